"""Feasibility probe: CUDA IPC peer mapping between torchrun ranks + peer read bandwidth (development tool)."""
import ctypes as C, os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
n = 256 * 1024 * 1024 // 4
x = torch.full((n,), float(rank + 1), device="cuda")
class Handle(C.Structure):
    _fields_ = [("b", C.c_ubyte * 64)]
rt.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), Handle, C.c_uint]
rt.cudaIpcGetMemHandle.argtypes = [C.POINTER(Handle), C.c_void_p]
h = Handle()
rc = rt.cudaIpcGetMemHandle(C.byref(h), C.c_void_p(x.data_ptr()))
hb = torch.tensor(list(bytes(h.b)), dtype=torch.uint8, device="cuda")
allh = [torch.empty_like(hb) for _ in range(world)]
dist.all_gather(allh, hb)
peer = (rank + 1) % world
ph = Handle(); C.memmove(C.byref(ph), bytes(allh[peer].cpu().tolist()), 64)
pp = C.c_void_p()
can = C.c_int(0); rt.cudaDeviceCanAccessPeer(C.byref(can), lr, peer)
rc2 = rt.cudaIpcOpenMemHandle(C.byref(pp), ph, 1)
y = torch.empty(n, device="cuda")
torch.cuda.synchronize(); dist.barrier()
rt.cudaMemcpy(C.c_void_p(y.data_ptr()), pp, C.c_size_t(n * 4), 3)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5):
    rt.cudaMemcpyAsync(C.c_void_p(y.data_ptr()), pp, C.c_size_t(n * 4), 3, None)
torch.cuda.synchronize(); dt = (time.time() - t0) / 5
print(f"rank {rank}: get={rc} canAccessPeer={can.value} open={rc2} peer value={y[0].item()} (expect {peer+1}) peer copy {n*4/dt/1e9:.0f} GB/s", flush=True)
dist.barrier(); dist.destroy_process_group()
