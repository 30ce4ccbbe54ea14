import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from open_ludwig_b200 import cabi, multigpu as mg
from open_ludwig_b200.host import synthetic as syn
from util import default_params
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
dims = (6, 4, 4)
lv = syn.make_box_level(*dims)
f, rho, vel = syn.noise_state(lv)
p = default_params(tuple(8 * d for d in dims), strict=0)
def single(steps):
    c = cabi.Context(device=lr); c.add_level(lv)
    for w, a in ((cabi.F, f), (cabi.F_TEMP, f), (cabi.VEL, vel), (cabi.VEL_TEMP, vel), (cabi.RHO, rho)): c.upload(0, w, a)
    c.step_batch(1, steps, 0.03, p); c.sync()
    o = {n: c.download(0, w) for n, w in (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("rho", cabi.RHO))}; c.close(); return o
def multi(steps, blocking):
    c = mg.init_context(None, lr); c.add_level(lv); mg.attach_peers(c, dev)
    if blocking:
        def bar():
            c.lib.ludwig_sync(c._h); dist.barrier()
        c.set_barrier(bar)
    loc = c.local_blocks(0)
    for w, a in ((cabi.F, f), (cabi.F_TEMP, f), (cabi.VEL, vel), (cabi.VEL_TEMP, vel)): c.upload_local(0, w, a[:, loc])
    c.upload_local(0, cabi.RHO, rho[loc])
    c.sync(); dist.barrier()
    c.step_batch(1, steps, 0.03, p); c.sync(); dist.barrier()
    o = {}
    for n, w in (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("rho", cabi.RHO)):
        a = torch.from_numpy(c.download(0, w)).to(dev); dist.all_reduce(a); o[n] = a.cpu().numpy()
    dist.barrier(); c.close(); return o
for steps in (1, 2, 9):
    r = single(steps)
    for blocking in (True, False):
        g = multi(steps, blocking)
        if rank == 0:
            bad = np.argwhere((r["rho"] != g["rho"]).any(axis=(1, 2, 3))).ravel()
            print(f"steps={steps} blocking={blocking}:", {k: bool(np.array_equal(r[k], g[k])) for k in r}, "bad blocks", bad[:12], len(bad), flush=True)
            if len(bad):
                b = bad[0]; d = np.argwhere(r["rho"][b] != g["rho"][b])
                print("   coords of block", lv.active_block_coords[b], "first bad cells zyx", d[:5].tolist(), flush=True)
dist.destroy_process_group()
