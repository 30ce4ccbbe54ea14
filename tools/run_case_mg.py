"""Strong scaling of a named case over N GPUs (one process per GPU; launch with torchrun, or plain python for N = 1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_case_mg.py bunny_fine 6

Every rank builds the (identical) domain on the host, creates a partitioned context, attaches the peers and steps in
lock-step; forces are reduced over the ranks.  Prints true MLUPS (max time over ranks, CUDA events on the library's stream).
"""
import os, sys, time
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
# host threads (set before anything loads an OpenMP runtime; torchrun exports OMP_NUM_THREADS=1): rank 0 builds the domain
# with every core, the other ranks only map the cached arrays
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 8) if rank == 0 else "2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from open_ludwig_b200 import cabi, multigpu as mg
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir
from open_ludwig_b200.solver import make_params, ramp_velocity

name, steps = sys.argv[1], int(sys.argv[2])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
case, ov = CASE_OVERRIDES[name]
t0 = time.time()
if world > 1:
    # rank 0 builds the domain with every host thread and caches it in shared memory; the others map the arrays
    cache = f"/dev/shm/ludwig_domain_{name}_{os.getppid()}"
    if rank == 0:
        dom = D.load_case(case_dir(case), ov, verbose=True, build_tri_map=False)
        D.save_domain(dom, cache)
    else:
        while not os.path.exists(os.path.join(cache, "domain.pkl")):
            time.sleep(0.5)
        dom = D.load_domain(cache)
else:
    dom = D.load_case(case_dir(case), ov, verbose=True, build_tri_map=False)
if rank == 0:
    print(f"domain build {time.time()-t0:.1f}s cells {dom.total_cells/1e6:.1f}M updates/coarse step {dom.cell_updates_per_coarse_step/1e6:.0f}M", flush=True)
ctx = cabi.Context(device=lr)
if world > 1:
    ctx.set_partition(rank, world)
    if not os.environ.get("LUDWIG_NO_PLAN"):
        ctx.set_partition_plan(dom.levels)       # parents, children and neighbours of one region on one GPU
t0 = time.time()
for lv in dom.levels:
    ctx.add_level(lv)
if world > 1:
    mg.attach_peers(ctx, dev)
m = dom.mesh
mesh = ctx.create_mesh(m.centers, m.normals, m.areas)
p = dom.params
forces = ctx.create_forces(mesh, p.rho_physical, p.u_physical, p.reference_area, p.reference_chord, p.moment_center, dom.cfg.symmetric)
ctx.init_equilibrium()
params = make_params(dom, strict=False)
ctx.sync()
if rank == 0:
    print(f"upload {time.time()-t0:.1f}s device GB (rank 0) {ctx.device_bytes()/1e9:.1f}", flush=True)
stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
u = ramp_velocity(dom.cfg.u_target, 8, dom.cfg.ramp_steps)
ctx.step_batch(1, 2, u, params); ctx.sync()            # warm-up (builds the fast-mode tables)


def timed(n):
    """max over ranks of the device time of n coarse steps (CUDA events on the library's stream)"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1: dist.barrier()
    e0.record(stream)
    ctx.step_batch(3, n, u, params)
    e1.record(stream)
    ctx.sync()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


ms = timed(steps)
if world > 1 and os.environ.get("LUDWIG_AB"):
    # A/B on the same processes: the round-1 barrier (stream-ordered NCCL all-reduce through the callback) against the
    # library's own peer-flag barrier kernel
    bstream, flag = stream, torch.zeros(1, device=dev)
    def nccl_barrier():
        with torch.cuda.stream(bstream):
            dist.all_reduce(flag)
    ctx.set_barrier(nccl_barrier)
    ctx.step_batch(1, 1, u, params); ctx.sync()
    ms_nccl = timed(steps)
    ctx.clear_barrier()
    if rank == 0:
        print(f"AB barrier: native {ms/steps:.2f} ms/step, nccl callback {ms_nccl/steps:.2f} ms/step", flush=True)
if os.environ.get("LUDWIG_PROFILE") is not None:   # per-level / per-class device time of every rank
    ctx.profile_enable(True)
    ctx.step_batch(3, steps, u, params); ctx.sync()
    ctx.profile_read(); lv = ctx.profile_levels()
    ctx.profile_enable(False)
    loc = [len(ctx.local_blocks(i)) for i in range(len(dom.levels))]
    print(f"rank {rank}: blocks {loc} per level [ms/coarse step]: " + " | ".join(
        f"L{i+1} {d['level_step']/steps:.2f} (k1p {d['k1_plain']/steps:.2f} pg {d['k1_plain_ghost']/steps:.2f} ft {d['k1_feature']/steps:.2f} fu {d['k1_full']/steps:.2f} pre {d['interface_prepass']/steps:.2f} bz {d['bouzidi']/steps:.2f} bar {d['barrier']/steps:.2f} unp {d['halo_unpack']/steps:.2f} pk {d['halo_pack']/steps:.2f})" for i, d in enumerate(lv)), flush=True)
if world > 1: dist.barrier()
aero = ctx.compute_aerodynamics(forces, len(dom.levels) - 1, p.mesh_offset, p.velocity_scale, p.rho_physical, 5)
stats = ctx.flow_stats(0)
if world > 1:
    aero = mg.reduce_aero(aero, dev); stats = mg.reduce_stats(stats, dev)
if rank == 0:
    sec = ms * 1e-3
    print(f"RESULT case={name} n_gpus={world} steps={steps} s/step={sec/steps:.4f} true_MLUPS={dom.cell_updates_per_coarse_step*steps/sec/1e6:.0f} "
          f"ref_MLUPS={dom.total_cells*steps/sec/1e6:.0f} Cd={aero['Cd']:.6e} Cl={aero['Cl']:.6e} rho_min={stats['rho_min']:.6f} rho_max={stats['rho_max']:.6f}", flush=True)
if world > 1: dist.barrier()
ctx.close()
if world > 1:
    dist.destroy_process_group()
    if rank == 0:
        import shutil
        shutil.rmtree(cache, ignore_errors=True)
