"""Multi-GPU correctness check (run under torchrun): the partitioned run must reproduce the single-GPU run bit for
bit in fast mode (same kernels, same per-cell arithmetic; only WHERE a neighbour block lives changes).
  - synthetic single-level box (periodic y/z, open x)
  - synthetic two-level case with sphere, Bouzidi, wall model, sponge, interface interpolation, forces
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from open_ludwig_b200 import cabi, multigpu as mg
from open_ludwig_b200.host import synthetic as syn
from util import default_params
import test_k1_features_gpu as T

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
# one GPU per rank when the box has them (NCCL plumbing, NVLink peer pulls, native barrier); otherwise the ranks SHARE cuda:0
# (gloo plumbing, CUDA-IPC mappings of the other process's allocations on the same device, blocking host barrier): the same
# one-process-per-GPU code path, testable on a one-GPU box
SHARED = torch.cuda.device_count() < world
lr = 0 if SHARED else lr
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if SHARED:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=dev)
BARRIERS = ("host",) if SHARED else ("native", "nccl", "native-mirror")
if rank == 0: print(f"mg_check: world {world}, {'shared cuda:0 (gloo + host barrier)' if SHARED else 'one GPU per rank (nccl)'}", flush=True)

def gather_field(ctx, level, which, lv):
    """global field assembled from every rank's local blocks"""
    a = torch.from_numpy(ctx.download(level, which)).to(mg._comm_device(dev))      # zeros outside the local blocks
    dist.all_reduce(a)
    return a.cpu().numpy()

def run_box(partitioned, steps=9, barrier="native", mirror=False, strict=0):
    dims = (6, 4, 4)
    lv = syn.make_box_level(*dims)
    f, rho, vel = syn.noise_state(lv)
    p = default_params(tuple(8 * d for d in dims), strict=strict)
    ctx = mg.init_context(None, lr) if partitioned else cabi.Context(device=lr)
    if mirror: ctx.set_option("halo_mirror", 1)
    ctx.add_level(lv)
    if partitioned:
        mg.attach_peers(ctx, dev, barrier)
        loc = ctx.local_blocks(0)
        from open_ludwig_b200 import partition
        assert np.array_equal(loc, partition.local_blocks(lv.active_block_coords, rank, world, level=lv)), "library partition != host mirror"
        ctx.upload_local(0, cabi.F, f[:, loc]); ctx.upload_local(0, cabi.F_TEMP, f[:, loc])
        ctx.upload_local(0, cabi.VEL, vel[:, loc]); ctx.upload_local(0, cabi.VEL_TEMP, vel[:, loc]); ctx.upload_local(0, cabi.RHO, rho[loc])
    else:
        for w, a in ((cabi.F, f), (cabi.F_TEMP, f), (cabi.VEL, vel), (cabi.VEL_TEMP, vel), (cabi.RHO, rho)):
            ctx.upload(0, w, a)
    ctx.step_batch(1, steps, 0.03, p); ctx.sync(); dist.barrier()
    if partitioned:
        out = {n: gather_field(ctx, 0, w, lv) for n, w in (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("rho", cabi.RHO), ("vel", cabi.VEL))}
        st = mg.reduce_stats(ctx.flow_stats(0), dev)
    else:
        out = {n: ctx.download(0, w) for n, w in (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("rho", cabi.RHO), ("vel", cabi.VEL))}
        st = ctx.flow_stats(0)
    dist.barrier(); ctx.close()
    return out, st

def run_two_level(partitioned, steps=40, plan=False, strict=0):
    levels = T.build_case()
    cells = tuple(8 * d for d in T.DIMS)
    p = default_params(cells, strict=strict, wall_model_active=1, use_temporal=1, inlet_turbulence=0.02)
    ctx = mg.init_context(None, lr) if partitioned else cabi.Context(device=lr)
    if partitioned and plan in ("rcb", "rcb_yz"): ctx.set_option("partition", plan)
    if partitioned and plan is True:
        ctx.set_partition_plan(levels)
    for lv in levels:
        ctx.add_level(lv)
    if partitioned:
        mg.attach_peers(ctx, dev, "host" if SHARED else None)
    ctx.init_equilibrium()
    centers, nrm, areas = T.sphere_mesh()
    mesh = ctx.create_mesh(centers, nrm, areas)
    forces = ctx.create_forces(mesh, 1.225, 10.0, 1.0, 1.0, (20.0, 16.0, 16.0), False)
    ctx.step_batch(1, steps, 0.02, p); ctx.sync(); dist.barrier()
    assert ctx.self_check() == 0, "index-table self-check failed"
    aero = ctx.compute_aerodynamics(forces, 1, (0.0, 0.0, 0.0), 300.0, 1.225, 5)
    if partitioned:
        aero = mg.reduce_aero(aero, dev)
        out = {f"L{i}{n}": gather_field(ctx, i, w, levels[i]) for i in range(2) for n, w in (("f", cabi.F), ("rho", cabi.RHO), ("vel", cabi.VEL))}
    else:
        out = {f"L{i}{n}": ctx.download(i, w) for i in range(2) for n, w in (("f", cabi.F), ("rho", cabi.RHO), ("vel", cabi.VEL))}
    dist.barrier(); ctx.close()
    return out, aero

ok = True
for strict in (0, 1):
    ref, sref = run_box(False, strict=strict)
    for barrier in BARRIERS:   # peer-flag barrier kernel / NCCL callback / packed halo mirrors (opt-in) / blocking host barrier
        got, sgot = run_box(True, barrier=barrier.split("-")[0], mirror=barrier.endswith("mirror"), strict=strict)
        for k in ref:
            same = np.array_equal(ref[k].view(np.int32), got[k].view(np.int32))
            ok &= same
            if rank == 0: print(f"box strict={strict} barrier={barrier} {k}: bit-identical={same} maxdiff={np.abs(ref[k]-got[k]).max():.3e}", flush=True)
    if rank == 0: print("box stats", sref["n_fluid"] == sgot["n_fluid"], abs(sref["rho_mean"] - sgot["rho_mean"]) < 1e-12, sref["rho_min"] == sgot["rho_min"], flush=True)
    ref, aref = run_two_level(False, strict=strict)
    for plan in (False, True, "rcb", "rcb_yz"):   # per-level cost-weighted Morton cut, the spatially aligned plan, per-level RCB boxes (all axes / y and z only)
        got, agot = run_two_level(True, plan=plan, strict=strict)
        for k in ref:
            same = np.array_equal(ref[k].view(np.int32), got[k].view(np.int32))
            ok &= same
            if rank == 0: print(f"two-level strict={strict} plan={plan} {k}: bit-identical={same} maxdiff={np.abs(ref[k]-got[k]).max():.3e}", flush=True)
        ok &= abs(aref["Cd"] - agot["Cd"]) <= 1e-9 * abs(aref["Cd"]) + 1e-15
        if rank == 0: print(f"aero strict={strict} plan={plan} Cd", aref["Cd"], agot["Cd"], "rel", abs(aref["Cd"] - agot["Cd"]) / abs(aref["Cd"]), flush=True)
if rank == 0:
    print("MG_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
