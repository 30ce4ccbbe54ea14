"""One-process-per-GPU plumbing around the C ABI's multi-GPU entry points (include/ludwig_b200.h, "multi-GPU").

torch.distributed is used for plumbing only: the all-gather of the CUDA-IPC handles and the final reductions of partial
statistics / forces.  The data path itself has no collective: K1 pulls remote neighbour blocks through NVLink peer
mappings inside the kernel, and the cross-rank barrier after every level step is the library's own peer-flag kernel
(an NCCL all-reduce can be registered instead for A/B runs).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import cabi


def init_context(lib_path=None, local_rank=None) -> cabi.Context:
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = rank if local_rank is None else local_rank
    ctx = cabi.Context(lib_path, local_rank)
    ctx.set_partition(rank, world)
    return ctx


def _comm_device(device):
    """Tensors of the plumbing collectives live on the GPU for NCCL and on the host for gloo (several ranks sharing one GPU:
    NCCL refuses duplicate devices)."""
    return device if dist.get_backend() == "nccl" else torch.device("cpu")


def attach_peers(ctx: cabi.Context, device: torch.device, barrier: str | None = None):
    """Exchange the IPC handles of every rank's state.  Call after the last ctx.add_level().

    barrier = "native" (default): the library's own peer-flag barrier kernel (flag stores over NVLink, stream-ordered,
    no host call per barrier).  "nccl": a stream-ordered NCCL all-reduce of one float registered through
    ludwig_set_barrier_callback (the round-1 path, kept for A/B measurements).  "host": a blocking host barrier (stream
    synchronise + dist.barrier()) — slow, but the only safe choice when several ranks share ONE GPU (a spinning barrier
    kernel would wait for a peer whose kernels cannot run beside it); used to test the one-process-per-GPU path on a
    one-GPU box."""
    barrier = barrier or "native"
    world = dist.get_world_size()
    mine = ctx.ipc_export()
    cd = _comm_device(device)
    t = torch.tensor(list(mine), dtype=torch.uint8, device=cd)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allt)
    ctx.ipc_attach(blob, len(mine))
    dist.barrier()          # every rank has opened every peer mapping before anyone steps
    if barrier == "native":
        return None
    if barrier == "host":
        def host_barrier():
            torch.cuda.synchronize(device)
            dist.barrier()
        ctx.set_barrier(host_barrier)
        return host_barrier
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=device)
    flag = torch.zeros(1, device=device)

    def barrier():
        # ordered on the library's stream: the all-reduce starts when this rank's kernels so far are done, and
        # later launches on that stream wait for it.  No host synchronisation.
        with torch.cuda.stream(stream):
            dist.all_reduce(flag)

    ctx._mg_keepalive = (stream, flag)
    ctx.set_barrier(barrier)
    return barrier


def reduce_stats(stats: dict, device: torch.device) -> dict:
    """Combine per-rank ludwig_flow_stats results (diagnostics.jl:56-94 over the whole level): ONE all-gather of the six partial
    values per rank and a host-side reduction in rank order (sums are deterministic; a NaN density / velocity stays NaN as in the
    library and in Julia's minimum / maximum, which MIN / MAX all-reduces would drop)."""
    device = _comm_device(device)
    mine = torch.tensor([stats["n_fluid"], stats["rho_mean"] * stats["n_fluid"], stats["kinetic_energy"], stats["rho_min"], stats["rho_max"],
                         stats["v_max"]], dtype=torch.float64, device=device)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    a = torch.stack(parts).cpu().numpy()
    n = float(a[:, 0].sum())
    nan_rho = bool(np.isnan(a[:, 3]).any() or np.isnan(a[:, 4]).any())
    nan_v = bool(np.isnan(a[:, 5]).any())
    return {"n_fluid": n, "rho_mean": float(a[:, 1].sum()) / max(n, 1.0), "rho_min": float("nan") if nan_rho else float(a[:, 3].min()),
            "rho_max": float("nan") if nan_rho else float(a[:, 4].max()), "v_max": float("nan") if nan_v else float(a[:, 5].max()),
            "kinetic_energy": float(a[:, 2].sum())}


def reduce_aero(aero: dict, device: torch.device) -> dict:
    """Every output of ludwig_compute_aerodynamics is linear in the per-rank partial sums."""
    keys = list(aero)
    t = torch.tensor([aero[k] for k in keys], dtype=torch.float64, device=_comm_device(device))
    dist.all_reduce(t)
    return dict(zip(keys, t.tolist()))


def use_all_host_threads() -> int:
    """Size the OpenMP team of the host-side domain builder at run time: torchrun exports OMP_NUM_THREADS=1 to its workers and
    libgomp reads that when it is loaded, so a rank that builds the domain would otherwise do it on one core."""
    import ctypes
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return 1


def load_domain_shared(name: str, log=None):
    """The domain of a named case on every rank of the process group: rank 0 builds it with every host thread and the
    other ranks map the arrays from a private /dev/shm directory (mkdtemp, 0700, name broadcast by rank 0: nobody else can
    plant a pickle there).  Returns (domain, build seconds)."""
    import shutil
    import tempfile
    import time
    from .host import domain as D
    from .host.cases import CASE_OVERRIDES, case_dir
    mg = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(), dist.get_world_size()) if mg else (0, 1)
    case, ov = CASE_OVERRIDES[name]
    t0 = time.time()
    if rank == 0:
        from .host import domain as _d
        _d.host_lib()                 # load the builder (and its OpenMP runtime) ...
        use_all_host_threads()        # ... then give it every core this process may use
    if world > 1:
        box = [tempfile.mkdtemp(prefix="ludwig_domain_", dir="/dev/shm") if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        cache = box[0]
        if rank == 0:
            dom = D.load_case(case_dir(case), ov, verbose=False, build_tri_map=False)
            D.save_domain(dom, cache)
        dist.barrier()
        if rank != 0:
            dom = D.load_domain(cache)
        dist.barrier()
        if rank == 0:
            shutil.rmtree(cache, ignore_errors=True)     # the mappings stay valid after the unlink
    else:
        dom = D.load_case(case_dir(case), ov, verbose=False, build_tri_map=False)
    build_s = time.time() - t0
    if log is not None and rank == 0:
        log(f"[strong] {name}: domain build {build_s:.1f}s, {dom.total_cells / 1e6:.1f} M cells, {dom.cell_updates_per_coarse_step / 1e6:.0f} M updates per coarse step")
    return dom, build_s


def run_case_strong(name: str, steps: int, local_rank: int, *, strict: bool = False, options: dict | None = None, plan: bool = False,
                    ramp_steps: int | None = None, profile_steps: int = 0, log=None, dom=None, all_ranks_levels: bool = False,
                    uniform_start: bool = False, param_overrides: dict | None = None) -> dict:
    """One strong-scaling measurement of a named case (open_ludwig_b200.host.cases) over the ranks of the current process
    group (or a single GPU when torch.distributed is not initialised): every rank creates its partitioned context from the
    shared domain (load_domain_shared), attaches the peers and steps `steps` coarse steps along the driver's cosine ramp
    (main.jl:168-176, one batch per step).  Device time = max over ranks of CUDA events on the library's stream.
    uniform_start: instead of the rest state and the ramp, the whole domain starts as a uniform flow at the target inlet velocity
    (ludwig_init_uniform_flow) — an impulsive start, so that the body carries O(1) forces after a few coarse steps and the
    Cd / Cl printed for 1 and N GPUs compare a developed force, not round-off around zero.
    Returns the record (identical on every rank)."""
    import time
    from .solver import make_params, ramp_velocity

    mg = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(), dist.get_world_size()) if mg else (0, 1)
    dev = torch.device("cuda", local_rank)
    build_s = 0.0
    if dom is None:
        dom, build_s = load_domain_shared(name, log)
    t0 = time.time()
    ctx = cabi.Context(device=local_rank, options=options)
    try:
        if world > 1:
            ctx.set_partition(rank, world)
            if plan:
                ctx.set_partition_plan(dom.levels)
        for lv in dom.levels:
            ctx.add_level(lv)
        if world > 1:
            attach_peers(ctx, dev)
        m = dom.mesh
        mesh = ctx.create_mesh(m.centers, m.normals, m.areas)
        p = dom.params
        forces = ctx.create_forces(mesh, p.rho_physical, p.u_physical, p.reference_area, p.reference_chord, p.moment_center, dom.cfg.symmetric)
        if uniform_start:
            ctx.init_uniform_flow(float(dom.cfg.u_target))
        else:
            ctx.init_equilibrium()
        params = make_params(dom, strict=strict)
        for k, v in (param_overrides or {}).items():              # experiments only (e.g. wall_model_active=0 to time a feature alone)
            setattr(params, k, type(getattr(params, k))(v))
        ctx.sync()
        upload_s = time.time() - t0
        ramp = 1 if uniform_start else (ramp_steps or dom.cfg.ramp_steps)
        stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
        t = 1
        for _ in range(2):                                     # warm-up (builds the lazily built tables)
            ctx.step_batch(t, 1, ramp_velocity(dom.cfg.u_target, t, ramp), params); t += 1
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ctx.launch_count()
        if mg:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            ctx.step_batch(t, 1, ramp_velocity(dom.cfg.u_target, t, ramp), params); t += 1
        e1.record(stream)
        ctx.sync()
        launches = ctx.launch_count() - n0
        tm = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if mg:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm[0]) / steps
        levels_ms = None
        if profile_steps > 0:                                  # per-level / per-class device time (separate, untimed pass)
            ctx.profile_enable(True)
            for _ in range(profile_steps):
                ctx.step_batch(t, 1, ramp_velocity(dom.cfg.u_target, t, ramp), params); t += 1
            ctx.sync()
            ctx.profile_read()
            levels_ms = [{k: v / profile_steps for k, v in d.items()} for d in ctx.profile_levels()]
            ctx.profile_enable(False)
        if mg:
            dist.barrier()
        aero = ctx.compute_aerodynamics(forces, len(dom.levels) - 1, p.mesh_offset, p.velocity_scale, p.rho_physical, 5)
        stats = ctx.flow_stats(0)
        if mg:
            aero = reduce_aero(aero, dev); stats = reduce_stats(stats, dev)
        blocks = [len(ctx.local_blocks(i)) for i in range(len(dom.levels))]
        all_levels = None
        if all_ranks_levels and mg and levels_ms is not None:
            box = [None] * world
            dist.all_gather_object(box, {"blocks": blocks, "levels_ms": levels_ms})
            all_levels = box
        gb = ctx.device_bytes() / 1e9
        if mg:
            dist.barrier()
    finally:
        ctx.close()
    upd = dom.cell_updates_per_coarse_step
    return {"case": name, "n_gpus": world, "cells": dom.total_cells, "levels": len(dom.levels), "cell_updates_per_coarse_step": upd,
            "coarse_steps": steps, "steps_run": t - 1, "ramp_steps": ramp, "initial_state": "uniform flow at u_target (impulsive start)" if uniform_start else "rest + cosine ramp", "ms_per_coarse_step": ms, "mlups_true": upd / (ms * 1e-3) / 1e6,
            "mlups_reference_style": dom.total_cells / (ms * 1e-3) / 1e6, "fp_mode": "strict" if strict else "fast",
            "Cd": aero["Cd"], "Cl": aero["Cl"], "rho_min": stats["rho_min"], "rho_max": stats["rho_max"], "v_max": stats["v_max"],
            "partition": (options or {}).get("partition", "plan" if plan else "morton"), "options": options or {},
            "blocks_rank0": blocks, "device_gb_rank0": gb, "launches_rank0": int(launches), "rank0_levels_ms": levels_ms, "all_ranks": all_levels,
            "domain_build_s": build_s, "upload_s": upload_s}
