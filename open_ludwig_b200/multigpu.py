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


def attach_peers(ctx: cabi.Context, device: torch.device, barrier: str | None = None):
    """Exchange the IPC handles of every rank's state.  Call after the last ctx.add_level().

    barrier = "native" (default): the library's own peer-flag barrier kernel (flag stores over NVLink, stream-ordered,
    no host call per barrier).  barrier = "nccl": a stream-ordered NCCL all-reduce of one float registered through
    ludwig_set_barrier_callback (the round-1 path, kept for A/B measurements; LUDWIG_BARRIER=nccl selects it)."""
    import os
    barrier = barrier or os.environ.get("LUDWIG_BARRIER", "native")
    world = dist.get_world_size()
    mine = ctx.ipc_export()
    t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allt)
    ctx.ipc_attach(blob, len(mine))
    dist.barrier()          # every rank has opened every peer mapping before anyone steps
    if barrier == "native":
        return None
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
    """Combine per-rank ludwig_flow_stats results (diagnostics.jl:56-94 over the whole level)."""
    s = torch.tensor([stats["n_fluid"], stats["rho_mean"] * stats["n_fluid"], stats["kinetic_energy"]], dtype=torch.float64, device=device)
    mn = torch.tensor([stats["rho_min"]], dtype=torch.float64, device=device)
    mx = torch.tensor([stats["rho_max"], stats["v_max"]], dtype=torch.float64, device=device)
    dist.all_reduce(s); dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    n = float(s[0])
    return {"n_fluid": n, "rho_mean": float(s[1]) / max(n, 1.0), "rho_min": float(mn[0]), "rho_max": float(mx[0]),
            "v_max": float(mx[1]), "kinetic_energy": float(s[2])}


def reduce_aero(aero: dict, device: torch.device) -> dict:
    """Every output of ludwig_compute_aerodynamics is linear in the per-rank partial sums."""
    keys = list(aero)
    t = torch.tensor([aero[k] for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return dict(zip(keys, t.tolist()))
