"""Checkpoint / restart of the solver state (SURVEY.md §8(f) N4 — the reference has none: its state lives only in
device arrays and the output directory is wiped at start, main.jl:79).

A checkpoint is a DIRECTORY: ``meta.json`` plus one ``shard_<rank>.npz`` per rank.  A shard holds, per level, the reference
(1-based table order, here 0-based) indices of the rank's blocks and their state in the reference's block-SoA layout
(f, f_temp, vel, vel_temp, rho — exactly the BlockLevel fields the A-B schedule reads), moved with
ludwig_level_download_local / _upload_local: no rank ever holds more than its own blocks on the host (a 512^3-per-GPU box
has 116 GB of populations on 8 GPUs).  Blocks are keyed by reference index, so a checkpoint restores into ANY rank count or
partition rule — 8 ranks -> 1 GPU, Morton -> RCB, one process per GPU -> one ludwig_multi.  The pre-step density the
temporal interface blend reads is not part of it: a restart happens between coarse steps, where every level's "old" density
is its current one.  Restart is bit-exact (tests/test_checkpoint_gpu.py).
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from . import cabi

_FIELDS = (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("vel", cabi.VEL), ("vel_temp", cabi.VEL_TEMP), ("rho", cabi.RHO))
FORMAT = 2


def _rank_views(ctx):
    """The per-rank contexts this process drives: [ctx] for a Context, every rank of a MultiContext."""
    if isinstance(ctx, cabi.MultiContext):
        return [ctx.rank_ctx(r) for r in range(ctx.world)]
    return [ctx]


def save(path: str, ctx, next_step: int) -> None:
    """Write this process's shard(s).  `next_step` is the coarse step the run would execute next (main.jl:168 `t`).  With
    one process per GPU every rank calls this with the same `path` (a shared directory); rank 0 also writes meta.json."""
    ctx.sync()
    os.makedirs(path, exist_ok=True)
    n_levels = len(ctx.n_blocks)
    for c in _rank_views(ctx):
        out = {}
        for lvl in range(n_levels):
            loc = c.local_blocks(lvl)
            out[f"L{lvl}_blocks"] = loc.astype(np.int32)
            for name, which in _FIELDS:
                out[f"L{lvl}_{name}"] = c.download_local(lvl, which, len(loc))
        tmp = os.path.join(path, f".shard_{c.rank}.tmp.npz")
        np.savez(tmp, **out)
        os.replace(tmp, os.path.join(path, f"shard_{c.rank}.npz"))      # a crash never leaves a half-written shard under the final name
        if c.rank == 0:
            with open(os.path.join(path, "meta.json"), "w") as fh:
                json.dump({"format": FORMAT, "next_step": int(next_step), "n_levels": n_levels, "world": int(c.world),
                           "n_blocks": [int(n) for n in ctx.n_blocks]}, fh)


def load(path: str, ctx) -> int:
    """Restore a checkpoint into a context (any rank count / partition) whose levels were created from the same domain.
    Returns next_step.  Raises if a block of this context is in no shard (incomplete checkpoint)."""
    with open(os.path.join(path, "meta.json")) as fh:
        meta = json.load(fh)
    if meta.get("format") != FORMAT:
        raise ValueError(f"unsupported checkpoint format {meta.get('format')!r}")
    if meta["n_levels"] != len(ctx.n_blocks) or list(meta["n_blocks"]) != [int(n) for n in ctx.n_blocks]:
        raise ValueError("checkpoint was written for a different domain (levels / blocks per level differ)")
    shards = sorted(glob.glob(os.path.join(path, "shard_*.npz")))
    if len(shards) != meta["world"]:
        raise ValueError(f"checkpoint has {len(shards)} shards, meta.json says {meta['world']}")
    files = [np.load(s) for s in shards]
    for c in _rank_views(ctx):
        for lvl in range(meta["n_levels"]):
            want = c.local_blocks(lvl)
            # where every wanted block sits: (shard, row)
            src_shard = np.full(len(want), -1, np.int64); src_row = np.zeros(len(want), np.int64)
            pos = {int(b): i for i, b in enumerate(want)}
            for si, z in enumerate(files):
                have = z[f"L{lvl}_blocks"]
                idx = np.fromiter((pos.get(int(b), -1) for b in have), np.int64, len(have))
                m = idx >= 0
                src_shard[idx[m]] = si; src_row[idx[m]] = np.nonzero(m)[0]
            if (src_shard < 0).any():
                raise ValueError(f"level {lvl}: {int((src_shard < 0).sum())} blocks of rank {c.rank} are in no shard")
            for name, which in _FIELDS:
                ncomp = 1 if name == "rho" else (3 if name.startswith("vel") else 27)
                buf = np.empty((len(want), 8, 8, 8) if ncomp == 1 else (ncomp, len(want), 8, 8, 8), np.float32)
                for si, z in enumerate(files):
                    m = src_shard == si
                    if not m.any():
                        continue
                    a = z[f"L{lvl}_{name}"]
                    if ncomp == 1:
                        buf[m] = a[src_row[m]]
                    else:
                        buf[:, m] = a[:, src_row[m]]
                c.upload_local(lvl, which, buf)
    return int(meta["next_step"])
