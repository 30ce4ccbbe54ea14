"""Checkpoint / restart of the solver state (SURVEY.md §8(f) N4 — the reference has none: its state lives only in
device arrays and the output directory is wiped at start, main.jl:79).

A checkpoint is the per-level state in the reference's own block-SoA layout (f, f_temp, vel, vel_temp, rho — exactly the
BlockLevel fields the A-B schedule reads) plus the step counter; it goes through ludwig_level_download / _upload, so it is
independent of the library's internal Morton / block-major layout and can be restored into a different build or GPU count.
Restart is bit-exact (tests/test_checkpoint_gpu.py).
"""
from __future__ import annotations

import numpy as np

from . import cabi

_FIELDS = (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("vel", cabi.VEL), ("vel_temp", cabi.VEL_TEMP), ("rho", cabi.RHO))


def save(path: str, ctx: cabi.Context, next_step: int) -> None:
    """Write the state of every level; `next_step` is the coarse step the run would execute next (main.jl:168 `t`)."""
    ctx.sync()
    out = {"next_step": np.int64(next_step), "n_levels": np.int64(len(ctx.n_blocks))}
    for lvl in range(len(ctx.n_blocks)):
        for name, which in _FIELDS:
            out[f"L{lvl}_{name}"] = ctx.download(lvl, which)
        if lvl + 1 < len(ctx.n_blocks):
            # levels with children keep the pre-step density for the temporal interface blend (blocks.jl:199-205)
            try:
                out[f"L{lvl}_rho_old"] = ctx.download(lvl, cabi.RHO_OLD)
            except cabi.LudwigError:
                pass
    np.savez(path, **out)


def load(path: str, ctx: cabi.Context) -> int:
    """Restore a checkpoint into a context whose levels were created from the same domain.  Returns next_step."""
    z = np.load(path)
    if int(z["n_levels"]) != len(ctx.n_blocks):
        raise ValueError("checkpoint has a different number of levels")
    for lvl in range(len(ctx.n_blocks)):
        for name, which in _FIELDS:
            ctx.upload(lvl, which, z[f"L{lvl}_{name}"])
    return int(z["next_step"])
