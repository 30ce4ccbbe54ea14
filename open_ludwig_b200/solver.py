"""Host driver: Python mirror of the kept Julia driver's ``solve_main`` loop (main.jl:54-249) on top of the C ABI.

It does exactly what main.jl does around the kernel boundary — upload (``adapt``), ``init_eq!``, the batched
step loop with the Float32 cosine ramp evaluated once per ``async_depth`` batch (main.jl:168-176), the
diagnostics / force cadence (main.jl:183-211) — and nothing else (no VTK, no CSV wiping).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

from . import cabi
from .host.domain import Domain


@dataclass
class DiagRow:
    """One row of the reference's console table / convergence.csv (main.jl:203-209) + forces.csv fields."""
    step: int
    u_inlet: float
    rho_min: float
    stats: dict
    aero: Optional[dict]
    wall_s: float


def ramp_velocity(u_target: float, batch_end: int, ramp_steps: int) -> float:
    """main.jl:173-174 in Float32: prog = 0.5f0*(1f0 - cos(Float32(pi)*batch_end/RAMP)), u = U_TARGET*prog."""
    f32 = np.float32
    if batch_end <= ramp_steps:
        arg = f32(f32(np.pi) * f32(batch_end)) / f32(ramp_steps)
        # Julia's cos(::Float32) is correctly rounded in practice; evaluate in double and round once
        prog = f32(0.5) * (f32(1.0) - f32(math.cos(float(arg))))
    else:
        prog = f32(1.0)
    return float(f32(u_target) * prog)


def make_params(domain: Domain, strict: bool) -> cabi.Params:
    cfg, p = domain.cfg, domain.params
    return cabi.Params(c_wale=cfg.c_wale, nu_sgs_bg=cfg.nu_sgs_background, inlet_turbulence=cfg.inlet_turbulence,
                       q_min_threshold=cfg.q_min_threshold, wall_model_active=int(p.wall_model_active),
                       use_temporal=int(cfg.temporal_interpolation), sponge_blend=int(cfg.sponge_blend_distributions),
                       symmetric=int(cfg.symmetric), domain_nx=p.nx_coarse, domain_ny=p.ny_coarse, domain_nz=p.nz_coarse,
                       strict_fp=int(strict))


class Simulation:
    """solve_main (main.jl:54-249) for one case on one device."""

    def __init__(self, domain: Domain, lib_path: Optional[str] = None, device: int = 0, strict: bool = False):
        self.domain = domain
        self.ctx = cabi.Context(lib_path, device)
        self.params = make_params(domain, strict)
        for lv in domain.levels:                      # main.jl:98
            self.ctx.add_level(lv)
        m = domain.mesh
        self.mesh = self.ctx.create_mesh(m.centers, m.normals, m.areas)          # main.jl:101
        self.ctx.init_equilibrium()                                                # main.jl:126-135
        p = domain.params
        self.forces = None
        if domain.cfg.force_enabled:                                               # main.jl:143-155
            self.forces = self.ctx.create_forces(self.mesh, p.rho_physical, p.u_physical, p.reference_area,
                                                 p.reference_chord, p.moment_center, domain.cfg.symmetric)
        self.t = 1
        self.rows: List[DiagRow] = []
        self._t0 = time.time()

    def close(self):
        self.ctx.close()

    def aerodynamics(self) -> dict:
        p = self.domain.params
        return self.ctx.compute_aerodynamics(self.forces, len(self.domain.levels) - 1, p.mesh_offset, p.velocity_scale,
                                             p.rho_physical, 5)

    def run(self, steps: Optional[int] = None, on_row: Optional[Callable[[DiagRow], None]] = None) -> List[DiagRow]:
        cfg = self.domain.cfg
        steps = cfg.steps if steps is None else steps
        batch = cfg.gpu_async_depth
        diag = cfg.diag_freq
        while self.t <= steps:                                      # main.jl:168-232
            batch_end = min(self.t + batch - 1, steps)
            actual = batch_end - self.t + 1
            u_curr = ramp_velocity(cfg.u_target, batch_end, cfg.ramp_steps)
            self.ctx.step_batch(self.t, actual, u_curr, self.params)
            if batch_end % diag < actual or batch_end == steps:
                diag_step = (batch_end // diag) * diag
                if self.t <= diag_step <= batch_end:
                    stats = self.ctx.flow_stats(0)
                    aero = self.aerodynamics() if self.forces is not None else None
                    row = DiagRow(diag_step, u_curr, stats["rho_min"], stats, aero, time.time() - self._t0)
                    self.rows.append(row)
                    if on_row:
                        on_row(row)
            self.t = batch_end + 1
        self.ctx.sync()
        return self.rows
