"""Helpers shared by the parity tests: run the same case through two libraries exporting the C ABI."""
import numpy as np

from open_ludwig_b200 import cabi


def default_params(level_dims, strict=1, **kw):
    p = dict(c_wale=0.5, nu_sgs_bg=0.0005, inlet_turbulence=0.01, q_min_threshold=0.001, wall_model_active=0,
             use_temporal=0, sponge_blend=1, symmetric=0, domain_nx=level_dims[0], domain_ny=level_dims[1],
             domain_nz=level_dims[2], strict_fp=strict)
    p.update(kw)
    return cabi.Params(**p)


def load_state(ctx, level, f, rho, vel):
    ctx.upload(level, cabi.F, f)
    ctx.upload(level, cabi.F_TEMP, f)
    ctx.upload(level, cabi.VEL, vel)
    ctx.upload(level, cabi.VEL_TEMP, vel)
    ctx.upload(level, cabi.RHO, rho)


def fetch_state(ctx, level):
    return {n: ctx.download(level, w) for n, w in (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("rho", cabi.RHO),
                                                    ("vel", cabi.VEL), ("vel_temp", cabi.VEL_TEMP))}


def max_ulp_diff(a, b):
    """Largest difference in units of the last place between two float32 arrays (same-sign finite values)."""
    ai = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return int(np.abs(ai - bi).max())


def rel_err_rho_u(ref, got):
    """north_star's error metrics: max|d rho|/rho and max|d u| / max|u_ref|."""
    e_rho = float(np.max(np.abs(got["rho"] - ref["rho"]) / np.abs(ref["rho"])))
    umax = float(np.max(np.abs(ref["vel"]))) or 1.0
    e_u = float(np.max(np.abs(got["vel"] - ref["vel"])) / umax)
    return e_rho, e_u
