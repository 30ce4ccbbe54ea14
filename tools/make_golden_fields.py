"""Field-level golden fixtures of the SHIPPED full-size cases (BASELINE configs 4 and 5: CASES/Wing_5_deg, CASES/Stanford_bunny),
made offline with the CPU oracle (oracle/_build/libludwig_oracle.so) on the reference's own case files:

    python tools/make_golden_fields.py wing5 3
    python tools/make_golden_fields.py bunny 3

A few coarse steps from the uniform-flow impulsive start (forces are O(10) at once, so Cd / Cl / Cmy compare a developed force),
then per level: SHA-256 of the rho and vel arrays in the reference layout (the strict CUDA build is held to IDENTITY), their
Float64 sums and extrema (what a tolerance comparison of the fast build uses), plus the coefficients and the flow statistics.
One coarse step of these cases is 1.0-1.1 G cell updates: minutes per step on the CPU, hence offline fixtures
(tests/golden/<case>_fields_cpu_oracle.json, read by tests/test_full_size_cases_gpu.py).
"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from open_ludwig_b200 import cabi
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir
from open_ludwig_b200.solver import make_params


def field_record(ctx, n_levels):
    out = []
    for i in range(n_levels):
        rho = ctx.download(i, cabi.RHO); vel = ctx.download(i, cabi.VEL); velt = ctx.download(i, cabi.VEL_TEMP)
        out.append({"rho_sha256": hashlib.sha256(rho.tobytes()).hexdigest(), "vel_sha256": hashlib.sha256(vel.tobytes()).hexdigest(),
                    "vel_temp_sha256": hashlib.sha256(velt.tobytes()).hexdigest(),
                    "rho_sum": float(rho.astype(np.float64).sum()), "rho_min": float(rho.min()), "rho_max": float(rho.max()),
                    "vel_abs_sum": float(np.abs(vel.astype(np.float64)).sum()), "vel_max": float(np.abs(vel).max()),
                    "vel_temp_abs_sum": float(np.abs(velt.astype(np.float64)).sum())})
    return out


def run(name, steps, lib, strict=True):
    case, ov = CASE_OVERRIDES[name]
    dom = D.load_case(case_dir(case), ov, verbose=True, build_tri_map=False)
    p = dom.params
    params = make_params(dom, strict=strict)
    t0 = time.time()
    with cabi.Context(lib) as c:
        for lv in dom.levels:
            c.add_level(lv)
        mesh = c.create_mesh(dom.mesh.centers, dom.mesh.normals, dom.mesh.areas)
        forces = c.create_forces(mesh, p.rho_physical, p.u_physical, p.reference_area, p.reference_chord, p.moment_center, dom.cfg.symmetric)
        c.init_uniform_flow(float(dom.cfg.u_target))
        for t in range(1, steps + 1):
            c.step_batch(t, 1, float(dom.cfg.u_target), params)
            c.sync()
            print(f"step {t} [{time.time() - t0:.0f}s]", flush=True)
        aero = c.compute_aerodynamics(forces, len(dom.levels) - 1, p.mesh_offset, p.velocity_scale, p.rho_physical, 5)
        stats = c.flow_stats(0)
        rec = {"case": name, "steps": steps, "backend": c.backend, "initial_state": "ludwig_init_uniform_flow(u_target)", "u_inlet": float(dom.cfg.u_target),
               "blocks": [lv.n_blocks for lv in dom.levels], "cells": dom.total_cells, "aero": aero, "stats": stats, "levels": field_record(c, len(dom.levels))}
    return rec


if __name__ == "__main__":
    name, steps = sys.argv[1], int(sys.argv[2])
    lib = os.path.join(ROOT, "oracle", "_build", "libludwig_oracle.so")
    rec = run(name, steps, lib)
    dst = os.path.join(ROOT, "tests", "golden", f"{name}_fields_cpu_oracle.json")
    json.dump(rec, open(dst, "w"), indent=1)
    print("wrote", dst, "Cd", rec["aero"]["Cd"], "Cl", rec["aero"]["Cl"])
