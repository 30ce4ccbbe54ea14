"""Generates tests/golden/*.json by running the CPU oracle (oracle/_build/libludwig_oracle.so) on the reference's
own case files.  Run in the build container (needs /root/reference or baseline/_ref/CASES):

    python tools/make_golden.py sphere_re1m 1000     # RESULTS_SPHERE_RE1M.txt configuration, rows at 200..1000
    python tools/make_golden.py ball1m_coarse 500    # SURVEY §8(d) config 1 (single level, res 7)
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir
from open_ludwig_b200.solver import Simulation

name, steps = sys.argv[1], int(sys.argv[2])
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "oracle", "_build", "libludwig_oracle.so")
case, ov = CASE_OVERRIDES[name]
dom = D.load_case(case_dir(case), ov, verbose=True)
sim = Simulation(dom, lib, strict=True)
print("backend", sim.ctx.backend, "cells", dom.total_cells, flush=True)
t0 = time.time()
def show(r):
    print(f"{r.step:6d} u={r.u_inlet:.6f} rho_min={r.rho_min:.6f} Cd={r.aero['Cd']:.6f} Cl={r.aero['Cl']:.6f} [{time.time()-t0:.0f}s]", flush=True)
rows = sim.run(steps, on_row=show)
out = {"case": name, "overrides": ov, "steps": steps, "backend": sim.ctx.backend,
       "reports": [r.__dict__ for r in dom.reports],
       "rows": [{"step": r.step, "u_inlet": r.u_inlet, "rho_min": r.rho_min, "stats": r.stats, "aero": r.aero} for r in rows]}
dst = os.path.join(ROOT, "tests", "golden", f"{name}_{sim.ctx.backend.replace('-', '_')}.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst)
