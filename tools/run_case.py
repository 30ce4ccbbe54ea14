"""Runs a named case (open_ludwig_b200/host/cases.py) through libludwig_b200.so and prints the diagnostics rows."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir
from open_ludwig_b200.solver import Simulation

ap = argparse.ArgumentParser()
ap.add_argument("name"); ap.add_argument("steps", type=int)
ap.add_argument("--strict", type=int, default=1); ap.add_argument("--lib", default=None); ap.add_argument("--json", default=None)
a = ap.parse_args()
case, ov = CASE_OVERRIDES[a.name]
t0 = time.time()
dom = D.load_case(case_dir(case), ov, verbose=True)
print(f"domain build {time.time()-t0:.1f}s cells {dom.total_cells/1e6:.2f}M updates/coarse step {dom.cell_updates_per_coarse_step/1e6:.1f}M", flush=True)
sim = Simulation(dom, a.lib, strict=bool(a.strict))
print("backend", sim.ctx.backend, "device MB", sim.ctx.device_bytes() / 1e6, flush=True)
t0 = time.time()
def show(r):
    print(f"{r.step:6d} u={r.u_inlet:.6f} rho_min={r.rho_min:.6f} Cd={r.aero['Cd']:.6f} Cl={r.aero['Cl']:.6f} Cmy={r.aero['Cmy']:.6f} [{time.time()-t0:.1f}s]", flush=True)
rows = sim.run(a.steps, on_row=show)
dt = time.time() - t0
print(f"{a.steps} steps in {dt:.2f}s: {dom.cell_updates_per_coarse_step*a.steps/dt/1e6:.0f} MLUPS true, {dom.total_cells*a.steps/dt/1e6:.0f} ref-MLUPS")
if a.json:
    json.dump({"case": a.name, "strict": a.strict, "rows": [{"step": r.step, "u_inlet": r.u_inlet, "rho_min": r.rho_min, "stats": r.stats, "aero": r.aero} for r in rows]}, open(a.json, "w"), indent=1)
