"""Where do fast and strict modes differ?  Runs a case in both modes and classifies the per-cell differences."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from open_ludwig_b200 import cabi
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir
from open_ludwig_b200.solver import Simulation
name = sys.argv[1]; steps = int(sys.argv[2])
case, ov = CASE_OVERRIDES[name]
dom = D.load_case(case_dir(case), ov)
out = []
for strict in (True, False):
    sim = Simulation(dom, None, strict=strict)
    sim.run(steps)
    out.append([{ "rho": sim.ctx.download(i, cabi.RHO), "vel": sim.ctx.download(i, cabi.VEL)} for i in range(len(dom.levels))])
    sim.close()
for i, lv in enumerate(dom.levels):
    a, b = out[0][i], out[1][i]
    drho = np.abs(a["rho"] - b["rho"]); dv = np.abs(a["vel"] - b["vel"]).max(axis=0)
    nt = lv.neighbor_table
    iface = (nt == 0).any(axis=0)                      # block has a missing neighbour
    cls = {"obstacle": lv.obstacle > 0, "nearwall": (lv.wall_dist < 10) & (lv.obstacle == 0), "sponge": lv.sponge > 0,
           "iface_block": np.broadcast_to(iface[:, None, None, None], lv.obstacle.shape) & (lv.obstacle == 0),
           "all": np.ones_like(lv.obstacle, bool)}
    sig = np.abs(a["rho"] - 1).max()
    print(f"L{i+1}: max|rho-1|={sig:.3e} max|u|={np.abs(a['vel']).max():.3e}")
    for k, m in cls.items():
        if m.any():
            print(f"   {k:12s} n={int(m.sum()):8d} max|drho|={drho[m].max():.3e} mean|drho|={drho[m].mean():.3e} max|du|={dv[m].max():.3e} mean|du|={dv[m].mean():.3e} mean(drho signed)={(b['rho']-a['rho'])[m].mean():+.3e}")
