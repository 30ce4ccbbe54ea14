"""Quick single-level K1 timing through the C ABI (development tool; bench.py is the contract)."""
import argparse, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--nb", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--strict", type=int, default=0)
ap.add_argument("--noise", type=int, default=1)
a = ap.parse_args()
t0 = time.time()
lv = syn.make_box_level(a.nb, a.nb, a.nb)
print(f"topology {time.time()-t0:.1f}s nb={lv.n_blocks} cells={lv.n_cells/1e6:.1f}M", flush=True)
n = a.nb * 8
p = cabi.Params(c_wale=0.5, nu_sgs_bg=0.0005, inlet_turbulence=0.01, q_min_threshold=0.001, wall_model_active=0,
                use_temporal=0, sponge_blend=1, symmetric=0, domain_nx=n, domain_ny=n, domain_nz=n, strict_fp=a.strict)
with cabi.Context() as c:
    t0 = time.time(); c.add_level(lv); print(f"add_level {time.time()-t0:.1f}s", flush=True)
    if a.noise:
        t0 = time.time(); f, rho, vel = syn.noise_state(lv); print(f"noise_state {time.time()-t0:.1f}s", flush=True)
        t0 = time.time()
        c.upload(0, cabi.F, f); c.upload(0, cabi.F_TEMP, f); c.upload(0, cabi.VEL, vel); c.upload(0, cabi.VEL_TEMP, vel); c.upload(0, cabi.RHO, rho)
        print(f"upload {time.time()-t0:.1f}s", flush=True)
    else:
        c.init_equilibrium()
    c.step_batch(1, a.warmup, 0.03, p); c.sync()
    t0 = time.time(); c.step_batch(1 + a.warmup, a.steps, 0.03, p); c.sync(); dt = time.time() - t0
    mlups = lv.n_cells * a.steps / dt / 1e6
    print(f"RESULT nb={a.nb} strict={a.strict} ms/step={dt/a.steps*1e3:.3f} MLUPS={mlups:.0f} GB/s@216={mlups*216e-3:.0f} GB/s@244={mlups*244e-3:.0f} dev_bytes={c.device_bytes()/1e9:.2f}GB")
    print(c.flow_stats(0))
