"""A/B of library options on the synthetic single-level box through the C ABI, all variants in ONE process on one GPU
(development tool; bench.py is the contract).  The state is the device-side uniform flow (no 34 GB host upload), so the numbers
compare variants with each other, they are not bench values.

  python tools/ab_box.py --nb 64 --steps 30  "name|strict|block_order=xslab8,strict_loop=4"  "base|fast|" ...
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--nb", type=int, default=64)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--warmup", type=int, default=6)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--profile-steps", type=int, default=6)
ap.add_argument("--e2e-steps", type=int, default=20)
ap.add_argument("variants", nargs="+")
a = ap.parse_args()
lv = syn.make_box_level(a.nb, a.nb, a.nb)
n = a.nb * 8
by_ctx = {}
for v in a.variants:                         # variants that differ only in run-time options share a context (and its level tables)
    name, fp, opts = v.split("|")
    o = dict(kv.split("=", 1) for kv in opts.split(",") if kv)
    early = tuple(sorted((k, x) for k, x in o.items() if k in ("block_order", "partition", "halo_mirror")))
    by_ctx.setdefault(early, []).append((name, fp, {k: x for k, x in o.items() if (k, x) not in early}))
DEFAULTS = {"strict_loop": "1", "strict_occupancy": "5", "cta_threads": "auto", "prefetch_distance": "0", "strict_kernel": "reg", "fast_kernel": "direct"}
for early, vs in by_ctx.items():
    with cabi.Context(options=dict(early)) as c:
        c.add_level(lv)
        c.init_uniform_flow(0.03)
        t = 1
        for rep in range(a.repeat):
            for name, fp, o in vs:
                for k, x in {**DEFAULTS, **o}.items():
                    c.set_option(k, x)
                p = cabi.Params(c_wale=0.5, nu_sgs_bg=0.0005, inlet_turbulence=0.01, q_min_threshold=0.001, wall_model_active=0, use_temporal=0,
                                sponge_blend=1, symmetric=0, domain_nx=n, domain_ny=n, domain_nz=n, strict_fp=int(fp == "strict"))
                c.step_batch(t, a.warmup, 0.03, p); t += a.warmup; c.sync()
                t0 = time.perf_counter(); c.step_batch(t, a.steps, 0.03, p); c.sync(); dt = time.perf_counter() - t0; t += a.steps
                k_ms, k_l, k_cells, classes = 0.0, 0, 0, {}
                if a.profile_steps:
                    c.profile_enable(True); c.step_batch(t, a.profile_steps, 0.03, p); t += a.profile_steps
                    k_ms, k_l, k_cells = c.profile_read(); classes = {k: round(v / a.profile_steps, 3) for k, v in c.profile_classes().items() if v}; c.profile_enable(False)
                t0 = time.perf_counter()
                for _ in range(a.e2e_steps):
                    c.step_batch(t, 1, 0.03, p); t += 1; st = c.flow_stats(0)
                e2e = (time.perf_counter() - t0) / max(a.e2e_steps, 1)
                print(f"AB {name:28s} {fp:6s} early={dict(early)} opts={o} rep={rep} ms/step={dt / a.steps * 1e3:.3f} MLUPS={lv.n_cells * a.steps / dt / 1e6:.0f} "
                      f"plain_kernel_ms={k_ms / max(k_l, 1):.3f} frac216={(k_cells / max(k_l, 1)) * 216 / (max(k_ms, 1e-9) / max(k_l, 1) * 1e-3) / 1e9 / 6456.5:.3f} classes={classes} e2e_ms={e2e * 1e3:.3f}", flush=True)
        print("stats", c.flow_stats(0), flush=True)
