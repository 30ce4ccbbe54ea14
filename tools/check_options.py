"""Bit-identity of option sets: runs the synthetic two-level feature case (tests/test_k1_features_gpu.py) and the noise box with the
library defaults and with each given option set, and compares every downloaded field word for word.

  python tools/check_options.py "block_order=xslab8" "block_order=xslab4,strict_loop=4"
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
import test_k1_features_gpu as F
from util import default_params, load_state, fetch_state

LIB = os.path.join(ROOT, "open_ludwig_b200", "csrc", "libludwig_b200.so")


def box(opts, strict, nb=(12, 10, 9), steps=6):
    lv = syn.make_box_level(*nb)
    f, rho, vel = syn.noise_state(lv)
    p = default_params(tuple(8 * d for d in nb), strict=strict)
    with cabi.Context(LIB, options=opts) as c:
        c.add_level(lv); load_state(c, 0, f, rho, vel)
        c.step_batch(1, steps, 0.03, p); c.sync()
        return {"L0": fetch_state(c, 0)}


def same(a, b):
    bad = []
    for lvl in a:
        for name in a[lvl]:
            if not np.array_equal(a[lvl][name].view(np.int32), b[lvl][name].view(np.int32)):
                bad.append((lvl, name, int((a[lvl][name].view(np.int32) != b[lvl][name].view(np.int32)).sum())))
    return bad


rc = 0
for strict in (1, 0):
    levels = F.build_case()
    base2 = F.run(LIB, levels, 10, strict, True)[0]
    baseb = box(None, strict)
    for spec in sys.argv[1:]:
        opts = dict(kv.split("=", 1) for kv in spec.split(",") if kv)
        bad = same(base2, F.run(LIB, levels, 10, strict, True, options=opts)[0]) + same(baseb, box(opts, strict))
        print(f"CHECK strict={strict} {spec}: {'IDENTICAL' if not bad else 'DIFFERENT ' + str(bad)}", flush=True)
        rc |= bool(bad)
sys.exit(rc)
