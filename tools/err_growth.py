"""Error growth of the fast CUDA path vs the oracle (development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
from util import default_params, fetch_state, load_state, rel_err_rho_u
ORACLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build", "libludwig_oracle.so")
dims = (6, 6, 6)
amp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.003
lv = syn.make_box_level(*dims)
state = syn.noise_state(lv, amp_u=amp)
cells = tuple(8 * d for d in dims)
ctxs = []
for lib, strict in ((ORACLE, 1), (None, 1), (None, 0)):
    c = cabi.Context(lib); c.add_level(lv); load_state(c, 0, *state); ctxs.append((c, default_params(cells, strict=strict)))
t = 1
for n in (1, 1, 3, 5, 10, 30, 50, 100):
    outs = []
    for c, p in ctxs:
        c.step_batch(t, n, 0.03, p); c.sync(); outs.append(fetch_state(c, 0))
    t += n
    es = rel_err_rho_u(outs[0], outs[1]); ef = rel_err_rho_u(outs[0], outs[2])
    fe = float(np.max(np.abs(outs[0]["f"] - outs[2]["f"])))
    print(f"step {t-1:4d}: strict e_rho={es[0]:.2e} e_u={es[1]:.2e} | fast e_rho={ef[0]:.2e} e_u={ef[1]:.2e} max|df|={fe:.2e} umax={np.abs(outs[0]['vel']).max():.4f}")
