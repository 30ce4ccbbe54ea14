"""north_star's field-level parity bar, measured as it is stated: "for FP32, after 1000 steps the max relative error in rho
and u must be <= 1e-5", against the CPU oracle run here on the same inputs (7-15 s of host time).

  * BASELINE config 1: CASES/ball1m, coarsest single-level grid (72 x 64 x 64 cells, wall model, Bouzidi cells, sponge,
    all four kinds of domain face), driven exactly like main.jl (batched cosine ramp), 1000 steps;
  * the hashed-noise box of config 2 at 48^3, 1000 steps.

Metrics as north_star / VERDICT define them: max|d rho| / rho and max|d u| / max|u|.  The measured numbers are printed
(pytest -s, and into the assertion messages).

STRICT mode (the mode bench.py measures) is held to identity, not to a tolerance: same bits as the oracle after 1000 steps.
FAST mode (opt-in: FMA contraction, regrouped sums, MUFU) is held to the bound documented in DESIGN.md section 3: any two FP32
evaluation orders of this scheme drift apart by acoustic round-off noise of ~1 ulp of a population (4e-7 absolute in u),
measured here 2.4e-5 of max|u| on the noise box and 1.2e-4 on config 1 (max|u| = 0.018) — outside north_star's bar, which is why it
is not the benched mode.
"""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host import synthetic as syn
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir, have_case
from open_ludwig_b200.solver import Simulation
from util import default_params, load_state

pytestmark = pytest.mark.gpu
STEPS = 1000
FIELDS = (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("rho", cabi.RHO), ("vel", cabi.VEL), ("vel_temp", cabi.VEL_TEMP))


def errors(ref, got):
    """max|d rho|/rho, max|d u|/max|u| over both velocity buffers, max|d f|, number of differing words"""
    e_rho = float(np.max(np.abs(got["rho"] - ref["rho"]) / np.abs(ref["rho"])))
    umax = max(float(np.max(np.abs(ref["vel"]))), float(np.max(np.abs(ref["vel_temp"]))))
    e_u = max(float(np.max(np.abs(got[k] - ref[k]))) for k in ("vel", "vel_temp")) / umax
    e_f = max(float(np.max(np.abs(got[k] - ref[k]))) for k in ("f", "f_temp"))
    n_diff = sum(int(np.count_nonzero(got[k].view(np.int32) != ref[k].view(np.int32))) for k in ref)
    return e_rho, e_u, e_f, n_diff, umax


@pytest.fixture(scope="module")
def config1():
    if not have_case("ball1m"):
        pytest.skip("reference case files not available (tools/fetch_cases.py)")
    case, ov = CASE_OVERRIDES["ball1m_coarse"]
    return D.load_case(case_dir(case), ov, verbose=False)


def run_config1(dom, lib, strict):
    sim = Simulation(dom, lib, strict=strict)
    rows = sim.run(STEPS)
    out = {n: sim.ctx.download(0, w) for n, w in FIELDS}
    sim.close()
    return out, rows[-1]


def test_config1_1000_steps_strict_is_bit_identical(config1, oracle_lib, cuda_lib):
    ref, rref = run_config1(config1, oracle_lib, True)
    got, rgot = run_config1(config1, cuda_lib, True)
    e_rho, e_u, e_f, n_diff, umax = errors(ref, got)
    print(f"\nconfig 1, {STEPS} steps, STRICT vs oracle: e_rho={e_rho:.3e} e_u={e_u:.3e} max|df|={e_f:.3e} differing words={n_diff} (max|u|={umax:.4f})")
    assert n_diff == 0 and e_rho == 0.0 and e_u == 0.0, (e_rho, e_u, e_f, n_diff)
    assert rgot.rho_min == rref.rho_min
    for k in ("Cd", "Cl", "Cmy"):      # K3 maps are bit-identical; K4 sums them in FP64 (oracle: FP32 sequential, as the reference)
        assert rgot.aero[k] == pytest.approx(rref.aero[k], rel=2e-4, abs=1e-9), k


def test_config1_1000_steps_fast_within_documented_bound(config1, oracle_lib, cuda_lib):
    ref, rref = run_config1(config1, oracle_lib, True)
    got, rgot = run_config1(config1, cuda_lib, False)
    e_rho, e_u, e_f, n_diff, umax = errors(ref, got)
    print(f"\nconfig 1, {STEPS} steps, FAST vs oracle: e_rho={e_rho:.3e} e_u={e_u:.3e} max|df|={e_f:.3e} (max|u|={umax:.4f})")
    assert e_rho <= 1e-5, e_rho                                    # north_star's bar holds for rho
    assert e_u <= 3e-4, e_u                                        # documented fast-mode bound (NOT north_star's 1e-5); measured 1.2e-4
    assert e_u * umax <= 5e-6 and e_f <= 4e-6, (e_u * umax, e_f)   # i.e. a few ulp of a population, absolute
    assert rgot.aero["Cd"] == pytest.approx(rref.aero["Cd"], rel=1e-3)   # north_star: Cd within 0.1 %


@pytest.mark.parametrize("strict", [1, 0])
def test_noise_box_1000_steps(oracle_lib, cuda_lib, strict):
    dims = (6, 6, 6)
    lv = syn.make_box_level(*dims)
    state = syn.noise_state(lv)
    cells = tuple(8 * d for d in dims)
    out = {}
    for name, lib, s in (("ref", oracle_lib, 1), ("got", cuda_lib, strict)):
        with cabi.Context(lib) as c:
            c.add_level(lv)
            load_state(c, 0, *state)
            c.step_batch(1, STEPS, 0.03, default_params(cells, strict=s))
            c.sync()
            out[name] = {n: c.download(0, w) for n, w in FIELDS}
    e_rho, e_u, e_f, n_diff, umax = errors(out["ref"], out["got"])
    print(f"\nnoise box 48^3, {STEPS} steps, {'STRICT' if strict else 'FAST'} vs oracle: e_rho={e_rho:.3e} e_u={e_u:.3e} max|df|={e_f:.3e} differing words={n_diff}")
    if strict:
        assert n_diff == 0
    else:
        assert e_rho <= 1e-5 and e_u <= 3e-4 and e_f <= 4e-6, (e_rho, e_u, e_f)      # measured 2.4e-5
