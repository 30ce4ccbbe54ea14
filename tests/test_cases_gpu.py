"""Case-level parity against committed oracle fixtures (tests/golden/*_cpu_oracle.json, made by tools/make_golden.py
with the CPU oracle on the reference's own case files):

  config 1  ball1m, coarsest single-level grid (surface_resolution 7, num_levels 1), 500 steps — wall model, Bouzidi on
            level 1, sponge, all four domain-face BCs
  config 4' Wing_5_deg at reduced resolution (3 levels, 2.65 M cells): symmetric half model (forces doubled, Fy = 0),
            WMLES, inlet turbulence hash, 36 871 Bouzidi cells — Cd / Cl / Cmy

strict: everything but powf/logf is bit-identical -> coefficients to ~1e-4 relative; fast: north_star's 0.1 %.
"""
import json
import os

import pytest

from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir, have_case
from open_ludwig_b200.solver import Simulation

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_case("ball1m"), reason="reference case files not available")]
HERE = os.path.dirname(os.path.abspath(__file__))


def _fixture(name):
    with open(os.path.join(HERE, "golden", f"{name}_cpu_oracle.json")) as fh:
        return json.load(fh)


def _run(name, lib, strict, steps):
    case, ov = CASE_OVERRIDES[name]
    dom = D.load_case(case_dir(case), ov)
    sim = Simulation(dom, lib, strict=strict)
    rows = {r.step: r for r in sim.run(steps)}
    sim.close()
    return dom, rows


@pytest.mark.parametrize("strict", [True, False])
def test_config1_ball1m_coarse(cuda_lib, strict):
    fx = _fixture("ball1m_coarse")
    dom, rows = _run("ball1m_coarse", cuda_lib, strict, fx["steps"])
    assert [r.n_blocks for r in dom.reports] == [r["n_blocks"] for r in fx["reports"]]
    assert dom.reports[0].n_boundary_cells == fx["reports"][0]["n_boundary_cells"] > 0
    for ref in fx["rows"]:
        g = rows[ref["step"]]
        assert g.u_inlet == ref["u_inlet"]
        assert g.stats["n_fluid"] == ref["stats"]["n_fluid"]
        assert abs(g.rho_min - ref["rho_min"]) <= (2e-6 if strict else 1e-5)
        if ref["step"] >= 300:          # earlier rows: |F| < 1e-3 of its final value, round-off dominated
            assert g.aero["Cd"] == pytest.approx(ref["aero"]["Cd"], rel=3e-4 if strict else 1e-3), ref["step"]


@pytest.mark.parametrize("strict", [True, False])
def test_config4_wing_reduced(cuda_lib, strict):
    if not have_case("Wing_5_deg"):
        pytest.skip("Wing_5_deg case files not available")
    fx = _fixture("wing5_small")
    dom, rows = _run("wing5_small", cuda_lib, strict, fx["steps"])
    assert dom.cfg.symmetric and [r.n_blocks for r in dom.reports] == [r["n_blocks"] for r in fx["reports"]]
    last = fx["rows"][-1]
    g = rows[last["step"]]
    assert g.stats["n_fluid"] == last["stats"]["n_fluid"]
    assert abs(g.rho_min - last["rho_min"]) <= (2e-6 if strict else 1e-5)
    for key in ("Cd", "Cl", "Cmy"):
        assert g.aero[key] == pytest.approx(last["aero"][key], rel=5e-4 if strict else 1e-3, abs=1e-7), key
    assert g.aero["Fy"] == 0.0 and g.aero["Mx"] == 0.0          # symmetry plane: side force and roll/yaw moments vanish
