"""The partitioned (multi-GPU) code path, verified on ONE GPU: N virtual ranks of a ludwig_multi share cuda:0.

Every rank owns its range of blocks, pulls halo layers / parent cells / x_ff populations / surface cells of the other
ranks through the same peer tables a real N-GPU run uses (here the "peers" are other allocations on the same device),
and the cross-rank barriers are the group's stream-ordered event waits.  The partitioned run must reproduce the
single-context run BIT FOR BIT in both FP modes, for every partition rule and both halo strategies — only WHERE a
neighbour block lives changes, never the per-cell arithmetic.  (On a box with >= 2 GPUs tests/test_multigpu_gpu.py runs
the one-process-per-GPU flavour of the same check over NVLink.)
"""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
from util import default_params

import test_k1_features_gpu as T

pytestmark = pytest.mark.gpu

STEPS = 40      # the pressure wave of the impulsive start reaches the sphere after ~35 coarse steps: forces are O(1), not round-off
FIELDS = (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("rho", cabi.RHO), ("vel", cabi.VEL), ("vel_temp", cabi.VEL_TEMP))


def run_two_level(n_ranks, steps, strict, options=None, plan=False, levels=None):
    levels = levels or T.build_case()
    cells = tuple(8 * d for d in T.DIMS)
    p = default_params(cells, strict=strict, wall_model_active=1, use_temporal=1, inlet_turbulence=0.02)
    centers, nrm, areas = T.sphere_mesh()
    with cabi.MultiContext(n_ranks, devices=[0] * n_ranks, options=options) as m:
        if plan:
            m.set_partition_plan(levels)
        for lv in levels:
            m.add_level(lv)
        m.init_equilibrium()
        h = m.create_forces(centers, nrm, areas, 1.225, 10.0, 1.0, 1.0, (20.0, 16.0, 16.0), False)
        # two batches with force / statistics calls in between, as the driver's diagnostics cadence does (main.jl:183-211)
        m.step_batch(1, steps // 2, 0.02, p)
        mid = m.compute_aerodynamics(h, len(levels) - 1, (0.0, 0.0, 0.0), 300.0, 1.225, 5)
        m.step_batch(1 + steps // 2, steps - steps // 2, 0.02, p)
        m.sync()
        out = {f"L{i}{n}": m.download(i, w) for i in range(len(levels)) for n, w in FIELDS}
        aero = m.compute_aerodynamics(h, len(levels) - 1, (0.0, 0.0, 0.0), 300.0, 1.225, 5)
        maps = m.download_force_maps(h, len(areas))
        stats = [m.flow_stats(i) for i in range(len(levels))]
        owners = [[len(m.rank_ctx(r).local_blocks(i)) for r in range(n_ranks)] for i in range(len(levels))]
        assert m.self_check() == 0          # every index table / peer offset the kernels dereference is in range (host-side bounds checks)
    return out, aero, maps, stats, mid, owners


def assert_same(ref, got, what):
    (o0, a0, m0, s0, mid0, _), (o1, a1, m1, s1, mid1, own) = ref, got
    for k in o0:
        assert np.array_equal(o0[k].view(np.int32), o1[k].view(np.int32)), (what, k, float(np.abs(o0[k] - o1[k]).max()))
    for x, y in zip(m0, m1):
        assert np.array_equal(x.view(np.int32), y.view(np.int32)), (what, "force maps")
    for k in ("Fx", "Fy", "Fz", "Mx", "My", "Mz", "Cd", "Cl", "Cs", "Cmy"):   # FP64 sums of the same FP32 terms in another order
        assert a1[k] == pytest.approx(a0[k], rel=1e-12, abs=1e-18), (what, k)
        assert mid1[k] == pytest.approx(mid0[k], rel=1e-12, abs=1e-18), (what, "mid", k)
    for x, y in zip(s0, s1):
        assert x["n_fluid"] == y["n_fluid"] and x["rho_min"] == y["rho_min"] and x["rho_max"] == y["rho_max"] and x["v_max"] == y["v_max"]
        assert y["rho_mean"] == pytest.approx(x["rho_mean"], rel=1e-13)
        assert y["kinetic_energy"] == pytest.approx(x["kinetic_energy"], rel=1e-12)
    assert all(min(o) >= 1 for o in own), own


@pytest.fixture(scope="module")
def single():
    return {strict: run_two_level(1, STEPS, strict) for strict in (0, 1)}


@pytest.mark.parametrize("strict", [0, 1])
@pytest.mark.parametrize("n_ranks,options,plan", [
    (2, None, False),                          # cost-weighted Morton ranges per level
    (8, None, False),
    (8, None, True),                           # spatially aligned plan
    (8, {"partition": "rcb"}, False),          # per-level recursive coordinate bisection
    (8, {"partition": "rcb_yz"}, False),       # the same, never cutting across x
    (3, {"partition": "rcb_yz"}, False),
    (4, {"halo_mirror": 1}, False),            # packed halo exchange into local mirrors
    (4, {"remote_order": "interleave", "fork_max_blocks": 0}, False),
])
def test_virtual_ranks_two_level_bit_identical(single, strict, n_ranks, options, plan):
    """Two-level case with every feature (interfaces with temporal blend, sphere with Bouzidi links, wall model, sponge,
    domain faces, forces): N virtual ranks == 1 context, bit for bit, Cd/Cl to 1e-12."""
    got = run_two_level(n_ranks, STEPS, strict, options=options, plan=plan)
    assert_same(single[strict], got, (n_ranks, options, plan, strict))
    assert abs(got[1]["Cd"]) > 1e-3          # the comparison is about a developed force, not about noise around zero


@pytest.mark.parametrize("strict", [0, 1])
def test_virtual_ranks_box_with_noise_state(strict):
    """Single-level box (open x faces, periodic y/z: every block has neighbours on another rank through the wrap), hashed
    initial state uploaded in the reference layout: 1, 2 and 8 ranks give the same bits after 9 steps."""
    dims = (8, 4, 4)
    lv = syn.make_box_level(*dims)
    f, rho, vel = syn.noise_state(lv)
    p = default_params(tuple(8 * d for d in dims), strict=strict)
    res = {}
    for n in (1, 2, 8):
        with cabi.MultiContext(n, devices=[0] * n) as m:
            m.add_level(lv)
            for w, a in ((cabi.F, f), (cabi.F_TEMP, f), (cabi.VEL, vel), (cabi.VEL_TEMP, vel), (cabi.RHO, rho)):
                m.upload(0, w, a)
            m.step_batch(1, 9, 0.03, p)
            m.sync()
            res[n] = {k: m.download(0, w) for k, w in FIELDS}, m.flow_stats(0)
    for n in (2, 8):
        for k in res[1][0]:
            assert np.array_equal(res[1][0][k].view(np.int32), res[n][0][k].view(np.int32)), (n, k)
        assert res[1][1]["rho_min"] == res[n][1]["rho_min"] and res[1][1]["n_fluid"] == res[n][1]["n_fluid"]


def test_virtual_ranks_reduced_wing(single):
    """Config 4's case file at the size the CPU oracle can afford (3 levels, 2.65 M cells, symmetric half model, WMLES,
    inlet turbulence): 8 virtual ranks with the RCB-yz partition against one context, 6 coarse steps, fast mode."""
    from open_ludwig_b200.host import domain as D
    from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir, have_case
    from open_ludwig_b200.solver import make_params, ramp_velocity
    if not have_case("Wing_5_deg"):
        pytest.skip("case files not shipped (tools/fetch_cases.py)")
    case, ov = CASE_OVERRIDES["wing5_small"]
    dom = D.load_case(case_dir(case), ov, verbose=False, build_tri_map=False)
    params = make_params(dom, strict=False)
    u = ramp_velocity(dom.cfg.u_target, 40, dom.cfg.ramp_steps)
    pr = dom.params
    res = {}
    for n, opts in ((1, None), (8, {"partition": "rcb_yz"}), (5, None)):
        with cabi.MultiContext(n, devices=[0] * n, options=opts) as m:
            for lv in dom.levels:
                m.add_level(lv)
            m.init_equilibrium()
            h = m.create_forces(dom.mesh.centers, dom.mesh.normals, dom.mesh.areas, pr.rho_physical, pr.u_physical, pr.reference_area,
                                pr.reference_chord, pr.moment_center, dom.cfg.symmetric)
            m.step_batch(1, 6, u, params)
            m.sync()
            res[n] = ({f"L{i}{k}": m.download(i, w) for i in range(len(dom.levels)) for k, w in (("f", cabi.F), ("rho", cabi.RHO), ("vel", cabi.VEL))},
                      m.compute_aerodynamics(h, len(dom.levels) - 1, pr.mesh_offset, pr.velocity_scale, pr.rho_physical, 5))
    for n in (8, 5):
        for k in res[1][0]:
            assert np.array_equal(res[1][0][k].view(np.int32), res[n][0][k].view(np.int32)), (n, k)
        for k in ("Cd", "Cl", "Cmy"):
            assert res[n][1][k] == pytest.approx(res[1][1][k], rel=1e-11, abs=1e-16), (n, k)
