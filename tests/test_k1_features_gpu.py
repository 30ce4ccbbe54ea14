"""Differential tests of every K1/K2/K3/K4 branch on a synthetic two-level case (no STL needed):
domain faces (inlet, outlet, y/z mirror), outlet sponge with population blending, a solid sphere on the fine
level (full-way bounce-back), wall-model forcing, Bouzidi links, 2:1 interface interpolation with temporal blend,
surface forces and flow statistics.  CUDA through the C ABI vs the CPU oracle on the same seeded inputs.
"""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
from util import default_params, fetch_state, max_ulp_diff, rel_err_rho_u

pytestmark = pytest.mark.gpu

DIMS = (6, 4, 4)                      # level-1 blocks -> 48 x 32 x 32 cells


def build_case(wall_model=True, bouzidi=True):
    l1 = syn.make_box_level(*DIMS, tau=0.5006, periodic_y=False, periodic_z=False, temporal_storage=True)
    syn.add_outlet_sponge(l1, DIMS[0] * 8)
    l2 = syn.sub_level(l1, (2, 2, 2), (4, 3, 3), tau=0.5003, level_id=2)
    # sphere in level-2 global cell coordinates (level-2 cells are half size): centre of the refined region
    syn.add_sphere_obstacle(l2, centre=(2 * 8 * 2.5, 2 * 8 * 2.0, 2 * 8 * 2.0), radius=7.3, with_bouzidi=bouzidi)
    return [l1, l2]


def sphere_mesh(n=400, seed=3, centre=(20.0, 16.0, 16.0), radius=3.65):
    """Random surface patches of the sphere in level-1 lattice units (dx_1 = 1, dx_2 = 0.5)."""
    rng = np.random.default_rng(seed)
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    centers = np.asarray(centre) + radius * nrm
    areas = np.full(n, 4 * np.pi * radius ** 2 / n)
    return centers, nrm, areas


def run(lib, levels, steps, strict, wall_model=True, fine_grained=False, options=None, **overrides):
    cells = tuple(8 * d for d in DIMS)
    kw = dict(strict=strict, wall_model_active=int(wall_model), use_temporal=1, inlet_turbulence=0.02)
    kw.update(overrides)
    p = default_params(cells, **kw)
    with cabi.Context(lib, options=options) as c:
        for lv in levels:
            c.add_level(lv)
        c.init_equilibrium()
        centers, nrm, areas = sphere_mesh()
        mesh = c.create_mesh(centers, nrm, areas)
        forces = c.create_forces(mesh, 1.225, 10.0, 1.0, 1.0, (20.0, 16.0, 16.0), False)
        if fine_grained:
            # the kept Julia driver's own recursion (solver_control.jl:21-143) over the fine-grained entry points
            for t in range(1, steps + 1):
                c.snapshot_old(0, t)
                c.level_step(0, t, 0, 0.0, 0.02, p)
                c.level_step(1, 2 * t, t, 0.0, 0.02, p)
                c.level_step(1, 2 * t + 1, t, 0.5, 0.02, p)
        else:
            c.step_batch(1, steps, 0.02, p)
        c.sync()
        out = {f"L{i}": fetch_state(c, i) for i in range(len(levels))}
        aero = c.compute_aerodynamics(forces, len(levels) - 1, (0.0, 0.0, 0.0), 300.0, 1.225, 5)
        maps = c.download_force_maps(forces, len(areas))
        stats = [c.flow_stats(i) for i in range(len(levels))]
    return out, aero, maps, stats


@pytest.mark.parametrize("wall_model", [False, True])
def test_strict_two_level(oracle_lib, cuda_lib, wall_model):
    levels = build_case()
    assert levels[1].n_boundary_cells > 100
    ref, aref, mref, sref = run(oracle_lib, levels, 12, 1, wall_model)
    got, agot, mgot, sgot = run(cuda_lib, levels, 12, 1, wall_model)
    for lvl in ref:
        for name in ref[lvl]:
            if wall_model:
                # powf/logf differ by <= 2 ulp between glibc and CUDA: not bit-exact, but within a few ulp of f
                assert np.allclose(ref[lvl][name], got[lvl][name], rtol=0, atol=2e-6), (lvl, name)
            else:
                assert np.array_equal(ref[lvl][name].view(np.int32), got[lvl][name].view(np.int32)), (lvl, name)
    if not wall_model:
        for a, b in zip(mref, mgot):
            assert np.array_equal(a.view(np.int32), b.view(np.int32))          # K3 bit-exact
        for k in ("Fx", "Fy", "Fz", "Mx", "My", "Mz", "Cd", "Cl"):
            assert agot[k] == pytest.approx(aref[k], rel=2e-4, abs=1e-9), k    # K4: FP64 tree vs FP32 sequential sum
        for a, b in zip(sref, sgot):
            assert a["n_fluid"] == b["n_fluid"] and a["rho_min"] == b["rho_min"] and a["rho_max"] == b["rho_max"]


def test_fine_grained_api_equals_batch(cuda_lib):
    """ludwig_level_step + ludwig_level_snapshot_old (explicit copy_to_old!) == ludwig_step_batch (copy-free)."""
    levels = build_case()
    a, *_ = run(cuda_lib, levels, 6, 1, True, fine_grained=False)
    b, *_ = run(cuda_lib, levels, 6, 1, True, fine_grained=True)
    for lvl in a:
        for name in a[lvl]:
            assert np.array_equal(a[lvl][name].view(np.int32), b[lvl][name].view(np.int32)), (lvl, name)


def test_fast_two_level(oracle_lib, cuda_lib):
    levels = build_case()
    ref, aref, _, _ = run(oracle_lib, levels, 40, 1, True)
    got, agot, _, _ = run(cuda_lib, levels, 40, 0, True)
    for lvl in ref:
        e_rho, e_u = rel_err_rho_u(ref[lvl], got[lvl])
        assert e_rho <= 1e-5 and e_u <= 5e-5, (lvl, e_rho, e_u)
    assert agot["Cd"] == pytest.approx(aref["Cd"], rel=1e-3)     # north_star: Cd within 0.1 %


@pytest.mark.parametrize("overrides", [dict(use_temporal=0), dict(sponge_blend=0), dict(inlet_turbulence=0.0, symmetric=1),
                                       dict(q_min_threshold=0.3), dict(c_wale=0.2, nu_sgs_bg=0.0)])
def test_parameter_variants(oracle_lib, cuda_lib, overrides):
    """Every scalar argument of perform_timestep_v2! that switches a code path: temporal blending off (new parent
    state only), sponge without population blending, no inlet noise + symmetric flag, a high Bouzidi q threshold
    (different active-link set), WALE without the background viscosity floor."""
    levels = build_case()
    ref, *_ = run(oracle_lib, levels, 8, 1, False, **overrides)
    got, *_ = run(cuda_lib, levels, 8, 1, False, **overrides)
    for lvl in ref:
        for name in ref[lvl]:
            assert np.array_equal(ref[lvl][name].view(np.int32), got[lvl][name].view(np.int32)), (overrides, lvl, name)
    fast, *_ = run(cuda_lib, levels, 8, 0, False, **overrides)
    for lvl in ref:
        # after 8 steps the fine level is still at rest (max|u| ~ 3e-5), so the error is bounded in absolute terms:
        # a few ulp of a population, i.e. FP32 round-off of the regrouped sums
        e_rho, _ = rel_err_rho_u(ref[lvl], fast[lvl])
        assert e_rho <= 1e-5, (overrides, lvl, e_rho)
        assert float(np.abs(ref[lvl]["vel"] - fast[lvl]["vel"]).max()) <= 1e-6, (overrides, lvl)
        assert float(np.abs(ref[lvl]["f"] - fast[lvl]["f"]).max()) <= 2e-6, (overrides, lvl)


def test_block_prepass_variant_is_bit_identical(cuda_lib):
    """Option prepass = block (one CTA per ghost block, parent cells staged in shared memory) performs the same arithmetic in
    the same order as the default one-thread-per-group interface pre-pass: identical bits after 10 two-level steps."""
    levels = build_case()
    a, *_ = run(cuda_lib, levels, 10, 0, True)
    b, *_ = run(cuda_lib, levels, 10, 0, True, options={"prepass": "block"})
    for lvl in a:
        for name in a[lvl]:
            assert np.array_equal(a[lvl][name].view(np.int32), b[lvl][name].view(np.int32)), (lvl, name)


@pytest.mark.parametrize("wall_model", [False, True])
def test_strict_packed_equals_generic_cross_check(cuda_lib, wall_model):
    """The shipped strict kernels (k1_strict.cu: packed FP32x2, ghost blocks + pre-pass) against the one-thread-per-cell
    kernel that keeps every branch of the reference inside it (option strict_generic): identical bits, wall model included
    (both call libdevice powf / logf)."""
    levels = build_case()
    a, *_ = run(cuda_lib, levels, 12, 1, wall_model)
    b, *_ = run(cuda_lib, levels, 12, 1, wall_model, options={"strict_generic": 1})
    for lvl in a:
        for name in a[lvl]:
            assert np.array_equal(a[lvl][name].view(np.int32), b[lvl][name].view(np.int32)), (lvl, name)


def test_options_are_validated(cuda_lib):
    with cabi.Context(cuda_lib) as c:
        with pytest.raises(cabi.LudwigError):
            c.set_option("no_such_option", 1)
        with pytest.raises(cabi.LudwigError):
            c.set_option("partition", "hilbert")
        c.set_option("partition", "rcb_yz")
        c.add_level(build_case()[0])
        with pytest.raises(cabi.LudwigError):     # partition rule is fixed once a level exists
            c.set_option("partition", "morton")



@pytest.mark.parametrize("variant", ["stash", "tma"])
@pytest.mark.parametrize("wall_model", [False, True])
def test_strict_kernel_variants_are_bit_identical(cuda_lib, variant, wall_model):
    """The three forms of the strict K1 — pulled populations in registers (default), in a shared-memory stash (3 CTAs / SM),
    persistent CTAs with cp.async.bulk (TMA) staged double-buffered block tiles — run the same strict_block body: same bits on
    the two-level case with every feature (all four launch classes, ghost blocks, domain faces, obstacles, sponge, wall model)."""
    levels = build_case()
    a, *_ = run(cuda_lib, levels, 12, 1, wall_model)
    b, *_ = run(cuda_lib, levels, 12, 1, wall_model, options={"strict_kernel": variant})
    for lvl in a:
        for name in a[lvl]:
            assert np.array_equal(a[lvl][name].view(np.int32), b[lvl][name].view(np.int32)), (variant, lvl, name)


def test_fast_tma_kernel_variant_is_bit_identical(cuda_lib):
    """fast_kernel = tma (persistent CTAs, cp.async.bulk staged tiles) runs the same fast_block body as the direct-load kernel."""
    levels = build_case()
    a, *_ = run(cuda_lib, levels, 12, 0, True)
    b, *_ = run(cuda_lib, levels, 12, 0, True, options={"fast_kernel": "tma"})
    for lvl in a:
        for name in a[lvl]:
            assert np.array_equal(a[lvl][name].view(np.int32), b[lvl][name].view(np.int32)), (lvl, name)


@pytest.mark.parametrize("strict", [0, 1])
@pytest.mark.parametrize("threads", [256, 128, 64])
def test_cta_threads_option_is_bit_identical(cuda_lib, strict, threads):
    """cta_threads = 128 / 64: a CTA takes 4 / 2 z-planes of a block instead of all 8 — same per-cell code, same bits."""
    levels = build_case()
    a, *_ = run(cuda_lib, levels, 10, strict, True)                                   # auto: 64 strict / 128 fast
    b, *_ = run(cuda_lib, levels, 10, strict, True, options={"cta_threads": threads})
    for lvl in a:
        for name in a[lvl]:
            assert np.array_equal(a[lvl][name].view(np.int32), b[lvl][name].view(np.int32)), (threads, lvl, name)


@pytest.mark.parametrize("occ", [4, 6])
def test_strict_occupancy_option_is_bit_identical(cuda_lib, occ):
    levels = build_case()
    a, *_ = run(cuda_lib, levels, 10, 1, True)
    b, *_ = run(cuda_lib, levels, 10, 1, True, options={"strict_occupancy": occ})
    for lvl in a:
        for name in a[lvl]:
            assert np.array_equal(a[lvl][name].view(np.int32), b[lvl][name].view(np.int32)), (occ, lvl, name)
