"""K1 single-level parity: CUDA (through the C ABI) vs the CPU oracle on seeded synthetic boxes.

strict_fp=1 must be BIT-EXACT (no wall model => no powf/logf on the path); strict_fp=0 (FMA + regrouped
sums) must stay within north_star's 1e-5 on rho and u.
"""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
from util import default_params, fetch_state, load_state, rel_err_rho_u

pytestmark = pytest.mark.gpu


def run(lib, lv, state, params, steps, u=0.03):
    with cabi.Context(lib) as c:
        c.add_level(lv)
        load_state(c, 0, *state)
        c.step_batch(1, steps, u, params)
        c.sync()
        return fetch_state(c, 0), c.flow_stats(0)


@pytest.mark.parametrize("dims,periodic", [((4, 3, 3), True), ((3, 2, 2), False), ((6, 6, 6), True)])
def test_strict_bit_exact(oracle_lib, cuda_lib, dims, periodic):
    lv = syn.make_box_level(*dims, periodic_y=periodic, periodic_z=periodic)
    state = syn.noise_state(lv)
    cells = tuple(8 * d for d in dims)
    p = default_params(cells, strict=1)
    ref, sref = run(oracle_lib, lv, state, p, 7)
    got, sgot = run(cuda_lib, lv, state, p, 7)
    for name in ref:
        assert np.array_equal(ref[name].view(np.int32), got[name].view(np.int32)), f"{name} differs"
    assert sref["n_fluid"] == sgot["n_fluid"]
    assert sref["rho_min"] == sgot["rho_min"] and sref["rho_max"] == sgot["rho_max"] and sref["v_max"] == sgot["v_max"]
    assert abs(sref["rho_mean"] - sgot["rho_mean"]) < 1e-12


def test_fast_within_tolerance(oracle_lib, cuda_lib):
    dims = (6, 6, 6)
    lv = syn.make_box_level(*dims)
    state = syn.noise_state(lv)
    cells = tuple(8 * d for d in dims)
    ref, _ = run(oracle_lib, lv, state, default_params(cells, strict=1), 50)
    got, _ = run(cuda_lib, lv, state, default_params(cells, strict=0), 50)
    e_rho, e_u = rel_err_rho_u(ref, got)
    assert e_rho <= 1e-5 and e_u <= 1e-5, (e_rho, e_u)   # tolerance stated by north_star
