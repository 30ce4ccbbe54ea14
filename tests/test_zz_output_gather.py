"""N3 output gather (ludwig_output_gather = io_vtk.jl:52-58,100-111 for a block list): the arrays the reference's VTK
writer fills — rho_arr, vel_mat (component fastest), obst_arr — for a subset of blocks in a caller-chosen order, against
the same arrays rebuilt with numpy from whole-field downloads (what the reference does).  The CPU oracle pins the
ABI / binding semantics without a GPU; the CUDA library is checked against the same expectation on the GPU box."""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn


def _case():
    lv = syn.make_box_level(3, 2, 2)
    nb = lv.n_blocks
    rng = np.random.default_rng(7)
    lv.obstacle = (rng.random((nb, 8, 8, 8)) < 0.1).astype(np.uint8)
    rho = (1.0 + 0.01 * rng.standard_normal((nb, 8, 8, 8))).astype(np.float32)
    vel = (0.03 * rng.standard_normal((3, nb, 8, 8, 8))).astype(np.float32)
    vel_temp = (0.03 * rng.standard_normal((3, nb, 8, 8, 8))).astype(np.float32)
    rho[1, 2, 3, 4] = np.nan; vel[0, 2, 0, 0, 1] = np.inf; vel_temp[2, 0, 7, 7, 7] = -np.inf    # io_vtk.jl:110-111
    blocks = np.array([5, 0, 11, 7, 3], np.int32)                                              # any order, any subset
    return lv, rho, vel, vel_temp, blocks


def _expected(lv, rho, vel, blocks):
    clean = lambda a: np.where(np.isfinite(a), a, np.float32(0)).astype(np.float32)
    e_rho = clean(rho[blocks].reshape(-1))
    e_vel = clean(vel[:, blocks].reshape(3, -1).T)
    e_obs = (lv.obstacle[blocks].reshape(-1) != 0).astype(np.uint8)
    return e_rho, e_vel, e_obs


def _check(lib):
    lv, rho, vel, vel_temp, blocks = _case()
    with cabi.Context(lib) as c:
        c.add_level(lv)
        c.upload(0, cabi.RHO, rho); c.upload(0, cabi.VEL, vel); c.upload(0, cabi.VEL_TEMP, vel_temp)
        for t_step, v in ((7, vel), (8, vel_temp)):                      # odd -> level.vel, even -> level.vel_temp (:56)
            g_rho, g_vel, g_obs = c.output_gather(0, t_step, blocks)
            e_rho, e_vel, e_obs = _expected(lv, rho, v, blocks)
            assert np.array_equal(g_rho.view(np.int32), e_rho.view(np.int32))
            assert np.array_equal(g_vel.view(np.int32), e_vel.view(np.int32))
            assert np.array_equal(g_obs, e_obs)
        with pytest.raises(cabi.LudwigError):
            c.output_gather(0, 1, np.array([lv.n_blocks], np.int32))     # one past the last block


def test_output_gather_oracle(oracle_lib):
    _check(oracle_lib)


@pytest.mark.gpu
def test_output_gather_cuda(cuda_lib):
    _check(cuda_lib)
