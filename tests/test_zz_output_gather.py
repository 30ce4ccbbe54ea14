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


# ---- the whole export (io_vtk.jl:17-111): valid-block rule + fields of every valid block -------------------------------------

def _two_level_case():
    import test_k1_features_gpu as T
    levels = T.build_case()
    rng = np.random.default_rng(11)
    state = []
    for lv in levels:
        nb = lv.n_blocks
        state.append(((1.0 + 0.01 * rng.standard_normal((nb, 8, 8, 8))).astype(np.float32),
                      (0.03 * rng.standard_normal((3, nb, 8, 8, 8))).astype(np.float32),
                      (0.03 * rng.standard_normal((3, nb, 8, 8, 8))).astype(np.float32)))
    state[1][0][3, 1, 1, 1] = np.nan
    return levels, state


def _expected_export(levels, state, t_step):
    """numpy restatement of export_merged_mesh_sync's block selection and field gathering"""
    sets = [set(map(tuple, np.asarray(lv.active_block_coords).tolist())) for lv in levels]
    rho_l, vel_l, obs_l, lvl_l, valid = [], [], [], [], []
    clean = lambda a: np.where(np.isfinite(a), a, np.float32(0)).astype(np.float32)
    for l, lv in enumerate(levels):
        keep = []
        for b, (bx, by, bz) in enumerate(np.asarray(lv.active_block_coords).tolist()):
            covered = l + 1 < len(levels) and all((2 * bx - 1 + dx, 2 * by - 1 + dy, 2 * bz - 1 + dz) in sets[l + 1]
                                                    for dz in (0, 1) for dy in (0, 1) for dx in (0, 1))
            if not covered:
                keep.append(b)
        keep = np.array(keep, np.int64)
        valid.append(keep)
        rho, vel, vel_temp = state[l]
        v = vel_temp if t_step % 2 == 0 else vel
        rho_l.append(clean(rho[keep].reshape(-1))); vel_l.append(clean(v[:, keep].reshape(3, -1).T))
        obs_l.append((np.asarray(lv.obstacle)[keep].reshape(-1) != 0).astype(np.uint8)); lvl_l.append(np.full(512 * len(keep), lv.level_id, np.int32))
    return valid, np.concatenate(rho_l), np.concatenate(vel_l), np.concatenate(obs_l), np.concatenate(lvl_l)


def _check_export(make_ctx):
    levels, state = _two_level_case()
    with make_ctx() as c:
        for lv in levels:
            c.add_level(lv)
        for l, (rho, vel, vel_temp) in enumerate(state):
            c.upload(l, cabi.RHO, rho); c.upload(l, cabi.VEL, vel); c.upload(l, cabi.VEL_TEMP, vel_temp)
        got_valid = c.output_valid_blocks()
        for t_step in (3, 4):
            valid, e_rho, e_vel, e_obs, e_lvl = _expected_export(levels, state, t_step)
            assert len(valid[0]) < levels[0].n_blocks and len(valid[1]) == levels[1].n_blocks      # some coarse blocks are covered
            for a, b in zip(valid, got_valid):
                assert np.array_equal(a, b)
            rho, vel, obs, lvl = c.output_export(t_step)
            assert np.array_equal(rho.view(np.int32), e_rho.view(np.int32)) and np.array_equal(vel.view(np.int32), e_vel.view(np.int32))
            assert np.array_equal(obs, e_obs) and np.array_equal(lvl, e_lvl)


def test_output_export_oracle(oracle_lib):
    _check_export(lambda: cabi.Context(oracle_lib))


@pytest.mark.gpu
def test_output_export_cuda(cuda_lib):
    _check_export(lambda: cabi.Context(cuda_lib))


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks", [1, 3])
def test_output_export_virtual_ranks(cuda_lib, n_ranks):
    """every rank writes the cells of its own valid blocks into the shared arrays: together, the single-context result"""
    _check_export(lambda: cabi.MultiContext(n_ranks, devices=[0] * n_ranks))


@pytest.mark.gpu
def test_output_gather_more_blocks_than_one_staging_chunk(cuda_lib):
    """5000 listed blocks > the 2048-block staging chunk: the double-buffered pipeline wraps around both buffers"""
    lv = syn.make_box_level(20, 16, 16)
    rng = np.random.default_rng(5)
    rho = rng.random((lv.n_blocks, 8, 8, 8), dtype=np.float32)
    vel = rng.random((3, lv.n_blocks, 8, 8, 8), dtype=np.float32)
    blocks = rng.permutation(lv.n_blocks)[:5000].astype(np.int32)
    with cabi.Context(cuda_lib) as c:
        c.add_level(lv)
        c.upload(0, cabi.RHO, rho); c.upload(0, cabi.VEL, vel)
        g_rho, g_vel, g_obs = c.output_gather(0, 1, blocks)
    assert np.array_equal(g_rho, rho[blocks].reshape(-1)) and np.array_equal(g_vel, vel[:, blocks].reshape(3, -1).T) and not g_obs.any()
