"""Checkpoint / restart (N4): run 2n steps == run n steps, checkpoint, restore into a fresh context, run n more —
bit for bit, single level and two levels (where the children blend the parent's pre-step state)."""
import numpy as np
import pytest

from open_ludwig_b200 import cabi, checkpoint
from open_ludwig_b200.host import synthetic as syn
from util import default_params, fetch_state, load_state
import test_k1_features_gpu as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("strict", [1, 0])
def test_restart_is_bit_exact_two_level(cuda_lib, tmp_path, strict):
    levels = T.build_case()
    cells = tuple(8 * d for d in T.DIMS)
    p = default_params(cells, strict=strict, wall_model_active=1, use_temporal=1, inlet_turbulence=0.02)
    n = 8                                       # a whole number of batches: restart happens between coarse steps
    with cabi.Context(cuda_lib) as c:
        for lv in levels:
            c.add_level(lv)
        c.init_equilibrium()
        c.step_batch(1, 2 * n, 0.02, p)
        c.sync()
        want = [fetch_state(c, i) for i in range(2)]
    path = str(tmp_path / "ckpt")
    with cabi.Context(cuda_lib) as c:
        for lv in levels:
            c.add_level(lv)
        c.init_equilibrium()
        c.step_batch(1, n, 0.02, p)
        checkpoint.save(path, c, n + 1)
    with cabi.Context(cuda_lib) as c:
        for lv in levels:
            c.add_level(lv)
        t = checkpoint.load(path, c)
        assert t == n + 1
        c.step_batch(t, n, 0.02, p)
        c.sync()
        got = [fetch_state(c, i) for i in range(2)]
    for a, b in zip(want, got):
        for k in a:
            assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), k


@pytest.mark.parametrize("strict", [1, 0])
def test_restart_across_rank_counts_and_partitions(cuda_lib, tmp_path, strict):
    """world = 2 virtual ranks (Morton ranges) write per-rank shards; the state is restored into 3 ranks with the RCB-yz
    partition and into ONE context, and both continue to the same bits as an uninterrupted single-context run."""
    levels = T.build_case()
    cells = tuple(8 * d for d in T.DIMS)
    p = default_params(cells, strict=strict, wall_model_active=1, use_temporal=1, inlet_turbulence=0.02)
    n = 6
    with cabi.Context(cuda_lib) as c:
        for lv in levels:
            c.add_level(lv)
        c.init_equilibrium()
        c.step_batch(1, 2 * n, 0.02, p)
        c.sync()
        want = [fetch_state(c, i) for i in range(2)]
    path = str(tmp_path / "ckpt_w2")
    with cabi.MultiContext(2, devices=[0, 0]) as m:
        for lv in levels:
            m.add_level(lv)
        m.init_equilibrium()
        m.step_batch(1, n, 0.02, p)
        checkpoint.save(path, m, n + 1)
    import os
    assert sorted(os.listdir(path)) == ["meta.json", "shard_0.npz", "shard_1.npz"]
    for world, opts in ((3, {"partition": "rcb_yz"}), (1, None)):
        with cabi.MultiContext(world, devices=[0] * world, options=opts) as m:
            for lv in levels:
                m.add_level(lv)
            t = checkpoint.load(path, m)
            assert t == n + 1
            m.step_batch(t, n, 0.02, p)
            m.sync()
            got = [{k: m.download(i, w) for k, w in (("f", cabi.F), ("f_temp", cabi.F_TEMP), ("rho", cabi.RHO), ("vel", cabi.VEL), ("vel_temp", cabi.VEL_TEMP))}
                   for i in range(2)]
        for a, b in zip(want, got):
            for k in a:
                assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), (world, k)
    os.remove(os.path.join(path, "shard_1.npz"))
    with cabi.Context(cuda_lib) as c:
        for lv in levels:
            c.add_level(lv)
        with pytest.raises(ValueError):
            checkpoint.load(path, c)
