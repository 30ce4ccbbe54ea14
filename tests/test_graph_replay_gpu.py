"""CUDA-graph replay of coarse steps (option graphs; abi.cu graph_coarse_step) against eager stepping: identical bits.

The two-level feature case exercises everything a captured coarse step contains — the interface pre-pass on its own stream, the
concurrent K1 launch classes (graph branches), the two Bouzidi phases — and everything that must NOT be baked into the graph:
the noise seed (inlet turbulence on: it changes every sub-step) and the ramped inlet velocity (it changes every batch)."""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
from util import default_params, fetch_state, load_state

import test_k1_features_gpu as T

pytestmark = pytest.mark.gpu


def run(cuda_lib, strict, graphs, batches, change_params_at=None):
    levels = T.build_case()
    cells = tuple(8 * d for d in T.DIMS)
    p = default_params(cells, strict=strict, wall_model_active=1, use_temporal=1, inlet_turbulence=0.02)
    with cabi.Context(cuda_lib, options={"graphs": graphs}) as c:
        for lv in levels:
            c.add_level(lv)
        c.init_uniform_flow(0.01)
        t = 1
        for i, n in enumerate(batches):
            if change_params_at == i:
                p = default_params(cells, strict=strict, wall_model_active=1, use_temporal=1, inlet_turbulence=0.02, c_wale=0.3)
            c.step_batch(t, n, 0.01 + 0.002 * i, p)          # a different inlet velocity every batch, as the ramp does
            t += n
        c.sync()
        return [fetch_state(c, i) for i in range(2)], c.graph_replays(), c.launch_count()


@pytest.mark.parametrize("strict", [0, 1])
def test_graph_replay_is_bit_identical_to_eager(cuda_lib, strict):
    batches = (3, 1, 4, 2, 5)                                  # odd and even batch lengths: both buffer parities start a batch
    eager, r0, n0 = run(cuda_lib, strict, 0, batches)
    graph, r1, n1 = run(cuda_lib, strict, 1, batches)
    assert r0 == 0 and r1 >= sum(batches) - 2                  # all but the two capturing steps were replays
    for a, b in zip(eager, graph):
        for k in a:
            assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), (strict, k)


def test_graph_is_recaptured_when_parameters_change(cuda_lib):
    batches = (4, 4, 4)
    eager, *_ = run(cuda_lib, 0, 0, batches, change_params_at=1)
    graph, replays, _ = run(cuda_lib, 0, 1, batches, change_params_at=1)
    assert replays >= 12 - 4
    for a, b in zip(eager, graph):
        for k in a:
            assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), k


def test_auto_mode_uses_graphs_for_multi_level_cases_only(cuda_lib):
    levels = T.build_case()
    p = default_params(tuple(8 * d for d in T.DIMS), strict=0, use_temporal=1)
    with cabi.Context(cuda_lib) as c:
        for lv in levels:
            c.add_level(lv)
        c.init_equilibrium()
        c.step_batch(1, 6, 0.02, p); c.sync()
        assert c.graph_replays() == 4
    lv = syn.make_box_level(4, 4, 4)
    with cabi.Context(cuda_lib) as c:
        c.add_level(lv)
        load_state(c, 0, *syn.noise_state(lv))
        c.step_batch(1, 6, 0.03, default_params((32, 32, 32), strict=0)); c.sync()
        assert c.graph_replays() == 0


def test_uniform_flow_initial_state_matches_the_oracle(oracle_lib, cuda_lib):
    """ludwig_init_uniform_flow (the strong-scaling record's initial condition) + 6 steps: CUDA strict == oracle bit for bit"""
    levels = T.build_case()
    p = default_params(tuple(8 * d for d in T.DIMS), strict=1, wall_model_active=1, use_temporal=1, inlet_turbulence=0.02)
    out = {}
    for name, lib in (("oracle", oracle_lib), ("cuda", cuda_lib)):
        with cabi.Context(lib) as c:
            for lv in levels:
                c.add_level(lv)
            c.init_uniform_flow(0.02)
            c.step_batch(1, 6, 0.02, p); c.sync()
            out[name] = [fetch_state(c, i) for i in range(2)]
    for a, b in zip(out["oracle"], out["cuda"]):
        for k in a:
            assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), k
