"""Launch-structure options only change HOW the K1 classes are launched (which blocks share a grid, how many CTAs, which register
budget): every field must stay identical word for word to the default configuration, in both FP modes, on a box whose inlet plane
holds x-only face blocks (merged into the plain launch), blocks on y / z faces and corners (general domain-face class), and on the
two-level feature case (obstacle, Bouzidi, sponge, wall model, interface pre-pass: no merge there)."""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
import test_k1_features_gpu as F
from util import default_params, load_state, fetch_state

pytestmark = pytest.mark.gpu
VARIANTS = [{"merge_face": 0}, {"face_persist": 2, "merge_face": 0}, {"strict_loop": 4}, {"strict_loop": 2, "merge_face": 0}, {"strict_feature_occupancy": 3},
            {"strict_feature_occupancy": 5, "merge_face": 0}, {"fork_max_blocks": 0}, {"single_stream": 1}, {"cta_threads": 256}, {"l2_fetch": 32}]


def walled_box(lib, opts, strict, nb=(7, 5, 4), steps=5):
    lv = syn.make_box_level(*nb, periodic_y=False, periodic_z=False)
    f, rho, vel = syn.noise_state(lv)
    p = default_params(tuple(8 * d for d in nb), strict=strict)
    with cabi.Context(lib, options=opts) as c:
        c.add_level(lv); load_state(c, 0, f, rho, vel)
        c.step_batch(1, steps, 0.03, p); c.sync()
        return fetch_state(c, 0)


@pytest.mark.parametrize("strict", [1, 0])
def test_fields_do_not_depend_on_the_launch_structure(cuda_lib, strict):
    levels = F.build_case()
    ref2 = F.run(cuda_lib, levels, 6, strict, True)[0]
    refb = walled_box(cuda_lib, None, strict)
    for opts in VARIANTS:
        got2 = F.run(cuda_lib, levels, 6, strict, True, options=opts)[0]
        gotb = walled_box(cuda_lib, opts, strict)
        for lvl in ref2:
            for name in ref2[lvl]:
                assert np.array_equal(ref2[lvl][name].view(np.int32), got2[lvl][name].view(np.int32)), (opts, lvl, name)
        for name in refb:
            assert np.array_equal(refb[name].view(np.int32), gotb[name].view(np.int32)), (opts, name)


def test_walled_box_strict_equals_oracle(cuda_lib, oracle_lib):
    """The merged launch's x-only body and the general face class against the CPU oracle (bit-identical)."""
    ref = walled_box(oracle_lib, None, 1)
    got = walled_box(cuda_lib, None, 1)
    for name in ref:
        assert np.array_equal(ref[name].view(np.int32), got[name].view(np.int32)), name
