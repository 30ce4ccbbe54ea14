"""The internal block order ("block_order" option: Morton curve, or x-slab order with T x T tiles in (y, z), the default) only
changes WHERE a block lives in HBM and when it runs: every field must be identical word for word, in both FP modes, on the
single-level noise box and on the two-level feature case (obstacle, Bouzidi, sponge, wall model, interface); and the library's
order must be the one the host mirror (open_ludwig_b200/partition.py) derives."""
import numpy as np
import pytest

from open_ludwig_b200 import cabi, partition
from open_ludwig_b200.host import synthetic as syn
import test_k1_features_gpu as F
from util import default_params, load_state, fetch_state

pytestmark = pytest.mark.gpu
ORDERS = ["morton", "xslab2", "xslab12", "xslab5"]


def box(lib, order, strict, nb=(12, 10, 9), steps=5):
    lv = syn.make_box_level(*nb)
    f, rho, vel = syn.noise_state(lv)
    p = default_params(tuple(8 * d for d in nb), strict=strict)
    with cabi.Context(lib, options={"block_order": order}) as c:
        c.add_level(lv); load_state(c, 0, f, rho, vel)
        loc = c.local_blocks(0)
        c.step_batch(1, steps, 0.03, p); c.sync()
        return fetch_state(c, 0), loc, lv


@pytest.mark.parametrize("strict", [1, 0])
def test_fields_do_not_depend_on_the_block_order(cuda_lib, strict):
    levels = F.build_case()
    ref2 = F.run(cuda_lib, levels, 8, strict, True, options={"block_order": ORDERS[0]})[0]
    refb, _, _ = box(cuda_lib, ORDERS[0], strict)
    for order in ORDERS[1:]:
        got2 = F.run(cuda_lib, levels, 8, strict, True, options={"block_order": order})[0]
        gotb, _, _ = box(cuda_lib, order, strict)
        for lvl in ref2:
            for name in ref2[lvl]:
                assert np.array_equal(ref2[lvl][name].view(np.int32), got2[lvl][name].view(np.int32)), (order, lvl, name)
        for name in refb:
            assert np.array_equal(refb[name].view(np.int32), gotb[name].view(np.int32)), (order, name)


def test_library_order_equals_host_mirror(cuda_lib):
    for order, t in (("morton", 0), ("xslab12", 12), ("xslab5", 5)):
        _, loc, lv = box(cuda_lib, order, 1, steps=1)
        assert np.array_equal(loc, partition.local_blocks(lv.active_block_coords, 0, 1, block_order=t)), order
    with cabi.Context(cuda_lib) as c:      # the library default is the mirror's default
        lv = syn.make_box_level(5, 14, 13)
        c.add_level(lv)
        assert np.array_equal(c.local_blocks(0), partition.local_blocks(lv.active_block_coords, 0, 1))


def test_block_order_is_validated(cuda_lib):
    with cabi.Context(cuda_lib) as c:
        with pytest.raises(cabi.LudwigError):
            c.set_option("block_order", "xslab1")
        with pytest.raises(cabi.LudwigError):
            c.set_option("block_order", "hilbert")
        c.add_level(syn.make_box_level(2, 2, 2))
        with pytest.raises(cabi.LudwigError):      # the order is baked into the level tables
            c.set_option("block_order", "morton")
