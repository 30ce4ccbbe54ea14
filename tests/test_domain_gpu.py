"""N2: the domain build's three brute-force phases on the device (csrc/domain_gpu.cu: SAT voxelisation, wall distance, q-map ray
casting in Float64, reference operation order) against the host build (host/domain_build.cpp, which reproduces every golden integer
of the reference's logs, tests/test_domain_golden.py): every table byte-identical, and the golden integers reproduced from the
device path too."""
import time

import numpy as np
import pytest

from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir, have_case

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_case("ball1m"), reason="reference case files not available")]
ARRAYS = ("obstacle", "wall_dist", "sponge", "q_map", "cell_block", "cell_x", "cell_y", "cell_z", "neighbor_table", "block_pointer", "tri_map")


def build_both(name, build_tri_map=True):
    case, ov = CASE_OVERRIDES[name]
    t0 = time.time(); host = D.load_case(case_dir(case), ov, build_tri_map=build_tri_map); t1 = time.time()
    dev = D.load_case(case_dir(case), ov, build_tri_map=build_tri_map, gpu_device=0); t2 = time.time()
    print(f"\n{name}: host build {t1 - t0:.2f} s, device-assisted build {t2 - t1:.2f} s ({host.total_cells / 1e6:.2f} M cells)")
    return host, dev


def assert_identical(host, dev):
    assert len(host.levels) == len(dev.levels)
    for a, b in zip(host.levels, dev.levels):
        assert a.n_boundary_cells == b.n_boundary_cells and a.bouzidi_enabled == b.bouzidi_enabled
        for k in ARRAYS:
            x, y = getattr(a, k), getattr(b, k)
            assert (x is None) == (y is None), k
            if x is not None:
                assert x.dtype == y.dtype and x.shape == y.shape and x.tobytes() == y.tobytes(), (a.level_id, k)
    for ra, rb in zip(host.reports, dev.reports):
        assert ra.__dict__ == rb.__dict__


def test_sphere_re1m_tables_identical_and_golden_integers():
    host, dev = build_both("sphere_re1m")
    assert_identical(host, dev)
    # RESULTS_SPHERE_RE1M.txt:60-106,160-163 reproduced from the device path
    assert [r.n_blocks for r in dev.reports] == [392, 1000, 1728]
    assert [r.filled_voxels for r in dev.reports] == [28, 548, 6084]
    assert dev.reports[-1].n_boundary_cells == 5824


def test_sphere_re10m_tables_identical_and_golden_integers():
    host, dev = build_both("sphere_re10m")
    assert_identical(host, dev)
    assert [r.n_blocks for r in dev.reports] == [512, 1728, 1856, 3552]          # RESULTS_SPHERE_RE10M.txt:60-116
    assert [r.filled_voxels for r in dev.reports] == [44, 778, 8342, 76288]
    assert dev.reports[-1].n_boundary_cells == 28400                              # :181


def test_bunny_tables_identical():
    if not have_case("Stanford_bunny"):
        pytest.skip("Stanford_bunny case files not available")
    host, dev = build_both("bunny_small")
    assert_identical(host, dev)
    assert dev.levels[-1].n_boundary_cells == 25825


def test_wing_tables_identical():
    if not have_case("Wing_5_deg"):
        pytest.skip("Wing_5_deg case files not available")
    host, dev = build_both("wing5_small", build_tri_map=False)
    assert_identical(host, dev)
