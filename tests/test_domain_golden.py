"""Golden integer KATs of the domain build (CPU, no GPU): the host-side restatement of the kept Julia driver's
setup code must reproduce EVERY integer the reference's shipped console logs print.

  RESULTS_SPHERE_RE1M.txt:48,60-106,160-163     (ball1m with surface_resolution 25, velocity 14.8)
  RESULTS_SPHERE_RE10M.txt:48,60-116,181-185    (ball1m as shipped)
  RESULTS_SPHERE_RE266K.txt:37,44,56-94,99,156-158  (ball1m with surface_resolution 25, velocity 4.0)
"""
import numpy as np
import pytest

from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir, have_case

pytestmark = pytest.mark.skipif(not have_case("ball1m"), reason="reference case files not available (run tools/fetch_cases.py)")


@pytest.fixture(scope="module")
def re1m():
    case, ov = CASE_OVERRIDES["sphere_re1m"]
    return D.load_case(case_dir(case), ov)


@pytest.fixture(scope="module")
def re10m():
    case, ov = CASE_OVERRIDES["sphere_re10m"]
    return D.load_case(case_dir(case), ov)


def test_re1m_scaling(re1m):
    p = re1m.params
    assert p.num_levels == 3                                         # RESULTS_SPHERE_RE1M.txt:48
    assert (p.bx_max, p.by_max, p.bz_max) == (8, 7, 7)              # :60
    assert round(p.re_number) == 986667                              # :48
    assert [f"{t:.6f}" for t in p.tau_levels] == ["0.500009", "0.500005", "0.500002"]   # :103
    assert [round(v, 3) for v in p.mesh_offset] == [4.25, 4.48, 4.48]                    # :160
    assert round(re1m.levels[-1].dx, 6) == 0.04                      # :161
    assert round(float(np.float32(p.rho_physical * p.velocity_scale ** 2)), 2) == 298137.78   # :162
    assert re1m.mesh.n_triangles == 20480                            # :34


def test_re1m_topology_counts(re1m):
    r = re1m.reports
    assert [x.n_blocks for x in r] == [392, 1000, 1728]              # :60,:80,:97
    assert [x.halo_added for x in r[1:]] == [988, 1660]              # :72,:84
    assert [x.filled_voxels for x in r] == [28, 548, 6084]           # :63,:75,:87
    assert r[2].n_boundary_cells == 5824                             # :94
    assert re1m.total_cells == 1597440                               # :106 "1.6 M"
    # near-wall counter of the reference is racy (412 / 405 for identical geometry in the two logs): order of magnitude only
    assert r[0].near_wall_cells == 412 and r[1].near_wall_cells == 1160


def test_re10m_counts(re10m):
    p, r = re10m.params, re10m.reports
    assert p.num_levels == 4 and (p.bx_max, p.by_max, p.bz_max) == (8, 8, 8)            # RESULTS_SPHERE_RE10M.txt:48,60
    assert [x.n_blocks for x in r] == [512, 1728, 1856, 3552]                            # :60-109
    assert [x.halo_added for x in r[1:]] == [1664, 1772, 3250]
    assert [x.filled_voxels for x in r] == [44, 778, 8342, 76288]                        # :63,:75,:87,:99
    assert r[3].n_boundary_cells == 28400                                                # :106
    assert [f"{t:.6f}" for t in p.tau_levels] == ["0.500008", "0.500004", "0.500002", "0.500001"]   # :116
    assert [round(v, 3) for v in p.mesh_offset] == [4.25, 4.655, 4.655]                  # :181


def test_re266k_scaling_and_counts():
    """The third shipped log: the Re = 1M grid at U = 4 m/s — same topology, different tau per level and force scales."""
    case, ov = CASE_OVERRIDES["sphere_re266k"]
    dom = D.load_case(case_dir(case), ov)
    p, r = dom.params, dom.reports
    assert p.num_levels == 3 and (p.bx_max, p.by_max, p.bz_max) == (8, 7, 7)            # RESULTS_SPHERE_RE266K.txt:44,56
    assert round(p.re_number) == 266667                                                  # :44
    assert [f"{t:.6f}" for t in p.tau_levels] == ["0.500034", "0.500017", "0.500008"]   # :99
    assert [x.n_blocks for x in r] == [392, 1000, 1728]                                  # :56,:76,:93
    assert [x.halo_added for x in r[1:]] == [988, 1660]                                  # :68,:80
    assert [x.filled_voxels for x in r] == [28, 548, 6084]                               # :59,:71,:83
    assert r[2].n_boundary_cells == 5824                                                 # :90
    assert [round(v, 3) for v in p.mesh_offset] == [4.25, 4.48, 4.48]                    # :156
    assert round(float(np.float32(p.rho_physical * p.velocity_scale ** 2)), 2) == 21777.78   # :158
    assert p.u_physical == 4.0


def test_tables_are_consistent(re1m):
    """Structural invariants of the tables handed to the kernels (1-based, 0 = none, column-major)."""
    for lv in re1m.levels:
        nb = lv.n_blocks
        c = lv.active_block_coords
        assert np.all(np.lexsort((c[:, 2], c[:, 1], c[:, 0])) == np.arange(nb))      # sort(collect(active_set))
        bp = lv.block_pointer                                                         # numpy [bz,by,bx]
        assert bp.shape == (c[:, 2].max(), c[:, 1].max(), c[:, 0].max())             # blocks.jl:104-115
        assert np.array_equal(bp[c[:, 2] - 1, c[:, 1] - 1, c[:, 0] - 1], np.arange(1, nb + 1))
        nt = lv.neighbor_table
        assert nt.shape == (27, nb) and nt.min() >= 0 and nt.max() <= nb
        assert np.array_equal(nt[13], np.arange(1, nb + 1))                          # dir (0,0,0) is the block itself
        for d in range(27):                                                           # symmetry: nbr of nbr in -dir is self
            has = nt[d] > 0
            assert np.array_equal(nt[26 - d][nt[d][has] - 1], np.nonzero(has)[0] + 1)
    fin = re1m.levels[-1]
    assert fin.q_map.dtype == np.float16 and fin.q_map.shape == (27, fin.n_blocks, 8, 8, 8)
    q = fin.q_map.astype(np.float32)
    assert q.min() >= 0 and q.max() <= 1.0 and np.all(q[13] == 0)
    has = (q > 0).any(axis=0)
    assert int(has.sum()) == fin.n_boundary_cells
    b, z, y, x = np.nonzero(has)
    assert np.array_equal(np.stack([b + 1, x + 1, y + 1, z + 1], 1),
                          np.stack([fin.cell_block, fin.cell_x, fin.cell_y, fin.cell_z], 1).astype(np.int64))
    assert fin.tri_map.min() >= 0 and fin.tri_map.max() <= re1m.mesh.n_triangles
    assert np.all((fin.tri_map > 0) == (fin.q_map > 0))
