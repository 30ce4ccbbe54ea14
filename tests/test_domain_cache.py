"""The domain cache used by multi-rank launches (rank 0 builds, the others map the arrays): a saved and re-loaded
Domain must hand the library exactly the same tables."""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir, have_case


@pytest.mark.skipif(not have_case("ball1m"), reason="case folder not available")
def test_save_load_round_trip(tmp_path):
    case, ov = CASE_OVERRIDES["sphere_re1m"]
    dom = D.load_case(case_dir(case), ov)
    D.save_domain(dom, str(tmp_path))
    back = D.load_domain(str(tmp_path))
    assert len(back.levels) == len(dom.levels) and back.total_cells == dom.total_cells
    assert back.cell_updates_per_coarse_step == dom.cell_updates_per_coarse_step
    for a, b in zip(dom.levels, back.levels):
        for k, v in vars(a).items():
            w = getattr(b, k)
            if isinstance(v, np.ndarray):
                assert w.dtype == v.dtype and w.shape == v.shape and np.array_equal(v, w), k
            else:
                assert v == w, k
        # the descriptor the ABI receives is built from memory-mapped, read-only arrays without copies
        d, keep = cabi.Context.make_desc(b)
        assert d.n_blocks == a.n_blocks and d.n_boundary_cells == (a.n_boundary_cells if a.bouzidi_enabled else 0)
    assert np.array_equal(back.mesh.centers, dom.mesh.centers) and back.params.mesh_offset == dom.params.mesh_offset
