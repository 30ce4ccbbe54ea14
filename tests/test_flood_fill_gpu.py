"""N2, flood fill on the device (csrc/domain_gpu.cu: ludwig_domain_flood_fill, a frontier of blocks) against the host restatement of
domain_generation.jl:114-203 (host/domain_build.cpp: ludwig_host_flood_fill, which reproduces the golden "Filled N interior voxels"
counts of the reference's logs, tests/test_domain_golden.py) on random mazes: dense random obstacle fields in which corridors wind
through and across blocks, with holes in the block set, enclosed cavities and unreachable blocks.  Arrays byte-identical."""
import ctypes as C

import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host.domain import host_lib

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def both(coords, dims, obstacle):
    nb = coords.shape[0]
    grid = np.zeros(dims, np.int32)                                  # [bx][by][bz], 1-based block index, 0 = none
    grid[coords[:, 0] - 1, coords[:, 1] - 1, coords[:, 2] - 1] = np.arange(1, nb + 1)
    grid_cm = np.ascontiguousarray(grid.transpose(2, 1, 0))          # Julia's column-major block_pointer bytes
    host = obstacle.copy(); dev = obstacle.copy()
    hl = host_lib()
    hl.ludwig_host_flood_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    hl.ludwig_host_flood_fill.restype = C.c_int64
    nh = hl.ludwig_host_flood_fill(_p(host), _p(coords), nb, _p(grid_cm), *dims)
    gl = cabi.load_library()
    nd = gl.ludwig_domain_flood_fill(0, _p(coords), nb, _p(grid), dims[0], dims[1], dims[2], _p(dev))
    assert nd >= 0, gl.ludwig_domain_last_error().decode()
    return nh, host, nd, dev


@pytest.mark.parametrize("seed,density,drop", [(1, 0.25, 0.0), (2, 0.40, 0.1), (3, 0.55, 0.2), (4, 0.62, 0.0), (5, 0.70, 0.05), (6, 0.0, 0.3)])
def test_random_mazes(seed, density, drop):
    rng = np.random.default_rng(seed)
    dims = (7, 5, 6)
    allc = np.array([(x, y, z) for x in range(1, dims[0] + 1) for y in range(1, dims[1] + 1) for z in range(1, dims[2] + 1)], np.int32)
    keep = rng.random(len(allc)) >= drop
    keep[allc[:, 0] == 1] |= rng.random(int((allc[:, 0] == 1).sum())) < 0.7          # some blocks on the seed plane
    coords = np.ascontiguousarray(allc[keep])
    nb = len(coords)
    obstacle = (rng.random((nb, 8, 8, 8)) < density).astype(np.uint8)
    free_blocks = rng.random(nb) < 0.3                                               # obstacle-free blocks take the one-node path
    obstacle[free_blocks] = 0
    walls = rng.random(nb) < 0.1                                                     # fully solid blocks cut regions off
    obstacle[walls] = 1
    nh, host, nd, dev = both(coords, dims, obstacle)
    assert nh == nd and host.tobytes() == dev.tobytes()
    assert 0 < nh < obstacle.size or density == 0.0


def test_corridor_longer_than_one_pass():
    """A single one-cell-wide corridor snaking through a row of blocks and back: the frontier has to revisit blocks."""
    dims = (6, 1, 1)
    coords = np.array([(x, 1, 1) for x in range(1, 7)], np.int32)
    obstacle = np.ones((6, 8, 8, 8), np.uint8)
    obstacle[:, 0, 0, :] = 0            # z = 0, y = 0 : corridor along x through every block
    obstacle[5, 0, :, 7] = 0            # turn in the last block (x = 7 column, y = 0..7)
    obstacle[:, 0, 7, :] = 0            # and all the way back along y = 7
    obstacle[0, 3, 7, 0] = 0; obstacle[0, 2, 7, 0] = 0; obstacle[0, 1, 7, 0] = 0   # a side pocket in the first block, reached last
    obstacle[2, 5, 5, 5] = 0            # an enclosed cavity: filled
    nh, host, nd, dev = both(coords, dims, obstacle)
    assert nh == nd == 1 and host.tobytes() == dev.tobytes()
    assert dev[2, 5, 5, 5] == 1 and dev[0, 3, 7, 0] == 0
