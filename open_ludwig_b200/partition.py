"""Host-side mirror of the library's block partition (csrc/abi.cu: morton3, ludwig_partition_starts,
ludwig_level_create): which rank owns which block of a level and which remote blocks a rank's neighbour tables
reference.  Pure NumPy: used by the CPU (gloo) tests and to size halos; the GPU tests check that
``Context.local_blocks`` returns exactly ``local_blocks(...)``.
"""
from __future__ import annotations

import numpy as np


def _spread3(v: np.ndarray) -> np.ndarray:
    x = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_order(coords_1based: np.ndarray) -> np.ndarray:
    """Internal (Morton) order of a level: reference indices (0-based) sorted by the interleaved key."""
    c = np.asarray(coords_1based, np.int64) - 1
    key = _spread3(c[:, 0]) | (_spread3(c[:, 1]) << np.uint64(1)) | (_spread3(c[:, 2]) << np.uint64(2))
    return np.argsort(key, kind="stable").astype(np.int32)


BLOCK_ORDER_DEFAULT = 12   # the library's default "block_order" option (csrc/ludwig_internal.h: opt_block_order)


def xslab_key(coords_1based: np.ndarray, tile: int) -> np.ndarray:
    """Mirror of the library's "block_order = xslab<T>" key (csrc/abi.cu, ludwig_level_create): T x T tiles in (y, z) along a
    Morton curve of the tiles, x-slices one after the other inside a tile."""
    c = np.asarray(coords_1based, np.int64) - 1
    t = np.uint64(tile)
    bx, by, bz = (c[:, i].astype(np.uint64) for i in range(3))
    tile_key = _spread3(by // t) | (_spread3(bz // t) << np.uint64(1))        # same ORDER as the library's 2-D interleave
    return ((tile_key * np.uint64(int(c[:, 0].max()) + 2) + bx) * t + by % t) * t + bz % t


def partition_starts(n_blocks: int, world: int) -> np.ndarray:
    return np.array([(n_blocks * r) // world for r in range(world + 1)], np.int32)


def block_costs(level) -> np.ndarray:
    """Mirror of ludwig_block_costs (csrc/abi.cu): relative cost of every block, reference order, float32."""
    nb = level.n_blocks
    obs = (np.asarray(level.obstacle).reshape(nb, 512) != 0)
    n_obs = obs.sum(axis=1)
    sp = (np.asarray(level.sponge).reshape(nb, 512) > 0).any(axis=1)
    wdv = np.asarray(level.wall_dist).reshape(nb, 512)
    wd = ((wdv > 0) & (wdv < 10)).any(axis=1)
    nt0 = np.asarray(level.neighbor_table) == 0
    miss = nt0.any(axis=0)
    miss_dx0 = nt0[np.arange(27) % 3 == 1].any(axis=0)                 # something missing that is not beyond an x face
    feat = (n_obs > 0) | sp | wd
    c = np.ones(nb, np.float32)
    c[feat] += np.float32(1.0)
    c[n_obs == 512] = np.float32(0.6)
    lean = ~miss_dx0 & ~feat & (level.level_id == 1)                    # feature-less inlet / outlet block: rides in the plain launch
    c[miss & ~lean] += np.float32(1.0)
    if level.bouzidi_enabled and level.cell_block is not None:
        for b in np.asarray(level.cell_block) - 1:          # sequential float32 adds, like the library
            c[b] = np.float32(c[b] + np.float32(0.004))
    return c


def weighted_starts(costs_morton: np.ndarray, world: int) -> np.ndarray:
    """Cut points of `world` contiguous ranges of (approximately) equal cost — ludwig_level_create's rule."""
    n = len(costs_morton)
    pre = np.concatenate([[0.0], np.cumsum(costs_morton.astype(np.float64))])
    st = np.zeros(world + 1, np.int64)
    if world == 1:
        st[1] = n
        return st
    for r in range(1, world):
        target = pre[n] * r / world
        cut = int(np.searchsorted(pre, target, side="left"))
        cut = max(cut, int(st[r - 1]) + 1)
        cut = min(cut, n - (world - r))
        st[r] = cut
    st[world] = n
    return st


def _starts(coords_1based, world, level=None):
    order = morton_order(coords_1based)
    if level is None or world == 1:
        return order, partition_starts(len(order), world)
    return order, weighted_starts(block_costs(level)[order], world)


def internal_order(coords_1based: np.ndarray, world: int, level=None, block_order: int = BLOCK_ORDER_DEFAULT) -> np.ndarray:
    """Reference indices (0-based) in the library's internal order: the owners' ranges are cut on the Morton order, then every
    range is walked in `block_order` (0: Morton, T: x-slab order with T x T tiles)."""
    order, st = _starts(coords_1based, world, level)
    if block_order <= 0:
        return order
    key = xslab_key(coords_1based, block_order)
    out = order.copy()
    for r in range(world):
        seg = order[st[r]:st[r + 1]]
        out[st[r]:st[r + 1]] = seg[np.argsort(key[seg], kind="stable")]
    return out


def local_blocks(coords_1based: np.ndarray, rank: int, world: int, level=None, block_order: int = BLOCK_ORDER_DEFAULT) -> np.ndarray:
    """0-based reference indices of the blocks rank `rank` owns, in internal order.  With `level` the cut is
    cost-weighted exactly as the library does it; without, equal block counts."""
    _, st = _starts(coords_1based, world, level)
    return internal_order(coords_1based, world, level, block_order)[st[rank]:st[rank + 1]]


def owner_of_ref(coords_1based: np.ndarray, world: int, level=None) -> np.ndarray:
    """owner rank of every block, indexed by reference index."""
    order, st = _starts(coords_1based, world, level)
    own = np.empty(len(order), np.int32)
    for r in range(world):
        own[order[st[r]:st[r + 1]]] = r
    return own


def remote_neighbours(neighbor_table: np.ndarray, coords_1based: np.ndarray, rank: int, world: int, level=None) -> np.ndarray:
    """Sorted reference indices (0-based) of the blocks owned by other ranks that `rank` pulls from."""
    own = owner_of_ref(coords_1based, world, level)
    mine = np.nonzero(own == rank)[0]
    nb = neighbor_table[:, mine]            # [27, n_mine], 1-based, 0 = none
    refs = np.unique(nb[nb > 0]) - 1
    return refs[own[refs] != rank]
