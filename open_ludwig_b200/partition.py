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


def partition_starts(n_blocks: int, world: int) -> np.ndarray:
    return np.array([(n_blocks * r) // world for r in range(world + 1)], np.int32)


def local_blocks(coords_1based: np.ndarray, rank: int, world: int) -> np.ndarray:
    """0-based reference indices of the blocks rank `rank` owns, in internal order."""
    order = morton_order(coords_1based)
    st = partition_starts(len(order), world)
    return order[st[rank]:st[rank + 1]]


def owner_of_ref(coords_1based: np.ndarray, world: int) -> np.ndarray:
    """owner rank of every block, indexed by reference index."""
    order = morton_order(coords_1based)
    st = partition_starts(len(order), world)
    own = np.empty(len(order), np.int32)
    for r in range(world):
        own[order[st[r]:st[r + 1]]] = r
    return own


def remote_neighbours(neighbor_table: np.ndarray, coords_1based: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Sorted reference indices (0-based) of the blocks owned by other ranks that `rank` pulls from."""
    own = owner_of_ref(coords_1based, world)
    mine = np.nonzero(own == rank)[0]
    nb = neighbor_table[:, mine]            # [27, n_mine], 1-based, 0 = none
    refs = np.unique(nb[nb > 0]) - 1
    return refs[own[refs] != rank]
