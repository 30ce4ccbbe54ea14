"""Synthetic cases built without an STL (SURVEY.md §8(d) "Config 2" and the differential-test inputs).

Everything here produces host arrays in the reference layout (see ``cabi.BlockLevel``), using the
reference's own conventions: blocks sorted lexicographically by (bx,by,bz) as ``sort(collect(active_set))``
does (domain.jl:171), 1-based tables, ``dir = (dx+1) + (dy+1)*3 + (dz+1)*9`` (domain_topology.jl:148).
"""
from __future__ import annotations

import numpy as np

from ..cabi import BlockLevel

CX = np.array([k % 3 - 1 for k in range(27)], dtype=np.int32)
CY = np.array([(k // 3) % 3 - 1 for k in range(27)], dtype=np.int32)
CZ = np.array([k // 9 - 1 for k in range(27)], dtype=np.int32)
_D2 = CX * CX + CY * CY + CZ * CZ
# physics_v2.jl:106: Float32 literals 8f0/27f0 ... evaluated in Float32
W = np.where(_D2 == 0, np.float32(8) / np.float32(27),
             np.where(_D2 == 1, np.float32(2) / np.float32(27),
                      np.where(_D2 == 2, np.float32(1) / np.float32(54), np.float32(1) / np.float32(216)))).astype(np.float32)


def gpu_hash(x: np.ndarray) -> np.ndarray:
    """physics_utils.jl:17-22 on uint32 arrays."""
    h = x.astype(np.uint32)
    h = (h ^ (h >> np.uint32(16))) * np.uint32(0x85EBCA6B)
    h = (h ^ (h >> np.uint32(13))) * np.uint32(0xC2B2AE35)
    return h ^ (h >> np.uint32(16))


def gradient_noise(gx, gy, gz, seed) -> np.ndarray:
    """physics_utils.jl:24-28 (wrapping Int32 arithmetic), vectorised."""
    with np.errstate(over="ignore"):
        c = (np.asarray(gx).astype(np.uint32) * np.uint32(374761393) + np.asarray(gy).astype(np.uint32) * np.uint32(668265263)
             + np.asarray(gz).astype(np.uint32) * np.uint32(1274126177) + np.uint32(seed & 0xFFFFFFFF))
        h = gpu_hash(c)
    return (h & np.uint32(0xFFFF)).astype(np.float32) / np.float32(32768.0) - np.float32(1.0)


def box_topology(nbx: int, nby: int, nbz: int, periodic_y: bool = True, periodic_z: bool = True, periodic_x: bool = False):
    """All blocks of an nbx x nby x nbz box, lexicographically sorted (bx major, bz fastest)."""
    bx, by, bz = np.meshgrid(np.arange(1, nbx + 1), np.arange(1, nby + 1), np.arange(1, nbz + 1), indexing="ij")
    coords = np.stack([bx.ravel(), by.ravel(), bz.ravel()], axis=1).astype(np.int32)  # already lexicographic
    nb = coords.shape[0]
    idx_of = lambda x, y, z: ((x - 1) * nby + (y - 1)) * nbz + (z - 1) + 1  # noqa: E731  (1-based)
    block_pointer = np.zeros((nbz, nby, nbx), dtype=np.int32)
    block_pointer[coords[:, 2] - 1, coords[:, 1] - 1, coords[:, 0] - 1] = np.arange(1, nb + 1, dtype=np.int32)
    nt = np.zeros((27, nb), dtype=np.int32)
    for d in range(27):
        x, y, z = coords[:, 0] + CX[d], coords[:, 1] + CY[d], coords[:, 2] + CZ[d]
        ok = np.ones(nb, dtype=bool)
        if periodic_x: x = (x - 1) % nbx + 1
        else: ok &= (x >= 1) & (x <= nbx)
        if periodic_y: y = (y - 1) % nby + 1
        else: ok &= (y >= 1) & (y <= nby)
        if periodic_z: z = (z - 1) % nbz + 1
        else: ok &= (z >= 1) & (z <= nbz)
        nt[d] = np.where(ok, idx_of(np.clip(x, 1, nbx), np.clip(y, 1, nby), np.clip(z, 1, nbz)), 0).astype(np.int32)
    return coords, block_pointer, nt


def make_box_level(nbx: int, nby: int, nbz: int, tau: float = 0.5006, dx: float = 1.0, level_id: int = 1,
                   periodic_y: bool = True, periodic_z: bool = True, temporal_storage: bool = False) -> BlockLevel:
    """obstacle-free box, sponge 0, wall_dist 100, x faces open (inlet/outlet
    equilibrium BCs, physics_kernels.jl:99-113), y/z periodic through the neighbour table."""
    coords, bp, nt = box_topology(nbx, nby, nbz, periodic_y, periodic_z)
    nb = coords.shape[0]
    return BlockLevel(level_id=level_id, dx=dx, tau=tau, block_pointer=bp, neighbor_table=nt, active_block_coords=coords,
                      obstacle=np.zeros((nb, 8, 8, 8), np.uint8), sponge=np.zeros((nb, 8, 8, 8), np.float32),
                      wall_dist=np.full((nb, 8, 8, 8), 100.0, np.float32), temporal_storage=temporal_storage)


def global_coords(level: BlockLevel):
    """1-based global cell coordinates gx,gy,gz as int32 arrays of shape [nb,8,8,8] (z,y,x order)."""
    c = level.active_block_coords.astype(np.int32)
    l = np.arange(1, 9, dtype=np.int32)
    gx = ((c[:, 0] - 1) * 8)[:, None, None, None] + l[None, None, None, :] + np.zeros((1, 8, 8, 1), np.int32)
    gy = ((c[:, 1] - 1) * 8)[:, None, None, None] + l[None, None, :, None] + np.zeros((1, 8, 1, 8), np.int32)
    gz = ((c[:, 2] - 1) * 8)[:, None, None, None] + l[None, :, None, None] + np.zeros((1, 1, 8, 8), np.int32)
    return gx, gy, gz


def equilibrium(rho, ux, uy, uz) -> np.ndarray:
    """calculate_equilibrium (physics_utils.jl:34-39) for all 27 directions -> [27, ...] float32."""
    rho, ux, uy, uz = (np.asarray(a, np.float32) for a in (rho, ux, uy, uz))
    usq = ux * ux + uy * uy + uz * uz
    out = np.empty((27,) + rho.shape, np.float32)
    for k in range(27):
        cu = np.float32(CX[k]) * ux + np.float32(CY[k]) * uy + np.float32(CZ[k]) * uz
        out[k] = rho * W[k] * (np.float32(1) + np.float32(3) * cu + np.float32(4.5) * cu * cu - np.float32(1.5) * usq)
    return out


def noise_state(level: BlockLevel, amp_rho: float = 0.01, u0: float = 0.03, amp_u: float = 0.003):
    """Initial state of config 2: f = f_eq(rho = 1 + amp_rho*n1, u = (u0 + amp_u*n2, amp_u*n3, amp_u*n4)),
    n_i = gradient_noise(gx,gy,gz,seed=i) — the reference's own hash, so every backend generates identical bits."""
    gx, gy, gz = global_coords(level)
    n = [gradient_noise(gx, gy, gz, s) for s in (1, 2, 3, 4)]
    rho = (np.float32(1) + np.float32(amp_rho) * n[0]).astype(np.float32)
    ux = (np.float32(u0) + np.float32(amp_u) * n[1]).astype(np.float32)
    uy = (np.float32(amp_u) * n[2]).astype(np.float32)
    uz = (np.float32(amp_u) * n[3]).astype(np.float32)
    f = equilibrium(rho, ux, uy, uz)
    vel = np.stack([ux, uy, uz], axis=0)
    return f, rho, vel


# ----------------------------------------------------------------------------------------------------
# Synthetic multi-level case for differential tests (no STL): exercises every branch of K1/K2/K3/K4.

def sub_level(parent: BlockLevel, lo, hi, tau: float, level_id: int, temporal_storage: bool = True) -> BlockLevel:
    """Child level made of the 8 children of every parent block with lo <= (bx,by,bz) <= hi (1-based, inclusive),
    built with the reference's conventions: children of block b are 2b-1, 2b (domain.jl:103-110), blocks sorted
    lexicographically (domain.jl:171), neighbour table 0 where no block (domain_topology.jl:135-160),
    block_pointer extents = max active coordinate (blocks.jl:104-115)."""
    coords = []
    for bx in range(lo[0], hi[0] + 1):
        for by in range(lo[1], hi[1] + 1):
            for bz in range(lo[2], hi[2] + 1):
                for dbx in (0, 1):
                    for dby in (0, 1):
                        for dbz in (0, 1):
                            coords.append((2 * bx - 1 + dbx, 2 * by - 1 + dby, 2 * bz - 1 + dbz))
    coords = np.array(sorted(coords), dtype=np.int32)
    nb = coords.shape[0]
    dimx, dimy, dimz = (int(coords[:, i].max()) for i in range(3))
    bp = np.zeros((dimz, dimy, dimx), np.int32)
    bp[coords[:, 2] - 1, coords[:, 1] - 1, coords[:, 0] - 1] = np.arange(1, nb + 1, dtype=np.int32)
    nt = np.zeros((27, nb), np.int32)
    for d in range(27):
        x, y, z = coords[:, 0] + CX[d], coords[:, 1] + CY[d], coords[:, 2] + CZ[d]
        ok = (x >= 1) & (x <= dimx) & (y >= 1) & (y <= dimy) & (z >= 1) & (z <= dimz)
        nt[d] = np.where(ok, bp[np.clip(z, 1, dimz) - 1, np.clip(y, 1, dimy) - 1, np.clip(x, 1, dimx) - 1], 0)
    return BlockLevel(level_id=level_id, dx=parent.dx / 2, tau=tau, block_pointer=bp, neighbor_table=nt,
                      active_block_coords=coords, obstacle=np.zeros((nb, 8, 8, 8), np.uint8),
                      sponge=np.zeros((nb, 8, 8, 8), np.float32), wall_dist=np.full((nb, 8, 8, 8), 100.0, np.float32),
                      temporal_storage=temporal_storage)


def add_sphere_obstacle(level: BlockLevel, centre, radius: float, seed: int = 7, with_bouzidi: bool = True):
    """Marks the cells inside a sphere (global cell coordinates of this level) as obstacle, sets wall_dist on the
    fluid cells touching it and builds a Bouzidi q_map / boundary-cell list (q from the analytic sphere
    intersection, rounded to Float16 like bouzidi_setup.jl:128)."""
    gx, gy, gz = global_coords(level)
    px, py, pz = gx - 0.5, gy - 0.5, gz - 0.5   # cell centres in lattice units
    r2 = (px - centre[0]) ** 2 + (py - centre[1]) ** 2 + (pz - centre[2]) ** 2
    level.obstacle[...] = (r2 < radius * radius).astype(np.uint8)
    dist = np.sqrt(r2) - radius
    near = (level.obstacle == 0) & (dist < 1.8)
    level.wall_dist[...] = np.where(near, np.maximum(dist, 0.05) * level.dx, 100.0).astype(np.float32)
    if not with_bouzidi:
        return level
    nb = level.n_blocks
    q = np.zeros((27, nb, 8, 8, 8), np.float64)
    for k in range(27):
        if k == 13:
            continue
        c = np.array([CX[k], CY[k], CZ[k]], np.float64)
        cn = np.linalg.norm(c)
        d = c / cn
        ox, oy, oz = px - centre[0], py - centre[1], pz - centre[2]
        bq = ox * d[0] + oy * d[1] + oz * d[2]
        disc = bq * bq - (r2 - radius * radius)
        t = -bq - np.sqrt(np.maximum(disc, 0.0))
        hit = (disc > 0) & (t > 1e-9)
        qq = np.where(hit, t / cn, 0.0)
        q[k] = np.where((qq > 0) & (qq <= 1.0), qq, 0.0)
    has = (q > 0).any(axis=0)
    b, z, y, x = np.nonzero(has)   # ordered by (b,z,y,x): the reference's single-thread order
    level.q_map = q.astype(np.float16)
    level.tri_map = np.zeros((27, nb, 8, 8, 8), np.int32)
    level.cell_block = (b + 1).astype(np.int32)
    level.cell_x, level.cell_y, level.cell_z = ((v + 1).astype(np.int8) for v in (x, y, z))
    level.n_boundary_cells = int(len(b))
    level.bouzidi_enabled = level.n_boundary_cells > 0
    return level


def add_outlet_sponge(level: BlockLevel, nx_cells: int, frac: float = 0.25):
    """Cosine outlet sponge like apply_sponge! (domain_generation.jl:215-289), in lattice units."""
    gx, _, _ = global_coords(level)
    px = gx - 0.5
    start = nx_cells * (1.0 - frac)
    s = np.where(px > start, 0.5 * (1.0 + np.cos(np.pi * (nx_cells - px) / (nx_cells * frac))), 0.0)
    level.sponge[...] = s.astype(np.float32)
    return level
