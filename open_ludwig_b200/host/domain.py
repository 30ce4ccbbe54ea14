"""Multi-level sparse-block domain build: host-side mirror of the kept Julia driver files ``domain.jl``
(:20-266), ``domain_topology.jl`` (:54-160) and ``blocks.jl`` (:89-188).  The heavy per-cell geometry work
(SAT voxelisation, flood fill, sponge, wall distance, Bouzidi q-map) is in ``domain_build.cpp``.

Block sets are kept as dense boolean grids indexed [bx-1, by-1, bz-1]; ``np.argwhere`` on such a grid yields
the blocks in the reference's order ``sort(collect(active_set))`` (lexicographic, bx major — domain.jl:171).
Every table this produces is in the reference's own convention (1-based, 0 = none) so that it can be diffed
bit-for-bit; the golden counts of RESULTS_SPHERE_RE1M.txt / RESULTS_SPHERE_RE10M.txt are checked in
tests/test_domain_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from ..cabi import BlockLevel
from .config import CaseConfig, DomainParameters, compute_domain_from_mesh, load_case_configuration
from .geometry import SolverMesh, load_mesh

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libludwig_host.so")
_lib = None


def host_lib() -> C.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "domain_build.cpp")
        if not os.path.exists(_LIB_PATH) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        lib = C.CDLL(_LIB_PATH)
        vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
        lib.ludwig_host_mark_surface_blocks.argtypes = [vp, i64, vp, f64, i32, i32, i32, vp]
        lib.ludwig_host_mark_surface_blocks.restype = None
        lib.ludwig_host_voxelize.argtypes = [vp, i64, vp, i32, f64, vp, vp]
        lib.ludwig_host_voxelize.restype = None
        lib.ludwig_host_flood_fill.argtypes = [vp, vp, i32, vp, i32, i32, i32]
        lib.ludwig_host_flood_fill.restype = i64
        lib.ludwig_host_sponge.argtypes = [vp, i32, f64, f64, f64, f64, f64, i32, vp]
        lib.ludwig_host_sponge.restype = None
        lib.ludwig_host_wall_distance.argtypes = [vp, i32, vp, f64, vp]
        lib.ludwig_host_wall_distance.restype = i64
        lib.ludwig_host_qmap.argtypes = [vp, i64, vp, i32, f64, vp, i64, vp, vp, vp]
        lib.ludwig_host_qmap.restype = i64
        lib.ludwig_host_scatter_qmap.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp]
        lib.ludwig_host_scatter_qmap.restype = None
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class LevelReport:
    """The numbers the reference prints per level (the golden logs' KATs)."""
    level: int
    n_blocks: int
    halo_added: int
    filled_voxels: int
    shell_voxels: int
    near_wall_cells: int
    n_boundary_cells: int
    active_links: int


@dataclass
class Domain:
    cfg: CaseConfig
    params: DomainParameters
    mesh: SolverMesh
    levels: List[BlockLevel]
    reports: List[LevelReport] = field(default_factory=list)

    @property
    def total_cells(self) -> int:
        return sum(lv.n_cells for lv in self.levels)

    @property
    def cell_updates_per_coarse_step(self) -> int:
        return sum(lv.n_cells * 2 ** (lv.level_id - 1) for lv in self.levels)


# ---- domain_topology.jl -----------------------------------------------------------------------------

def _dilate26(a: np.ndarray) -> np.ndarray:
    p = np.pad(a, 1)
    out = np.zeros_like(a)
    n0, n1, n2 = a.shape
    for dx in (0, 1, 2):
        for dy in (0, 1, 2):
            for dz in (0, 1, 2):
                out |= p[dx:dx + n0, dy:dy + n1, dz:dz + n2]
    return out


def _complete_octets(a: np.ndarray) -> np.ndarray:
    """All siblings ((b+1)÷2 equal) of the set blocks, clipped to the grid."""
    n0, n1, n2 = a.shape
    p = np.pad(a, ((0, n0 % 2), (0, n1 % 2), (0, n2 % 2)))
    par = p.reshape(p.shape[0] // 2, 2, p.shape[1] // 2, 2, p.shape[2] // 2, 2).any(axis=(1, 3, 5))
    full = np.repeat(np.repeat(np.repeat(par, 2, 0), 2, 1), 2, 2)
    return full[:n0, :n1, :n2]


def add_halo_blocks_with_siblings(active: np.ndarray, layers: int) -> np.ndarray:
    """domain_topology.jl:54-99."""
    active = active.copy()
    for _ in range(layers):
        new = _dilate26(active) & ~active
        sib = _complete_octets(new) & ~active & ~new
        active |= new | sib
    return active


def ensure_complete_parent_coverage(active: np.ndarray) -> np.ndarray:
    """domain_topology.jl:101-133 (a fixpoint: completing octets never creates an incomplete one)."""
    return active | _complete_octets(active)


def build_neighbor_table(coords: np.ndarray, dims) -> tuple:
    """domain_topology.jl:135-160 -> (neighbor_table [27,nb], full-extent pointer grid [bx,by,bz])."""
    nb = coords.shape[0]
    ptr = np.zeros(tuple(d + 2 for d in dims), np.int32)
    ptr[coords[:, 0], coords[:, 1], coords[:, 2]] = np.arange(1, nb + 1, dtype=np.int32)   # padded by 1
    nt = np.zeros((27, nb), np.int32)
    for d in range(27):
        dx, dy, dz = d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1
        nt[d] = ptr[coords[:, 0] + dx, coords[:, 1] + dy, coords[:, 2] + dz]
    return nt, ptr[1:-1, 1:-1, 1:-1]


# ---- domain.jl ----------------------------------------------------------------------------------------

def setup_multilevel_domain(cfg: CaseConfig, mesh: Optional[SolverMesh] = None, verbose: bool = False,
                            build_tri_map: bool = True, gpu_device: Optional[int] = None) -> Domain:
    """domain.jl:20-280.  gpu_device: run voxelisation, flood fill, wall distance and q-map ray casting on that
    CUDA device through libludwig_b200.so's ludwig_domain_* entry points (N2) instead of the host threads; the tables are
    byte-identical either way (tests/test_domain_gpu.py)."""
    lib = host_lib()
    glib = None
    if gpu_device is not None:
        from .. import cabi
        glib = cabi.load_library()

    def gcheck(rc, what):
        if rc < 0:
            raise RuntimeError(f"{what} failed ({rc}): {glib.ludwig_domain_last_error().decode()}")
        return rc
    if mesh is None:
        stl = os.path.join(cfg.case_dir, cfg.stl_file)
        if not os.path.isfile(stl):
            stl = os.path.join(cfg.case_dir, "model.stl")
        mesh = load_mesh(stl, scale=cfg.stl_scale)
    params = compute_domain_from_mesh(cfg, mesh.min_bounds, mesh.max_bounds)
    tris = np.ascontiguousarray(mesh.triangles, np.float64)
    n_tri = tris.shape[0]
    off = np.array(params.mesh_offset, np.float64)
    num_levels = params.num_levels
    pmin = np.array(params.mesh_min) + off
    pmax = np.array(params.mesh_max) + off
    wake_start_x = pmax[0] - (params.reference_length * 0.1)
    wake_end_x = pmax[0] + (params.reference_length * cfg.wake_length)
    wcy, wcz = (pmin[1] + pmax[1]) / 2.0, (pmin[2] + pmax[2]) / 2.0
    ww, wh = (pmax[1] - pmin[1]) * cfg.wake_width_factor, (pmax[2] - pmin[2]) * cfg.wake_height_factor
    wake_min_y, wake_max_y = wcy - ww / 2.0, wcy + ww / 2.0
    wake_min_z, wake_max_z = wcz - wh / 2.0, wcz + wh / 2.0

    levels: List[BlockLevel] = []
    reports: List[LevelReport] = []
    prev_active = None
    import time as _time
    phase_s: dict = {}

    def timed(name, fn, *a):
        t0 = _time.time(); r = fn(*a); phase_s[name] = phase_s.get(name, 0.0) + _time.time() - t0
        return r
    for lvl in range(1, num_levels + 1):
        scale = 2 ** (lvl - 1)
        dx = params.dx_coarse / scale
        tau = params.tau_levels[lvl - 1]
        dims = (params.bx_max * scale, params.by_max * scale, params.bz_max * scale)
        if lvl == 1:
            active = np.ones(dims, bool)
        else:
            prev_dx = params.dx_coarse / (2 ** (lvl - 2))
            prev_bs_phys = 8 * prev_dx
            if cfg.refinement_strategy != "geometry_first":
                raise NotImplementedError("only the geometry_first refinement strategy (all shipped cases) is restated")
            grid = np.zeros(dims, np.uint8)
            timed("mark_surface_blocks", lib.ludwig_host_mark_surface_blocks, _p(tris), n_tri, _p(off), dx, dims[0], dims[1], dims[2], _p(grid))
            active = grid.astype(bool)
            if cfg.wake_enabled:                      # domain.jl:88-112
                pc = levels[-1].active_block_coords.astype(np.float64)
                bmin, bmax = (pc - 1) * prev_bs_phys, pc * prev_bs_phys
                ov = ((bmin[:, 0] <= wake_end_x) & (bmax[:, 0] >= wake_start_x) & (bmin[:, 1] <= wake_max_y) & (bmax[:, 1] >= wake_min_y)
                      & (bmin[:, 2] <= wake_max_z) & (bmax[:, 2] >= wake_min_z))
                for cbx, cby, cbz in levels[-1].active_block_coords[ov]:
                    active[2 * cbx - 2:2 * cbx, 2 * cby - 2:2 * cby, 2 * cbz - 2:2 * cbz] = True
            # drop blocks whose parent (b+1)÷2 is not active on the previous level (domain.jl:114-127)
            par = np.repeat(np.repeat(np.repeat(prev_active, 2, 0), 2, 1), 2, 2)
            active &= par
        n_before = int(active.sum())
        active = add_halo_blocks_with_siblings(active, cfg.refinement_margin)
        active = ensure_complete_parent_coverage(active)
        n_after = int(active.sum())
        coords = (np.argwhere(active) + 1).astype(np.int32)      # lexicographic, 1-based
        nb = coords.shape[0]
        nt, full_ptr = timed("neighbor_table", build_neighbor_table, coords, dims)
        full_ptr_cm = np.ascontiguousarray(full_ptr.transpose(2, 1, 0))   # Julia [bx,by,bz] column-major bytes

        obstacle = np.zeros((nb, 8, 8, 8), np.uint8)
        sponge = np.zeros((nb, 8, 8, 8), np.float32)
        wall_dist = np.full((nb, 8, 8, 8), 100.0, np.float32)
        cflat = np.ascontiguousarray(coords)
        grid_ptr = np.ascontiguousarray(full_ptr, np.int32)       # [bx][by][bz], 1-based block index, 0 = none
        if glib is not None:
            gcheck(timed("voxelize", glib.ludwig_domain_voxelize, gpu_device, _p(tris), n_tri, _p(off), dx, _p(cflat), nb, _p(grid_ptr), dims[0], dims[1], dims[2],
                         _p(obstacle)), "ludwig_domain_voxelize")
        else:
            timed("voxelize", lib.ludwig_host_voxelize, _p(tris), n_tri, _p(cflat), nb, dx, _p(off), _p(obstacle))
        shell = int(obstacle.sum())
        if glib is not None:
            filled = int(gcheck(timed("flood_fill", glib.ludwig_domain_flood_fill, gpu_device, _p(cflat), nb, _p(grid_ptr), dims[0], dims[1], dims[2], _p(obstacle)),
                                "ludwig_domain_flood_fill"))
        else:
            filled = int(timed("flood_fill", lib.ludwig_host_flood_fill, _p(obstacle), _p(cflat), nb, _p(full_ptr_cm), dims[0], dims[1], dims[2]))
        timed("sponge", lib.ludwig_host_sponge, _p(cflat), nb, dx, params.domain_size[0], params.domain_size[1], params.domain_size[2],
              float(cfg.sponge_thickness), int(cfg.symmetric), _p(sponge))
        near = 0
        if cfg.wall_model_enabled:
            if glib is not None:
                nt_c = np.ascontiguousarray(nt, np.int32)
                near = int(gcheck(timed("wall_distance", glib.ludwig_domain_wall_distance, gpu_device, _p(nt_c), nb, _p(obstacle), dx, _p(wall_dist)), "ludwig_domain_wall_distance"))
            else:
                near = int(timed("wall_distance", lib.ludwig_host_wall_distance, _p(cflat), nb, _p(obstacle), dx, _p(wall_dist)))

        use_bouzidi = cfg.boundary_method == "bouzidi" and lvl > (num_levels - cfg.bouzidi_levels)   # bouzidi_common.jl:28-34
        q_map = tri_map = cell_block = cell_x = cell_y = cell_z = None
        n_bc = links = 0
        if use_bouzidi:
            def qmap(cap, c_, q_, t_):
                if glib is not None:
                    return int(gcheck(glib.ludwig_domain_qmap(gpu_device, _p(tris), n_tri, _p(off), dx, _p(cflat), nb, _p(grid_ptr), dims[0], dims[1],
                                                              dims[2], cap, c_, q_, t_), "ludwig_domain_qmap"))
                return int(lib.ludwig_host_qmap(_p(tris), n_tri, _p(cflat), nb, dx, _p(off), cap, c_, q_, t_))
            n_bc = timed("qmap", qmap, 0, None, None, None)
            cells = np.zeros((max(n_bc, 1), 4), np.int32)
            qv = np.zeros((max(n_bc, 1), 27), np.float64)
            tv = np.zeros((max(n_bc, 1), 27), np.int32)
            got = timed("qmap", qmap, n_bc, _p(cells), _p(qv), _p(tv))
            assert got == n_bc
            cells, qv, tv = cells[:n_bc], qv[:n_bc], tv[:n_bc]
            q_map = np.zeros((27, nb, 8, 8, 8), np.float16)
            # tri_map (bouzidi_setup.jl:85) is never read by a kernel: 108 B/cell of host memory, optional for huge levels
            tri_map = np.zeros((27, nb, 8, 8, 8), np.int32) if build_tri_map else None
            q16 = qv.astype(np.float16)               # Float16(q) round-to-nearest-even (bouzidi_setup.jl:128)
            cells_c, qv_c, q16_c, tv_c = (np.ascontiguousarray(a) for a in (cells, qv, q16, tv))
            timed("scatter_qmap", lib.ludwig_host_scatter_qmap, _p(cells_c), _p(qv_c), _p(q16_c), _p(tv_c) if tri_map is not None else None, n_bc, nb,
                  _p(q_map), _p(tri_map) if tri_map is not None else None)
            cell_block = cells[:, 0].astype(np.int32)
            cell_x, cell_y, cell_z = (cells[:, i].astype(np.int8) for i in (1, 2, 3))
            qf = q16.astype(np.float32)
            links = int(((qf > np.float32(cfg.q_min_threshold)) & (qf <= 1.0)).sum())

        # BlockLevel ctor (blocks.jl:89-188): block_pointer extents = max active coordinate per axis
        mx = coords.max(axis=0)
        bp = np.ascontiguousarray(full_ptr[:mx[0], :mx[1], :mx[2]].transpose(2, 1, 0))
        level = BlockLevel(level_id=lvl, dx=float(np.float32(dx)), tau=float(tau), block_pointer=bp, neighbor_table=nt,
                           active_block_coords=coords, obstacle=obstacle, sponge=sponge, wall_dist=wall_dist,
                           temporal_storage=bool(cfg.temporal_interpolation),
                           bouzidi_enabled=bool(use_bouzidi and n_bc > 0), n_boundary_cells=n_bc, q_map=q_map, tri_map=tri_map,
                           cell_block=cell_block, cell_x=cell_x, cell_y=cell_y, cell_z=cell_z)
        levels.append(level)
        reports.append(LevelReport(lvl, nb, n_after - n_before, filled, shell, near, n_bc, links))
        if verbose:
            print(f"--- Level {lvl} --- blocks {nb} (+{n_after - n_before} halo) shell {shell} filled {filled} near-wall {near} "
                  f"boundary cells {n_bc} links {links}", flush=True)
        prev_active = active
    if verbose:
        print("domain build phases [s]: " + ", ".join(f"{k} {v:.2f}" for k, v in phase_s.items()), flush=True)
    dom = Domain(cfg, params, mesh, levels, reports)
    dom.phase_s = phase_s
    return dom


def load_case(case_dir: str, overrides: Optional[dict] = None, verbose: bool = False, build_tri_map: bool = True,
              gpu_device: Optional[int] = None) -> Domain:
    """load_case_configuration + setup_multilevel_domain (main.jl:259-260, :90)."""
    return setup_multilevel_domain(load_case_configuration(case_dir, overrides), verbose=verbose, build_tri_map=build_tri_map,
                                   gpu_device=gpu_device)


# ---- domain cache (multi-rank launches: one rank builds, the others map the arrays) -------------------

_LEVEL_ARRAYS = ("block_pointer", "neighbor_table", "active_block_coords", "obstacle", "sponge", "wall_dist",
                 "q_map", "tri_map", "cell_block", "cell_x", "cell_y", "cell_z")


def save_domain(dom: Domain, folder: str) -> None:
    """Write a built Domain to `folder`: every level array as its own .npy, the small host objects pickled.
    The domain build is deterministic, so this is purely a cache: with N processes on one box (one per GPU) rank 0
    builds with all host threads and the others load_domain() — page-cache shared, read-only memory maps."""
    import pickle
    os.makedirs(folder, exist_ok=True)
    meta = []
    for i, lv in enumerate(dom.levels):
        scal = {k: v for k, v in vars(lv).items() if k not in _LEVEL_ARRAYS}
        have = []
        for k in _LEVEL_ARRAYS:
            a = getattr(lv, k)
            if a is not None:
                np.save(os.path.join(folder, f"L{i}_{k}.npy"), np.ascontiguousarray(a))
                have.append(k)
        meta.append((scal, have))
    with open(os.path.join(folder, "domain.pkl.tmp"), "wb") as f:
        pickle.dump({"cfg": dom.cfg, "params": dom.params, "mesh": dom.mesh, "reports": dom.reports, "levels": meta}, f, protocol=5)
    os.replace(os.path.join(folder, "domain.pkl.tmp"), os.path.join(folder, "domain.pkl"))   # written last: marks the cache complete


def load_domain(folder: str) -> Domain:
    import pickle
    with open(os.path.join(folder, "domain.pkl"), "rb") as f:
        d = pickle.load(f)
    levels = []
    for i, (scal, have) in enumerate(d["levels"]):
        arrs = {k: (np.load(os.path.join(folder, f"L{i}_{k}.npy"), mmap_mode="r") if k in have else None) for k in _LEVEL_ARRAYS}
        levels.append(BlockLevel(**scal, **arrs))
    return Domain(cfg=d["cfg"], params=d["params"], mesh=d["mesh"], levels=levels, reports=d["reports"])
