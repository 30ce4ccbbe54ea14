"""ctypes binding of the C ABI declared in include/ludwig_b200.h.

This is the Python twin of the Julia ``ccall`` glue in ``open_ludwig_b200/julia/LudwigB200.jl``
(Julia is not installed in this image, so every test drives the library through this module).

The default library is the CUDA product library ``open_ludwig_b200/csrc/libludwig_b200.so``.
There is NO CPU fallback: if that library is missing, loading fails loudly.  A different
library implementing the same ABI can be passed explicitly by path (the tests do that with the
CPU parity oracle); nothing in this package refers to such a library on its own.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "libludwig_b200.so")

# which-codes (include/ludwig_b200.h)
F, F_TEMP, F_POST, F_OLD, RHO, RHO_OLD, VEL, VEL_TEMP, VEL_OLD, OBSTACLE = range(10)
_NCOMP = {F: 27, F_TEMP: 27, F_POST: 27, F_OLD: 27, RHO: 1, RHO_OLD: 1, VEL: 3, VEL_TEMP: 3, VEL_OLD: 3, OBSTACLE: 1}

EXPORTED_SYMBOLS = (
    "ludwig_ctx_create", "ludwig_ctx_destroy", "ludwig_last_error", "ludwig_backend_name", "ludwig_sync",
    "ludwig_level_create", "ludwig_num_levels", "ludwig_level_upload", "ludwig_level_download",
    "ludwig_mesh_create", "ludwig_mesh_destroy", "ludwig_forces_create", "ludwig_forces_destroy",
    "ludwig_init_equilibrium", "ludwig_step_batch", "ludwig_level_step", "ludwig_level_snapshot_old",
    "ludwig_compute_aerodynamics", "ludwig_forces_download_maps", "ludwig_flow_stats", "ludwig_device_bytes",
    "ludwig_ctx_stream", "ludwig_launch_count", "ludwig_profile_enable", "ludwig_profile_read", "ludwig_profile_classes",
    "ludwig_partition_starts", "ludwig_block_costs", "ludwig_ctx_set_partition", "ludwig_partition_plan",
    "ludwig_ctx_set_partition_keys", "ludwig_set_barrier_callback", "ludwig_level_local_blocks",
    "ludwig_ipc_export", "ludwig_ipc_attach", "ludwig_level_upload_local", "ludwig_level_download_local",
    "ludwig_attach_inprocess", "ludwig_profile_levels", "ludwig_output_gather", "ludwig_partition_rcb",
    "ludwig_partition_rcb_axes", "ludwig_ctx_set_option",
    "ludwig_multi_create", "ludwig_multi_destroy", "ludwig_multi_last_error", "ludwig_multi_num_ranks", "ludwig_multi_ctx",
    "ludwig_multi_set_option", "ludwig_multi_set_partition_plan", "ludwig_multi_level_create", "ludwig_multi_level_upload",
    "ludwig_multi_level_download", "ludwig_multi_init_equilibrium", "ludwig_multi_step_batch", "ludwig_multi_sync",
    "ludwig_multi_flow_stats", "ludwig_multi_forces_create", "ludwig_multi_compute_aerodynamics",
    "ludwig_multi_forces_download_maps", "ludwig_multi_device_bytes",
    "ludwig_domain_last_error", "ludwig_domain_voxelize", "ludwig_domain_flood_fill", "ludwig_domain_wall_distance", "ludwig_domain_qmap",
    "ludwig_ctx_self_check", "ludwig_multi_self_check",
    "ludwig_graph_replays", "ludwig_init_uniform_flow", "ludwig_multi_init_uniform_flow",
    "ludwig_output_valid_blocks", "ludwig_output_export", "ludwig_multi_output_valid_blocks", "ludwig_multi_output_export",
)

BARRIER_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)   # returns 0 on success (include/ludwig_b200.h)


class LudwigError(RuntimeError):
    """Non-zero return from the library (the Julia side turns it into ``error(...)``,
    matching the per-case try/catch of main.jl:261-267)."""


class LevelDesc(C.Structure):
    _fields_ = [
        ("level_id", C.c_int32), ("n_blocks", C.c_int32),
        ("dim_x", C.c_int32), ("dim_y", C.c_int32), ("dim_z", C.c_int32),
        ("tau", C.c_float), ("dx", C.c_double),
        ("block_pointer", C.c_void_p), ("neighbor_table", C.c_void_p),
        ("map_x", C.c_void_p), ("map_y", C.c_void_p), ("map_z", C.c_void_p),
        ("obstacle", C.c_void_p), ("sponge", C.c_void_p), ("wall_dist", C.c_void_p),
        ("temporal_storage", C.c_int32), ("bouzidi_enabled", C.c_int32), ("n_boundary_cells", C.c_int32),
        ("q_map_f16", C.c_void_p), ("cell_block", C.c_void_p),
        ("cell_x", C.c_void_p), ("cell_y", C.c_void_p), ("cell_z", C.c_void_p),
    ]


class Params(C.Structure):
    """ludwig_params: the batch-constant scalar arguments of perform_timestep_v2!."""
    _fields_ = [
        ("c_wale", C.c_float), ("nu_sgs_bg", C.c_float), ("inlet_turbulence", C.c_float),
        ("q_min_threshold", C.c_float),
        ("wall_model_active", C.c_int32), ("use_temporal", C.c_int32), ("sponge_blend", C.c_int32),
        ("symmetric", C.c_int32),
        ("domain_nx", C.c_int32), ("domain_ny", C.c_int32), ("domain_nz", C.c_int32),
        ("strict_fp", C.c_int32),
    ]


def load_library(path: Optional[str] = None) -> C.CDLL:
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise LudwigError(
            f"{path} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
            "g.build()'` at the repo root). There is no CPU fallback on the product path.")
    lib = C.CDLL(path)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
    sig = {
        "ludwig_ctx_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "ludwig_ctx_destroy": (C.c_int, [vp]),
        "ludwig_last_error": (C.c_char_p, [vp]),
        "ludwig_backend_name": (C.c_char_p, []),
        "ludwig_sync": (C.c_int, [vp]),
        "ludwig_level_create": (C.c_int, [vp, C.POINTER(LevelDesc), C.POINTER(i32)]),
        "ludwig_num_levels": (C.c_int, [vp]),
        "ludwig_level_upload": (C.c_int, [vp, i32, i32, vp]),
        "ludwig_level_download": (C.c_int, [vp, i32, i32, vp]),
        "ludwig_mesh_create": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
        "ludwig_mesh_destroy": (C.c_int, [vp]),
        "ludwig_forces_create": (C.c_int, [vp, vp, f64, f64, f64, f64, C.POINTER(f64), i32, C.POINTER(vp)]),
        "ludwig_forces_destroy": (C.c_int, [vp]),
        "ludwig_init_equilibrium": (C.c_int, [vp]),
        "ludwig_step_batch": (C.c_int, [vp, i64, i32, f32, C.POINTER(Params)]),
        "ludwig_level_step": (C.c_int, [vp, i32, i64, i64, f32, f32, C.POINTER(Params)]),
        "ludwig_level_snapshot_old": (C.c_int, [vp, i32, i64]),
        "ludwig_compute_aerodynamics": (C.c_int, [vp, vp, i32, C.POINTER(f64), f64, f64, i32, C.POINTER(f64)]),
        "ludwig_forces_download_maps": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "ludwig_flow_stats": (C.c_int, [vp, i32, C.POINTER(f64)]),
        "ludwig_device_bytes": (C.c_int64, [vp]),
        "ludwig_ctx_stream": (vp, [vp]),
        "ludwig_launch_count": (C.c_int64, [vp]),
        "ludwig_profile_enable": (C.c_int, [vp, i32]),
        "ludwig_profile_read": (C.c_int, [vp, C.POINTER(f64), C.POINTER(i64), C.POINTER(i64)]),
        "ludwig_profile_classes": (C.c_int, [vp, C.POINTER(f64)]),
        "ludwig_partition_starts": (C.c_int, [i32, i32, vp]),
        "ludwig_block_costs": (C.c_int, [C.POINTER(LevelDesc), vp]),
        "ludwig_ctx_set_partition": (C.c_int, [vp, i32, i32]),
        "ludwig_partition_plan": (C.c_int, [vp, i32, i32, vp]),
        "ludwig_ctx_set_partition_keys": (C.c_int, [vp, vp, i32]),
        "ludwig_set_barrier_callback": (C.c_int, [vp, BARRIER_CB, vp]),
        "ludwig_level_local_blocks": (C.c_int, [vp, i32, C.POINTER(i32), vp]),
        "ludwig_level_upload_local": (C.c_int, [vp, i32, i32, vp]),
        "ludwig_level_download_local": (C.c_int, [vp, i32, i32, vp]),
        "ludwig_ipc_export": (C.c_int, [vp, vp, i64, C.POINTER(i64)]),
        "ludwig_ipc_attach": (C.c_int, [vp, vp, i64]),
        "ludwig_attach_inprocess": (C.c_int, [vp, C.POINTER(vp), i32]),
        "ludwig_profile_levels": (C.c_int, [vp, C.POINTER(f64), i32]),
        "ludwig_output_gather": (C.c_int, [vp, i32, i64, vp, i32, vp, vp, vp]),
        "ludwig_partition_rcb": (C.c_int, [C.POINTER(LevelDesc), i32, vp]),
        "ludwig_partition_rcb_axes": (C.c_int, [C.POINTER(LevelDesc), i32, i32, vp]),
        "ludwig_ctx_set_option": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
        "ludwig_multi_create": (C.c_int, [C.POINTER(vp), i32, vp]),
        "ludwig_multi_destroy": (C.c_int, [vp]),
        "ludwig_multi_last_error": (C.c_char_p, [vp]),
        "ludwig_multi_num_ranks": (i32, [vp]),
        "ludwig_multi_ctx": (vp, [vp, i32]),
        "ludwig_multi_set_option": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
        "ludwig_multi_set_partition_plan": (C.c_int, [vp, vp, i32]),
        "ludwig_multi_level_create": (C.c_int, [vp, C.POINTER(LevelDesc), C.POINTER(i32)]),
        "ludwig_multi_level_upload": (C.c_int, [vp, i32, i32, vp]),
        "ludwig_multi_level_download": (C.c_int, [vp, i32, i32, vp]),
        "ludwig_multi_init_equilibrium": (C.c_int, [vp]),
        "ludwig_multi_step_batch": (C.c_int, [vp, i64, i32, f32, C.POINTER(Params)]),
        "ludwig_multi_sync": (C.c_int, [vp]),
        "ludwig_multi_flow_stats": (C.c_int, [vp, i32, C.POINTER(f64)]),
        "ludwig_multi_forces_create": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, f64, f64, f64, f64, C.POINTER(f64), i32, C.POINTER(i32)]),
        "ludwig_multi_compute_aerodynamics": (C.c_int, [vp, i32, i32, C.POINTER(f64), f64, f64, i32, C.POINTER(f64)]),
        "ludwig_multi_forces_download_maps": (C.c_int, [vp, i32, vp, vp, vp, vp]),
        "ludwig_multi_device_bytes": (C.c_int64, [vp]),
        "ludwig_domain_last_error": (C.c_char_p, []),
        "ludwig_domain_voxelize": (C.c_int, [C.c_int, vp, i64, vp, f64, vp, i32, vp, i32, i32, i32, vp]),
        "ludwig_domain_flood_fill": (C.c_int64, [C.c_int, vp, i32, vp, i32, i32, i32, vp]),
        "ludwig_domain_wall_distance": (C.c_int64, [C.c_int, vp, i32, vp, f64, vp]),
        "ludwig_domain_qmap": (C.c_int64, [C.c_int, vp, i64, vp, f64, vp, i32, vp, i32, i32, i32, i64, vp, vp, vp]),
        "ludwig_ctx_self_check": (C.c_int64, [vp]),
        "ludwig_multi_self_check": (C.c_int64, [vp]),
        "ludwig_graph_replays": (C.c_int64, [vp]),
        "ludwig_init_uniform_flow": (C.c_int, [vp, f32]),
        "ludwig_multi_init_uniform_flow": (C.c_int, [vp, f32]),
        "ludwig_output_valid_blocks": (C.c_int, [vp, vp, vp]),
        "ludwig_output_export": (C.c_int, [vp, i64, vp, vp, vp, vp]),
        "ludwig_multi_output_valid_blocks": (C.c_int, [vp, vp, vp]),
        "ludwig_multi_output_export": (C.c_int, [vp, i64, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype, fn.argtypes = res, args
    return lib


@dataclass
class BlockLevel:
    """Host mirror of the reference's ``BlockLevel`` (blocks.jl:16-65) as it leaves
    ``setup_multilevel_domain``.  Arrays are numpy C-order with REVERSED axes, which is
    byte-identical to the Julia column-major arrays:

        Julia obstacle[x,y,z,b]        <-> numpy obstacle[b,z,y,x]
        Julia neighbor_table[b,dir]    <-> numpy neighbor_table[dir,b]
        Julia block_pointer[bx,by,bz]  <-> numpy block_pointer[bz,by,bx]
        Julia q_map[x,y,z,b,k]         <-> numpy q_map[k,b,z,y,x]

    Index values stay 1-based with 0 = none, exactly as the reference builds them.
    """
    level_id: int
    dx: float
    tau: float
    block_pointer: np.ndarray          # int32 [dimz,dimy,dimx]
    neighbor_table: np.ndarray         # int32 [27,nb]
    active_block_coords: np.ndarray    # int32 [nb,3] (bx,by,bz) 1-based, sorted lexicographically
    obstacle: np.ndarray               # uint8 [nb,8,8,8]
    sponge: np.ndarray                 # float32 [nb,8,8,8]
    wall_dist: np.ndarray              # float32 [nb,8,8,8]
    temporal_storage: bool = True
    bouzidi_enabled: bool = False
    n_boundary_cells: int = 0
    q_map: Optional[np.ndarray] = None       # float16 [27,nb,8,8,8]
    tri_map: Optional[np.ndarray] = None     # int32   [27,nb,8,8,8]  (never read by a kernel; host only)
    cell_block: Optional[np.ndarray] = None  # int32 [n_bc] 1-based
    cell_x: Optional[np.ndarray] = None      # int8  [n_bc] 1-based
    cell_y: Optional[np.ndarray] = None
    cell_z: Optional[np.ndarray] = None

    @property
    def n_blocks(self) -> int:
        return int(self.active_block_coords.shape[0])

    @property
    def n_cells(self) -> int:
        return self.n_blocks * 512


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _as(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One solver context = the reference's ``grids`` vector + lattice arrays on one device."""

    def __init__(self, lib_path: Optional[str] = None, device: int = 0, options: Optional[dict] = None):
        self.lib = load_library(lib_path)
        self._h = C.c_void_p()
        rc = self.lib.ludwig_ctx_create(C.byref(self._h), device)
        if rc != 0:
            raise LudwigError(f"ludwig_ctx_create failed ({rc}): is a CUDA device visible?")
        self._cb_error: Optional[BaseException] = None
        for k, v in (options or {}).items():
            self.set_option(k, v)
        self.n_blocks: list[int] = []
        self.rank, self.world = 0, 1
        self._meshes: list[C.c_void_p] = []
        self._forces: list[C.c_void_p] = []

    # -- plumbing ---------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.ludwig_last_error(self._h)
            err = LudwigError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
            if self._cb_error is not None:      # an exception raised inside the barrier callback (ctypes cannot propagate it)
                raise err from self._cb_error
            raise err

    def set_option(self, key: str, value):
        """ludwig_ctx_set_option: every behaviour switch of the library (it reads no environment variable)."""
        if isinstance(value, bool):
            value = int(value)
        self._check(self.lib.ludwig_ctx_set_option(self._h, str(key).encode(), str(value).encode()), f"ludwig_ctx_set_option({key})")

    @property
    def backend(self) -> str:
        return self.lib.ludwig_backend_name().decode()

    def close(self):
        if self._h:
            for f in self._forces:
                self.lib.ludwig_forces_destroy(f)
            for m in self._meshes:
                self.lib.ludwig_mesh_destroy(m)
            self.lib.ludwig_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.ludwig_sync(self._h), "ludwig_sync")

    def output_gather(self, level: int, t_step: int, blocks0: np.ndarray):
        """io_vtk.jl:52-58,100-111 for the listed blocks (0-based reference indices): (rho_arr [512 n], vel_mat [512 n, 3]
        = Julia's Matrix{Float32}(3, 512 n), obst_arr [512 n]) exactly as the VTK writer fills them."""
        ids = _as(np.asarray(blocks0) + 1, np.int32)
        n = ids.size
        rho = np.empty(512 * n, np.float32); vel = np.empty((512 * n, 3), np.float32); obs = np.empty(512 * n, np.uint8)
        self._check(self.lib.ludwig_output_gather(self._h, level, t_step, _ptr(ids), n, _ptr(rho), _ptr(vel), _ptr(obs)), "ludwig_output_gather")
        return rho, vel, obs

    def device_bytes(self) -> int:
        return int(self.lib.ludwig_device_bytes(self._h))

    _VALID, _EXPORT = "ludwig_output_valid_blocks", "ludwig_output_export"

    def output_valid_blocks(self):
        """io_vtk.jl:17-46: per level, the 0-based reference indices of the blocks the VTK export writes."""
        return _valid_blocks(self, len(self.n_blocks))

    def output_export(self, t_step: int):
        """io_vtk.jl:52-58,100-111 for every valid block: (rho_arr [N], vel_mat [N, 3], obst_arr [N], level_arr [N])."""
        return _export(self, t_step, len(self.n_blocks))

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t of the context (0 for a CPU backend) — wrap with torch.cuda.ExternalStream to time on it."""
        return int(self.lib.ludwig_ctx_stream(self._h) or 0)

    def launch_count(self) -> int:
        return int(self.lib.ludwig_launch_count(self._h))

    def self_check(self) -> int:
        """ludwig_ctx_self_check: number of index-table violations (0 = clean)."""
        n = int(self.lib.ludwig_ctx_self_check(self._h))
        if n < 0:
            self._check(n, "ludwig_ctx_self_check")
        return n

    def graph_replays(self) -> int:
        return int(self.lib.ludwig_graph_replays(self._h))

    def init_uniform_flow(self, ux: float):
        self._check(self.lib.ludwig_init_uniform_flow(self._h, C.c_float(ux)), "ludwig_init_uniform_flow")

    def profile_enable(self, on: bool = True):
        self._check(self.lib.ludwig_profile_enable(self._h, int(on)), "ludwig_profile_enable")

    def profile_read(self):
        ms, n, cells = C.c_double(), C.c_int64(), C.c_int64()
        self._check(self.lib.ludwig_profile_read(self._h, C.byref(ms), C.byref(n), C.byref(cells)), "ludwig_profile_read")
        return ms.value, n.value, cells.value

    PROFILE_CLASSES = ("k1_plain", "k1_plain_ghost", "k1_feature", "k1_full", "interface_prepass", "bouzidi", "barrier", "level_step")

    def profile_classes(self) -> dict:
        """Device ms per launch class of the last profile_read()."""
        out = (C.c_double * 8)()
        self._check(self.lib.ludwig_profile_classes(self._h, out), "ludwig_profile_classes")
        return {k: v for k, v in zip(self.PROFILE_CLASSES, list(out)) if k != "-"}

    def profile_levels(self) -> list:
        """Device ms per (level, launch class) of the last profile_read(): one dict per level."""
        n = int(self.lib.ludwig_num_levels(self._h))
        names = self.PROFILE_CLASSES + ("halo_unpack", "halo_pack", "-", "-")
        out = (C.c_double * (12 * n))()
        self._check(self.lib.ludwig_profile_levels(self._h, out, 12 * n), "ludwig_profile_levels")
        return [{k: out[12 * l + i] for i, k in enumerate(names) if k != "-"} for l in range(n)]

    # -- multi-GPU (one process per GPU) ----------------------------------------------------
    def set_partition(self, rank: int, world: int):
        self._check(self.lib.ludwig_ctx_set_partition(self._h, rank, world), "ludwig_ctx_set_partition")
        self.rank, self.world = rank, world

    def set_barrier(self, fn):
        """fn(): cross-rank barrier called by the library after every level step (keep a reference to the thunk).
        An exception inside fn makes the callback return non-zero: the library marks the context failed and the next
        call raises LudwigError chained to that exception (ctypes itself would only print and swallow it)."""
        def thunk(_user):
            try:
                fn()
                return 0
            except BaseException as e:          # noqa: BLE001 - must not escape into C
                self._cb_error = e
                return -1
        self._barrier_thunk = BARRIER_CB(thunk)
        self._check(self.lib.ludwig_set_barrier_callback(self._h, self._barrier_thunk, None), "ludwig_set_barrier_callback")

    def clear_barrier(self):
        """Back to the library's own peer-flag barrier."""
        self._check(self.lib.ludwig_set_barrier_callback(self._h, C.cast(None, BARRIER_CB), None), "ludwig_set_barrier_callback")

    def local_blocks(self, level: int) -> np.ndarray:
        """0-based reference indices of the blocks this rank owns, in the library's internal order."""
        n = C.c_int32()
        self._check(self.lib.ludwig_level_local_blocks(self._h, level, C.byref(n), None), "ludwig_level_local_blocks")
        out = np.empty(n.value, np.int32)
        self._check(self.lib.ludwig_level_local_blocks(self._h, level, C.byref(n), _ptr(out)), "ludwig_level_local_blocks")
        return out - 1

    def upload_local(self, level: int, which: int, arr: np.ndarray):
        """arr: [ncomp, n_local, 8,8,8] for this rank's blocks in local_blocks(level) order."""
        a = _as(arr, np.float32)
        self._check(self.lib.ludwig_level_upload_local(self._h, level, which, _ptr(a)), "ludwig_level_upload_local")

    def download_local(self, level: int, which: int, n_local: int) -> np.ndarray:
        nc = _NCOMP[which]
        out = np.empty((n_local, 8, 8, 8) if nc == 1 else (nc, n_local, 8, 8, 8), np.float32)
        self._check(self.lib.ludwig_level_download_local(self._h, level, which, _ptr(out)), "ludwig_level_download_local")
        return out

    def ipc_export(self) -> bytes:
        need = C.c_int64()
        self._check(self.lib.ludwig_ipc_export(self._h, None, 0, C.byref(need)), "ludwig_ipc_export")
        buf = (C.c_ubyte * max(need.value, 1))()
        self._check(self.lib.ludwig_ipc_export(self._h, buf, need.value, C.byref(need)), "ludwig_ipc_export")
        return bytes(buf)[:need.value]

    def ipc_attach(self, all_handles: bytes, bytes_per_rank: int):
        buf = (C.c_ubyte * len(all_handles)).from_buffer_copy(all_handles)
        self._check(self.lib.ludwig_ipc_attach(self._h, buf, bytes_per_rank), "ludwig_ipc_attach")

    @staticmethod
    def attach_inprocess(contexts):
        """Peers living in this process: contexts[r] is rank r of len(contexts) (ludwig_attach_inprocess)."""
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        for c in contexts:
            c._check(c.lib.ludwig_attach_inprocess(c._h, arr, len(contexts)), "ludwig_attach_inprocess")

    # -- upload (main.jl:98,101,145) ---------------------------------------------------
    @staticmethod
    def make_desc(lv: BlockLevel):
        """ludwig_level_desc for a host BlockLevel + the arrays that must stay alive while it is used."""
        nb = lv.n_blocks
        bp = _as(lv.block_pointer, np.int32)
        nt = _as(lv.neighbor_table, np.int32)
        assert nt.shape == (27, nb), nt.shape
        coords = _as(lv.active_block_coords, np.int32)
        mx, my, mz = (_as(coords[:, i], np.int32) for i in range(3))
        obs = _as(lv.obstacle, np.uint8)
        sp = _as(lv.sponge, np.float32)
        wd = _as(lv.wall_dist, np.float32)
        assert obs.size == nb * 512 and sp.size == nb * 512 and wd.size == nb * 512
        d = LevelDesc()
        d.level_id, d.n_blocks = lv.level_id, nb
        d.dim_z, d.dim_y, d.dim_x = bp.shape
        d.tau, d.dx = float(lv.tau), float(lv.dx)
        d.block_pointer, d.neighbor_table = _ptr(bp), _ptr(nt)
        d.map_x, d.map_y, d.map_z = _ptr(mx), _ptr(my), _ptr(mz)
        d.obstacle, d.sponge, d.wall_dist = _ptr(obs), _ptr(sp), _ptr(wd)
        d.temporal_storage = int(lv.temporal_storage)
        keep = [bp, nt, mx, my, mz, obs, sp, wd]
        bz = bool(lv.bouzidi_enabled and lv.n_boundary_cells > 0 and lv.q_map is not None)
        d.bouzidi_enabled, d.n_boundary_cells = int(bz), int(lv.n_boundary_cells if bz else 0)
        if bz:
            q = _as(lv.q_map, np.float16)
            assert q.size == nb * 512 * 27
            cb, cx, cy, cz = _as(lv.cell_block, np.int32), _as(lv.cell_x, np.int8), _as(lv.cell_y, np.int8), _as(lv.cell_z, np.int8)
            d.q_map_f16, d.cell_block = _ptr(q), _ptr(cb)
            d.cell_x, d.cell_y, d.cell_z = _ptr(cx), _ptr(cy), _ptr(cz)
            keep += [q, cb, cx, cy, cz]
        return d, keep

    def add_level(self, lv: BlockLevel) -> int:
        d, keep = self.make_desc(lv)
        idx = C.c_int32(-1)
        self._check(self.lib.ludwig_level_create(self._h, C.byref(d), C.byref(idx)), "ludwig_level_create")
        del keep
        self.n_blocks.append(lv.n_blocks)     # blocks of the whole level (this rank's share: local_blocks())
        return idx.value

    def set_partition_plan(self, levels):
        """Spatially aligned, cost-balanced cut of all levels (call after set_partition, before add_level)."""
        descs, keeps = zip(*[self.make_desc(lv) for lv in levels])
        arr = (C.POINTER(LevelDesc) * len(descs))(*[C.pointer(d) for d in descs])
        keys = (C.c_uint64 * (self.world + 1))()
        rc = self.lib.ludwig_partition_plan(arr, len(descs), self.world, keys)
        if rc != 0:
            raise LudwigError(f"ludwig_partition_plan failed ({rc})")
        self._check(self.lib.ludwig_ctx_set_partition_keys(self._h, keys, len(descs)), "ludwig_ctx_set_partition_keys")
        del keeps
        return list(keys)

    def upload(self, level: int, which: int, arr: np.ndarray):
        nb = self.n_blocks[level]
        dt = np.uint8 if which == OBSTACLE else np.float32
        a = _as(arr, dt)
        assert a.size == nb * 512 * _NCOMP[which], (a.shape, nb, which)
        self._check(self.lib.ludwig_level_upload(self._h, level, which, _ptr(a)), "ludwig_level_upload")

    def download(self, level: int, which: int) -> np.ndarray:
        nb = self.n_blocks[level]
        nc = _NCOMP[which]
        dt = np.uint8 if which == OBSTACLE else np.float32
        shape = (nb, 8, 8, 8) if nc == 1 else (nc, nb, 8, 8, 8)
        out = np.empty(shape, dtype=dt)
        self._check(self.lib.ludwig_level_download(self._h, level, which, _ptr(out)), "ludwig_level_download")
        return out

    def create_mesh(self, centers: np.ndarray, normals: np.ndarray, areas: np.ndarray) -> C.c_void_p:
        """geometry.jl:60-84: Float64 geometry -> Float32 SoA."""
        arrs = [_as(centers[:, i], np.float32) for i in range(3)] + [_as(normals[:, i], np.float32) for i in range(3)]
        arrs.append(_as(areas, np.float32))
        h = C.c_void_p()
        self._check(self.lib.ludwig_mesh_create(self._h, len(areas), *[_ptr(a) for a in arrs], C.byref(h)), "ludwig_mesh_create")
        self._meshes.append(h)
        return h

    def create_forces(self, mesh, rho_ref: float, u_ref: float, area_ref: float, chord_ref: float,
                      moment_center: Sequence[float], symmetric: bool) -> C.c_void_p:
        mc = (C.c_double * 3)(*[float(v) for v in moment_center])
        h = C.c_void_p()
        self._check(self.lib.ludwig_forces_create(self._h, mesh, rho_ref, u_ref, area_ref, chord_ref, mc, int(symmetric), C.byref(h)),
                    "ludwig_forces_create")
        self._forces.append(h)
        return h

    # -- stepping (main.jl:126-135, solver_control.jl:145) -----------------------------------
    def init_equilibrium(self):
        self._check(self.lib.ludwig_init_equilibrium(self._h), "ludwig_init_equilibrium")

    def step_batch(self, t_start: int, batch: int, u_curr: float, params: Params):
        self._check(self.lib.ludwig_step_batch(self._h, t_start, batch, C.c_float(u_curr), C.byref(params)), "ludwig_step_batch")

    def level_step(self, level: int, t_sub: int, parent_t_sub: int, temporal_weight: float, u_curr: float, params: Params):
        self._check(self.lib.ludwig_level_step(self._h, level, t_sub, parent_t_sub, C.c_float(temporal_weight),
                                               C.c_float(u_curr), C.byref(params)), "ludwig_level_step")

    def snapshot_old(self, level: int, t_sub: int):
        self._check(self.lib.ludwig_level_snapshot_old(self._h, level, t_sub), "ludwig_level_snapshot_old")

    # -- diagnostics (main.jl:186,197) ---------------------------------------------------------
    AERO_KEYS = ("Fx", "Fy", "Fz", "Mx", "My", "Mz", "Fx_p", "Fy_p", "Fz_p", "Fx_v", "Fy_v", "Fz_v",
                 "Cd", "Cl", "Cs", "Cmx", "Cmy", "Cmz")

    def compute_aerodynamics(self, forces, level: int, mesh_offset: Sequence[float], velocity_scale: float,
                             rho_phys: float, search_radius: int = 5) -> dict:
        off = (C.c_double * 3)(*[float(v) for v in mesh_offset])
        out = (C.c_double * 18)()
        self._check(self.lib.ludwig_compute_aerodynamics(self._h, forces, level, off, velocity_scale, rho_phys,
                                                         search_radius, out), "ludwig_compute_aerodynamics")
        return dict(zip(self.AERO_KEYS, list(out)))

    def download_force_maps(self, forces, n_triangles: int):
        maps = [np.empty(n_triangles, np.float32) for _ in range(4)]
        self._check(self.lib.ludwig_forces_download_maps(self._h, forces, *[_ptr(m) for m in maps]), "ludwig_forces_download_maps")
        return maps

    STATS_KEYS = ("n_fluid", "rho_mean", "rho_min", "rho_max", "v_max", "kinetic_energy")

    def flow_stats(self, level: int = 0) -> dict:
        out = (C.c_double * 6)()
        self._check(self.lib.ludwig_flow_stats(self._h, level, out), "ludwig_flow_stats")
        return dict(zip(self.STATS_KEYS, list(out)))


def _valid_blocks(obj, n_levels):
    n = np.zeros(n_levels, np.int32)
    obj._check(getattr(obj.lib, obj._VALID)(obj._h, _ptr(n), None), obj._VALID)
    flat = np.empty(int(n.sum()), np.int32)
    obj._check(getattr(obj.lib, obj._VALID)(obj._h, _ptr(n), _ptr(flat)), obj._VALID)
    out, o = [], 0
    for c in n:
        out.append(flat[o:o + c] - 1); o += c
    return out


def _export(obj, t_step, n_levels):
    n = np.zeros(n_levels, np.int32)
    obj._check(getattr(obj.lib, obj._VALID)(obj._h, _ptr(n), None), obj._VALID)
    N = 512 * int(n.sum())
    rho = np.zeros(N, np.float32); vel = np.zeros((N, 3), np.float32); obs = np.zeros(N, np.uint8); lvl = np.zeros(N, np.int32)
    obj._check(getattr(obj.lib, obj._EXPORT)(obj._h, t_step, _ptr(rho), _ptr(vel), _ptr(obs), _ptr(lvl)), obj._EXPORT)
    return rho, vel, obs, lvl


class MultiContext:
    """ludwig_multi: N ranks (N GPUs, or N virtual ranks on one GPU) driven by this one thread.  Same call sequence as
    ``Context`` for what the kept Julia driver does (main.jl:54-249); ``create_forces`` takes the mesh arrays directly."""

    def __init__(self, n_ranks: int, devices: Optional[Sequence[int]] = None, lib_path: Optional[str] = None, options: Optional[dict] = None):
        self.lib = load_library(lib_path)
        self._h = C.c_void_p()
        dev = None if devices is None else _as(np.asarray(devices), np.int32)
        rc = self.lib.ludwig_multi_create(C.byref(self._h), n_ranks, _ptr(dev))
        if rc != 0:
            raise LudwigError(f"ludwig_multi_create failed ({rc})")
        self.world = n_ranks
        self.n_blocks: list[int] = []
        for k, v in (options or {}).items():
            self.set_option(k, v)

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.ludwig_multi_last_error(self._h)
            raise LudwigError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self._h:
            self.lib.ludwig_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value):
        if isinstance(value, bool):
            value = int(value)
        self._check(self.lib.ludwig_multi_set_option(self._h, str(key).encode(), str(value).encode()), f"ludwig_multi_set_option({key})")

    def rank_ctx(self, rank: int) -> "Context":
        """A non-owning Context view of one rank (per-rank calls: local_blocks, upload_local, profile_*)."""
        c = Context.__new__(Context)
        c.lib, c._h = self.lib, C.c_void_p(self.lib.ludwig_multi_ctx(self._h, rank))
        c.n_blocks, c.rank, c.world, c._meshes, c._forces, c._cb_error = self.n_blocks, rank, self.world, [], [], None
        c.close = lambda: None            # the ludwig_multi owns it
        return c

    def set_partition_plan(self, levels):
        descs, keeps = zip(*[Context.make_desc(lv) for lv in levels])
        arr = (C.POINTER(LevelDesc) * len(descs))(*[C.pointer(d) for d in descs])
        self._check(self.lib.ludwig_multi_set_partition_plan(self._h, arr, len(descs)), "ludwig_multi_set_partition_plan")
        del keeps

    def add_level(self, lv: BlockLevel) -> int:
        d, keep = Context.make_desc(lv)
        idx = C.c_int32(-1)
        self._check(self.lib.ludwig_multi_level_create(self._h, C.byref(d), C.byref(idx)), "ludwig_multi_level_create")
        del keep
        self.n_blocks.append(lv.n_blocks)
        return idx.value

    def upload(self, level: int, which: int, arr: np.ndarray):
        a = _as(arr, np.uint8 if which == OBSTACLE else np.float32)
        assert a.size == self.n_blocks[level] * 512 * _NCOMP[which]
        self._check(self.lib.ludwig_multi_level_upload(self._h, level, which, _ptr(a)), "ludwig_multi_level_upload")

    def download(self, level: int, which: int) -> np.ndarray:
        nb, nc = self.n_blocks[level], _NCOMP[which]
        out = np.empty((nb, 8, 8, 8) if nc == 1 else (nc, nb, 8, 8, 8), dtype=np.uint8 if which == OBSTACLE else np.float32)
        self._check(self.lib.ludwig_multi_level_download(self._h, level, which, _ptr(out)), "ludwig_multi_level_download")
        return out

    def init_equilibrium(self):
        self._check(self.lib.ludwig_multi_init_equilibrium(self._h), "ludwig_multi_init_equilibrium")

    def self_check(self) -> int:
        n = int(self.lib.ludwig_multi_self_check(self._h))
        if n < 0:
            self._check(n, "ludwig_multi_self_check")
        return n

    def init_uniform_flow(self, ux: float):
        self._check(self.lib.ludwig_multi_init_uniform_flow(self._h, C.c_float(ux)), "ludwig_multi_init_uniform_flow")

    def step_batch(self, t_start: int, batch: int, u_curr: float, params: Params):
        self._check(self.lib.ludwig_multi_step_batch(self._h, t_start, batch, C.c_float(u_curr), C.byref(params)), "ludwig_multi_step_batch")

    def sync(self):
        self._check(self.lib.ludwig_multi_sync(self._h), "ludwig_multi_sync")

    def flow_stats(self, level: int = 0) -> dict:
        out = (C.c_double * 6)()
        self._check(self.lib.ludwig_multi_flow_stats(self._h, level, out), "ludwig_multi_flow_stats")
        return dict(zip(Context.STATS_KEYS, list(out)))

    def create_forces(self, centers, normals, areas, rho_ref, u_ref, area_ref, chord_ref, moment_center, symmetric) -> int:
        arrs = [_as(centers[:, i], np.float32) for i in range(3)] + [_as(normals[:, i], np.float32) for i in range(3)] + [_as(areas, np.float32)]
        mc = (C.c_double * 3)(*[float(v) for v in moment_center])
        h = C.c_int32(-1)
        self._check(self.lib.ludwig_multi_forces_create(self._h, len(areas), *[_ptr(a) for a in arrs], rho_ref, u_ref, area_ref, chord_ref,
                                                        mc, int(symmetric), C.byref(h)), "ludwig_multi_forces_create")
        return h.value

    def compute_aerodynamics(self, handle: int, level: int, mesh_offset, velocity_scale: float, rho_phys: float, search_radius: int = 5) -> dict:
        off = (C.c_double * 3)(*[float(v) for v in mesh_offset])
        out = (C.c_double * 18)()
        self._check(self.lib.ludwig_multi_compute_aerodynamics(self._h, handle, level, off, velocity_scale, rho_phys, search_radius, out),
                    "ludwig_multi_compute_aerodynamics")
        return dict(zip(Context.AERO_KEYS, list(out)))

    def download_force_maps(self, handle: int, n_triangles: int):
        maps = [np.empty(n_triangles, np.float32) for _ in range(4)]
        self._check(self.lib.ludwig_multi_forces_download_maps(self._h, handle, *[_ptr(m) for m in maps]), "ludwig_multi_forces_download_maps")
        return maps

    def device_bytes(self) -> int:
        return int(self.lib.ludwig_multi_device_bytes(self._h))

    _VALID, _EXPORT = "ludwig_multi_output_valid_blocks", "ludwig_multi_output_export"

    def output_valid_blocks(self):
        return _valid_blocks(self, len(self.n_blocks))

    def output_export(self, t_step: int):
        return _export(self, t_step, len(self.n_blocks))
