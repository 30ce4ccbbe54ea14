"""STL reader + per-triangle geometry: host-side mirror of ``geometry.jl`` (kept Julia driver file).

parse_binary_stl :116-135 (Float32 -> Float64 x scale), parse_ascii_stl :137-158, load_mesh :160-209 (format
sniffing, bounds), compute_geometry_properties :86-114 (normals / areas / centres in Float64).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class SolverMesh:
    triangles: np.ndarray   # float64 [n,3,3]  (triangle, vertex, xyz) in STL coordinates x scale
    min_bounds: tuple
    max_bounds: tuple
    normals: np.ndarray     # float64 [n,3]
    areas: np.ndarray       # float64 [n]
    centers: np.ndarray     # float64 [n,3]

    @property
    def n_triangles(self) -> int:
        return int(self.triangles.shape[0])


def _parse_binary(path: str, scale: float) -> np.ndarray:
    with open(path, "rb") as fh:
        fh.seek(80)
        count = int(np.frombuffer(fh.read(4), dtype="<u4")[0])
        rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("attr", "<u2")])
        data = np.frombuffer(fh.read(count * 50), dtype=rec, count=count)
    return data["v"].astype(np.float64) * scale


def _parse_ascii(path: str, scale: float) -> np.ndarray:
    tris, cur = [], []
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            s = line.strip()
            if s.startswith("vertex"):
                parts = s.split()
                if len(parts) >= 4:
                    cur.append((float(parts[1]) * scale, float(parts[2]) * scale, float(parts[3]) * scale))
            elif s.startswith("endloop"):
                if len(cur) == 3:
                    tris.append(cur)
                cur = []
    return np.array(tris, dtype=np.float64).reshape(-1, 3, 3)


def compute_geometry_properties(tri: np.ndarray):
    """geometry.jl:86-114."""
    v1, v2, v3 = tri[:, 0], tri[:, 1], tri[:, 2]
    e1, e2 = v2 - v1, v3 - v1
    cp = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1],
                   e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                   e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], axis=1)
    nrm = np.sqrt(cp[:, 0] * cp[:, 0] + cp[:, 1] * cp[:, 1] + cp[:, 2] * cp[:, 2])
    area = 0.5 * nrm
    with np.errstate(divide="ignore", invalid="ignore"):
        normal = np.where((area > 1e-12)[:, None], cp / (2.0 * area)[:, None], 0.0)
    centers = ((v1 + v2) + v3) / 3.0
    return normal, area, centers


def load_mesh(path: str, scale: float = 1.0) -> SolverMesh:
    """geometry.jl:160-209."""
    if not os.path.isfile(path):
        raise FileNotFoundError(f"STL file not found: {path}")
    size = os.path.getsize(path)
    is_binary = True
    if size < 84:
        is_binary = False
    else:
        with open(path, "rb") as fh:
            header = fh.read(5)
            if header.lower().startswith(b"solid"):
                fh.seek(80)
                count = int(np.frombuffer(fh.read(4), dtype="<u4")[0])
                if size != 84 + count * 50:
                    is_binary = False
    tri = _parse_binary(path, scale) if is_binary else _parse_ascii(path, scale)
    if tri.shape[0] == 0:
        raise ValueError("No triangles loaded.")
    pts = tri.reshape(-1, 3)
    mn, mx = pts.min(axis=0), pts.max(axis=0)
    normals, areas, centers = compute_geometry_properties(tri)
    return SolverMesh(np.ascontiguousarray(tri), tuple(float(v) for v in mn), tuple(float(v) for v in mx), normals, areas, centers)
