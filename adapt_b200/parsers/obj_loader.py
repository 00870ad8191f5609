"""Wavefront OBJ ingestion without pywavefront.

The reference (``parsers/obj_loader.py:21-80``) reads the *first material's* interleaved vertex
buffer produced by ``pywavefront.Wavefront(path, collect_faces=True)`` and slices it by the
material's ``vertex_format`` (``[T2F_][C3F_][N3F_]V3F``).  ``read_obj_interleaved`` below rebuilds
exactly that buffer: one record per face corner in file order, n-gons fan-triangulated, a default
material synthesised when a face precedes any ``usemtl``.  Everything after that
(``extract_obj_info``, ``calculate_surface_area``, ``apply_transform``) follows the reference
semantics, including its quirks: rotation about the mesh centroid by right multiplication, vertex
normals not rotated, scale parsed but never applied (obj_loader.py:100-122).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import numpy as np

from ..utils.tools import CONSOLE

__all__ = ["extract_obj_info", "apply_transform", "calculate_surface_area", "read_obj_interleaved",
           "TRIANGLE_MESH", "SPHERE"]

TRIANGLE_MESH = 0
SPHERE = 1


def _corner_index(tok: str, n: int) -> int:
    """OBJ indices are 1-based; negative values count from the end."""
    i = int(tok)
    return i - 1 if i > 0 else n + i


def read_obj_interleaved(path: str) -> Tuple[str, np.ndarray]:
    """Return (vertex_format, float32 buffer (n_corners, dim)) of the first material in the file.

    Fast path: a pure-triangle mesh whose faces all share one corner layout is decoded with numpy.
    """
    pos: List[List[float]] = []
    col: List[List[float]] = []
    tex: List[List[float]] = []
    nrm: List[List[float]] = []
    # material name -> list of face-corner token lists; insertion order == pywavefront's dict order
    faces: Dict[str, List[List[str]]] = {}
    active: Optional[str] = None
    base_dir = os.path.dirname(path)

    with open(path, "r", errors="ignore") as fh:
        for line in fh:
            if not line or line[0] == "#":
                continue
            parts = line.split()
            if not parts:
                continue
            key = parts[0]
            if key == "v":
                pos.append([float(parts[1]), float(parts[2]), float(parts[3])])
                if len(parts) >= 7:
                    col.append([float(parts[4]), float(parts[5]), float(parts[6])])
            elif key == "vn":
                nrm.append([float(parts[1]), float(parts[2]), float(parts[3])])
            elif key == "vt":
                tex.append([float(parts[1]), float(parts[2]) if len(parts) > 2 else 0.0])
            elif key == "f":
                if active is None:
                    active = "default0"
                faces.setdefault(active, []).append(parts[1:])
            elif key == "usemtl":
                active = parts[1] if len(parts) > 1 else "default0"
                faces.setdefault(active, [])
            elif key == "mtllib" and len(parts) > 1:
                mtl_path = os.path.join(base_dir, parts[1])
                if os.path.exists(mtl_path):
                    with open(mtl_path, "r", errors="ignore") as mf:
                        for ml in mf:
                            mp = ml.split()
                            if len(mp) >= 2 and mp[0] == "newmtl":
                                faces.setdefault(mp[1], [])

    first = None
    for name, flist in faces.items():
        first = flist
        break
    if first is None or len(first) == 0:
        raise ValueError("This wavefront onject file has no material but it is required.")

    # vertex layout is decided by the first corner of the first face (pywavefront behaviour)
    probe = first[0][0].split("/")
    has_vt = len(probe) >= 2 and probe[1] != ""
    has_vn = len(probe) == 3 and probe[2] != ""
    has_col = len(col) == len(pos) and len(col) > 0
    fmt_parts = []
    if has_vt:
        fmt_parts.append("T2F")
    if has_col:
        fmt_parts.append("C3F")
    if has_vn:
        fmt_parts.append("N3F")
    fmt_parts.append("V3F")
    vertex_format = "_".join(fmt_parts)

    P = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
    C = np.asarray(col, dtype=np.float64).reshape(-1, 3) if has_col else None
    T = np.asarray(tex, dtype=np.float64).reshape(-1, 2) if has_vt else None
    N = np.asarray(nrm, dtype=np.float64).reshape(-1, 3) if has_vn else None

    # fan triangulation: (c0, c_{k-1}, c_k); identity for triangles
    tri_corners: List[str] = []
    for f in first:
        if len(f) == 3:
            tri_corners.extend(f)
        else:
            for k in range(2, len(f)):
                tri_corners.extend((f[0], f[k - 1], f[k]))
    ncorner = len(tri_corners)
    vi = np.empty(ncorner, dtype=np.int64)
    ti = np.empty(ncorner, dtype=np.int64) if has_vt else None
    ni = np.empty(ncorner, dtype=np.int64) if has_vn else None
    for k, tok in enumerate(tri_corners):
        sp = tok.split("/")
        vi[k] = _corner_index(sp[0], len(P))
        if has_vt:
            ti[k] = _corner_index(sp[1], len(T))
        if has_vn:
            ni[k] = _corner_index(sp[2], len(N))
    cols = []
    if has_vt:
        cols.append(T[ti])
    if has_col:
        cols.append(C[vi])
    if has_vn:
        cols.append(N[ni])
    cols.append(P[vi])
    return vertex_format, np.concatenate(cols, axis=1).astype(np.float32)


def extract_obj_info(path: str, verbose: bool = True, auto_scale_uv: bool = False):
    """-> (mesh_faces (N,3,3), normals (N,3), vert_normal (N,3,3)|None, uv_coords (N,3,2)|None).

    Same slicing and geometric-normal formula as obj_loader.py:40-78 (cross(v1-v0, v2-v1), normalised).
    """
    vert_type, all_data = read_obj_interleaved(path)
    if "T" not in vert_type and verbose:
        CONSOLE.log(f"[blue]Attention: Object contains no uv-coordinates for vtype '{vert_type}'")
    all_parts = vert_type[:-1].split("F_")                                 # "N3F_V3F" -> ["N3", "V3"]
    start_dim = 0
    mesh_faces = None
    vert_normal = None
    uv_coords = None
    for part in all_parts:
        width = int(part[1:])
        if part.startswith("T"):
            uv_coords = all_data[:, start_dim:start_dim + 2]
            if auto_scale_uv:
                lo, hi = uv_coords.min(), uv_coords.max()
                uv_coords = (uv_coords - lo) / (hi - lo)
            uv_coords = uv_coords.reshape(-1, 3, 2)
        elif part.startswith("V"):
            mesh_faces = np.float32(all_data[:, start_dim:start_dim + 3]).reshape(-1, 3, 3)
        elif part.startswith("N"):
            vert_normal = np.float32(all_data[:, start_dim:start_dim + 3]).reshape(-1, 3, 3)
        start_dim += width
    assert mesh_faces is not None
    dp1 = mesh_faces[:, 1, :] - mesh_faces[:, 0, :]
    dp2 = mesh_faces[:, 2, :] - mesh_faces[:, 1, :]
    normals = np.cross(dp1, dp2)
    normals /= np.linalg.norm(normals, axis=-1, keepdims=True)
    if verbose:
        CONSOLE.log(f"Mesh loaded from '{path}', output shape: [blue]{mesh_faces.shape}[/blue]")
    return mesh_faces, normals, vert_normal, uv_coords


def calculate_surface_area(meshes: np.ndarray, _type: int = 0) -> float:
    """Total triangle area, or 4*pi*r^2 for a sphere (obj_loader.py:82-93)."""
    if _type == TRIANGLE_MESH:
        dv1 = meshes[:, 1] - meshes[:, 0]
        dv2 = meshes[:, 2] - meshes[:, 0]
        return float((np.linalg.norm(np.cross(dv1, dv2), axis=-1) / 2.0).sum(dtype=np.float64))
    if _type == SPHERE:
        radius = meshes[0, 1, 0]
        return float(4.0 * np.pi * radius ** 2)
    return 0.0


def is_uniform_scaling(scale: np.ndarray) -> bool:
    return not (scale[0] != scale[1] or scale[0] != scale[2])


def apply_transform(meshes, normals, trans_r, trans_t, trans_s):
    """Reference transform semantics (obj_loader.py:100-122): scale is *never applied*."""
    if trans_s is not None and not is_uniform_scaling(trans_s):
        CONSOLE.log("Warning: scaling for meshes should be uniform, otherwise normals should be re-computed.")
        trans_s[1] = trans_s[0]
        trans_s[2] = trans_s[0]
    if trans_r is not None:
        center = meshes.mean(axis=1).mean(axis=0)
        meshes = meshes - center
        meshes = meshes @ trans_r
        if normals is not None:
            normals = normals @ trans_r
        meshes = meshes + center
    if trans_t is not None:
        meshes = meshes + trans_t
    return meshes, normals
