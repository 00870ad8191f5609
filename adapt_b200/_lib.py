"""ctypes binding of libadapt_b200.so (include/adapt_b200.h) and the scene packer.

``pack_scene`` flattens the 4-tuple produced by ``scene_parsing`` into the C struct
``adapt_scene_desc`` -- the job ``PathTracer.__init__`` / ``load_primitives`` / ``initialze`` do with
Taichi fields in the reference (tracer/path_tracer.py:54-141,245-274; tracer/tracer_base.py:36-134).

There is deliberately no CPU fallback: ``load_library`` raises if the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np

from .bxdf.brdf import BXDF_DTYPE
from .emitters.abtract_source import EMITTER_DTYPE
from .la.cam_transform import fov2focal, np_rotation_between

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libadapt_b200.so")


class AdaptError(RuntimeError):
    pass


class adapt_bxdf(C.Structure):
    _fields_ = [("kind", C.c_int32), ("type", C.c_int32), ("is_delta", C.c_int32),
                ("k_d", C.c_float * 3), ("k_s", C.c_float * 3), ("k_g", C.c_float * 3), ("mean", C.c_float * 3),
                ("ior", C.c_float)]


class adapt_emitter(C.Structure):
    _fields_ = [("type", C.c_int32), ("obj_ref_id", C.c_int32), ("bool_bits", C.c_int32),
                ("intensity", C.c_float * 3), ("dir", C.c_float * 3), ("pos", C.c_float * 3),
                ("inv_area", C.c_float), ("r", C.c_float), ("emit_time", C.c_float), ("_pad", C.c_float)]


class adapt_medium(C.Structure):
    """include/adapt_b200.h: adapt_medium (reference bxdf/medium.py:71-78 + bxdf/phase.py:30-34)."""
    _fields_ = [("type", C.c_int32), ("ior", C.c_float), ("u_a", C.c_float * 3), ("u_s", C.c_float * 3), ("u_e", C.c_float * 3),
                ("par", C.c_float * 3), ("pdf", C.c_float * 3), ("_pad", C.c_int32 * 3)]


MEDIUM_DTYPE = np.dtype([("type", "<i4"), ("ior", "<f4"), ("u_a", "<f4", 3), ("u_s", "<f4", 3), ("u_e", "<f4", 3), ("par", "<f4", 3),
                         ("pdf", "<f4", 3), ("_pad", "<i4", 3)])
assert C.sizeof(adapt_medium) == MEDIUM_DTYPE.itemsize == 80


def medium_record(m) -> np.ndarray:
    """Medium_np (or None = transparent) -> one adapt_medium record."""
    rec = np.zeros((), dtype=MEDIUM_DTYPE)
    rec["type"], rec["ior"], rec["pdf"] = -1, 1.0, (1.0, 0.0, 0.0)
    if m is not None:
        rec["type"], rec["ior"] = m.type_id, m.ior
        for k in ("u_a", "u_s", "u_e", "par", "pdf"):
            rec[k] = np.broadcast_to(np.float32(getattr(m, k)), 3)
    return rec


assert C.sizeof(adapt_bxdf) == BXDF_DTYPE.itemsize == 64
assert C.sizeof(adapt_emitter) == EMITTER_DTYPE.itemsize == 64

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


class adapt_scene_desc(C.Structure):
    _fields_ = [
        ("n_prims", C.c_int32), ("n_objects", C.c_int32),
        ("primitives", _fp), ("n_g", _fp), ("n_s", _fp), ("uvs", _fp),
        ("obj_info", _ip), ("obj_aabb", _fp), ("emitter_id", _ip),
        ("bxdfs", C.POINTER(adapt_bxdf)),
        ("n_emitters", C.c_int32),
        ("emitters", C.POINTER(adapt_emitter)),
        ("width", C.c_int32), ("height", C.c_int32),
        ("cam_r", C.c_float * 9), ("cam_t", C.c_float * 3),
        ("inv_focal", C.c_float), ("half_w", C.c_float), ("half_h", C.c_float),
        ("do_crop", C.c_int32), ("start_x", C.c_int32), ("end_x", C.c_int32), ("start_y", C.c_int32), ("end_y", C.c_int32),
        ("max_bounce", C.c_int32), ("num_shadow_ray", C.c_int32), ("use_rr", C.c_int32), ("rr_bounce_th", C.c_int32),
        ("use_mis", C.c_int32), ("anti_alias", C.c_int32), ("stratified_sampling", C.c_int32),
        ("brdf_two_sides", C.c_int32), ("has_v_normal", C.c_int32),
        ("rr_threshold", C.c_float), ("world_ior", C.c_float),
        ("seed", C.c_uint64),
        ("device_id", C.c_int32), ("n_pixels", C.c_int32),
        ("pixel_list", _ip),
        ("pool_size", C.c_int32),
        ("accelerator", C.c_int32),
        ("bvh_builder", C.c_int32),
        ("reserved", C.c_int32 * 5),
        ("textures", C.c_void_p),
        ("tex_image", _fp * 3),
        ("tex_size", C.c_int32 * 3),
        ("integrator", C.c_int32),
        ("media", C.POINTER(adapt_medium)),
        ("n_devices", C.c_int32),
        ("device_ids", _ip),
    ]


class adapt_stats(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
        ("iterations", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("ms_logic", C.c_float), ("ms_closest", C.c_float), ("ms_shadow", C.c_float), ("ms_total", C.c_float),
        ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64),
        ("reserved", C.c_uint64 * 4),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


# every symbol include/adapt_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "adapt_create", "adapt_destroy", "adapt_render", "adapt_wait_enqueued", "adapt_sync", "adapt_read_accum", "adapt_load_accum",
    "adapt_read_pixels", "adapt_host_alloc", "adapt_host_free",
    "adapt_accum_device_ptr", "adapt_set_stream", "adapt_get_stats", "adapt_reset_stats", "adapt_intersect_batch", "adapt_bxdf_batch",
    "adapt_bvh_export", "adapt_bvh_export_wide", "adapt_update_geometry", "adapt_refit_geometry",
    "adapt_bvh_build", "adapt_free", "adapt_last_error", "adapt_version", "adapt_tile_partition",
]

_lib = None


def load_library(path: Optional[str] = None):
    """dlopen libadapt_b200.so and declare prototypes.  Raises AdaptError when it has not been built
    (run ``python -m adapt_b200.build`` or ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("ADAPT_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise AdaptError(f"{path} not found: the CUDA library is not built and there is no CPU fallback. "
                         "Run `python -m adapt_b200.build`.")
    lib = C.CDLL(path)
    H = C.c_void_p
    lib.adapt_create.argtypes = [C.POINTER(H), C.POINTER(adapt_scene_desc)]
    lib.adapt_create.restype = C.c_int
    lib.adapt_destroy.argtypes = [H]
    lib.adapt_destroy.restype = None
    lib.adapt_render.argtypes = [H, C.c_int32]
    lib.adapt_render.restype = C.c_int
    lib.adapt_sync.argtypes = [H]
    lib.adapt_sync.restype = C.c_int
    lib.adapt_read_accum.argtypes = [H, _fp, _ip]
    lib.adapt_read_accum.restype = C.c_int
    lib.adapt_tile_partition.argtypes = [C.c_int32] * 5 + [_ip, _ip, C.c_int32]
    lib.adapt_tile_partition.restype = C.c_int32
    lib.adapt_wait_enqueued.argtypes = [H]
    lib.adapt_wait_enqueued.restype = C.c_int
    lib.adapt_load_accum.argtypes = [H, _fp, C.c_int32]
    lib.adapt_load_accum.restype = C.c_int
    lib.adapt_read_pixels.argtypes = [H, _fp, _ip]
    lib.adapt_read_pixels.restype = C.c_int
    lib.adapt_host_alloc.argtypes = [C.c_uint64]
    lib.adapt_host_alloc.restype = C.c_void_p
    lib.adapt_host_free.argtypes = [C.c_void_p]
    lib.adapt_host_free.restype = None
    lib.adapt_accum_device_ptr.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.adapt_accum_device_ptr.restype = C.c_int
    lib.adapt_set_stream.argtypes = [H, C.c_void_p]
    lib.adapt_set_stream.restype = C.c_int
    lib.adapt_get_stats.argtypes = [H, C.POINTER(adapt_stats)]
    lib.adapt_get_stats.restype = C.c_int
    lib.adapt_reset_stats.argtypes = [H]
    lib.adapt_reset_stats.restype = C.c_int
    lib.adapt_intersect_batch.argtypes = [H, _fp, _fp, _fp, C.c_int32, C.c_int32, _ip, _ip, _fp, _fp, _fp]
    lib.adapt_intersect_batch.restype = C.c_int
    lib.adapt_bxdf_batch.argtypes = [H, C.c_int32, C.c_int32, _fp, _fp, _fp, _fp, C.c_int32, C.c_uint64, _fp, _fp, _fp, _fp, _fp, _ip]
    lib.adapt_bxdf_batch.restype = C.c_int
    lib.adapt_bvh_export.argtypes = [H, _ip, _ip, _ip, _ip, _fp, _fp, _fp]
    lib.adapt_bvh_export.restype = C.c_int
    lib.adapt_bvh_export_wide.argtypes = [H, _ip, _ip, C.POINTER(C.c_uint32)]
    lib.adapt_bvh_export_wide.restype = C.c_int
    lib.adapt_update_geometry.argtypes = [H, _fp, _fp, _fp]
    lib.adapt_update_geometry.restype = C.c_int
    lib.adapt_refit_geometry.argtypes = [H, _fp, _fp, _fp]
    lib.adapt_refit_geometry.restype = C.c_int
    lib.adapt_bvh_build.argtypes = [_fp, C.c_int32, _ip, C.c_int32, _fp, _fp,
                                    C.POINTER(_fp), C.POINTER(_fp), C.POINTER(_ip), C.POINTER(_ip), _ip, _ip]
    lib.adapt_bvh_build.restype = C.c_int
    lib.adapt_free.argtypes = [C.c_void_p]
    lib.adapt_free.restype = None
    lib.adapt_last_error.argtypes = []
    lib.adapt_last_error.restype = C.c_char_p
    lib.adapt_version.argtypes = []
    lib.adapt_version.restype = C.c_char_p
    if path == os.environ.get("ADAPT_B200_LIB", LIB_PATH):
        _lib = lib
    return lib


def check(lib, status: int, what: str = ""):
    if status != 0:
        msg = lib.adapt_last_error()
        raise AdaptError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class PackedScene:
    """adapt_scene_desc + the numpy arrays it points into (kept alive here)."""

    def __init__(self):
        self.desc = adapt_scene_desc()
        self.keep = {}
        self.host = {}      # host-side scalars the Renderer class exposes (focal, cam_orient, ...)

    def _ptr(self, name, arr, ctype):
        self.keep[name] = arr
        return arr.ctypes.data_as(C.POINTER(ctype))


def pack_scene(emitters: List, array_info: dict, objects: List, prop: dict, seed: int = 0, device_id: int = 0,
               pixel_list: Optional[np.ndarray] = None, pool_size: int = 0, max_bounce: Optional[int] = None,
               bvh_builder=0, integrator="pt", device_ids=None) -> PackedScene:
    ps = PackedScene()
    d = ps.desc
    film = prop["film"]
    w, h = int(film["width"]), int(film["height"])
    # ---- camera / film (tracer_base.py:36-75) ----
    crop_x, crop_y = film.get("crop_x", 0), film.get("crop_y", 0)
    crop_rx, crop_ry = film.get("crop_rx", 0), film.get("crop_ry", 0)
    do_crop = (crop_rx > 0) and (crop_ry > 0)
    if do_crop:
        sx, ex, sy, ey = crop_x - crop_rx, crop_x + crop_rx, crop_y - crop_ry, crop_y + crop_ry
    else:
        sx, sy, ex, ey = 0, 0, w, h
    focal = fov2focal(prop["fov"], min(w, h))
    cam_orient = np.array(prop["transform"][0], dtype=np.float64)
    cam_orient = cam_orient / np.linalg.norm(cam_orient)
    cam_t = np.float32(prop["transform"][1])
    cam_r = np.float32(np_rotation_between(np.float32([0, 0, 1]), cam_orient))
    ps.host.update(dict(w=w, h=h, crop_x=crop_x, crop_y=crop_y, crop_rx=crop_rx, crop_ry=crop_ry, do_crop=do_crop,
                        start_x=sx, end_x=ex, start_y=sy, end_y=ey, focal=focal, cam_orient=cam_orient, cam_t=cam_t,
                        cam_r=cam_r))
    d.width, d.height = w, h
    d.cam_r = (C.c_float * 9)(*cam_r.reshape(-1).tolist())
    d.cam_t = (C.c_float * 3)(*cam_t.tolist())
    d.inv_focal = float(np.float32(1.0 / focal))
    d.half_w, d.half_h = w / 2, h / 2
    d.do_crop, d.start_x, d.end_x, d.start_y, d.end_y = int(do_crop), sx, ex, sy, ey
    # ---- integrator flags (path_tracer.py:63-69) ----
    d.max_bounce = int(prop["max_bounce"] if max_bounce is None or max_bounce < 0 else max_bounce)
    d.num_shadow_ray = int(prop["num_shadow_ray"])
    d.use_rr = int(bool(prop["use_rr"]))
    d.rr_threshold = float(prop.get("rr_threshold", 0.1))
    d.rr_bounce_th = int(prop.get("rr_bounce_th", 4))
    d.use_mis = int(bool(prop["use_mis"]))
    d.anti_alias = int(bool(prop["anti_alias"]))
    d.stratified_sampling = int(bool(prop["stratified_sampling"]))
    d.brdf_two_sides = int(bool(prop.get("brdf_two_sides", False)))
    d.has_v_normal = int(bool(prop["has_vertex_normal"]))
    world = prop.get("world", None)
    d.world_ior = float(world.medium.ior) if world is not None else 1.0
    # ---- geometry (tracer_base.py:117-134) ----
    prims = _f32(array_info["primitives"]).reshape(-1, 9)
    n_prims = prims.shape[0]
    d.n_prims = n_prims
    d.n_objects = len(objects)
    d.primitives = ps._ptr("primitives", prims, C.c_float)
    d.n_g = ps._ptr("n_g", _f32(array_info["n_g"]).reshape(-1, 3), C.c_float)
    if d.has_v_normal:
        d.n_s = ps._ptr("n_s", _f32(array_info["n_s"]).reshape(-1, 9), C.c_float)
    else:
        d.n_s = None
    d.uvs = ps._ptr("uvs", _f32(array_info["uvs"]).reshape(-1, 6), C.c_float)
    # ---- per-object tables (path_tracer.py:245-274) ----
    obj_info = np.zeros((len(objects), 3), np.int32)
    aabbs = np.zeros((len(objects), 6), np.float32)
    emitter_id = np.full((len(objects),), -1, np.int32)
    bx = np.zeros((len(objects),), dtype=BXDF_DTYPE)
    em = np.zeros((max(len(emitters), 1),), dtype=EMITTER_DTYPE)
    for i, e in enumerate(emitters):
        em[i] = e.export()
        em[i]["obj_ref_id"] = -1
    acc = 0
    for i, obj in enumerate(objects):
        obj_info[i] = (acc, obj.tri_num, obj.type)
        acc += obj.tri_num
        bx[i] = obj.bsdf.export()
        aabbs[i, :3] = obj.aabb[0]
        aabbs[i, 3:] = obj.aabb[1]
        emitter_id[i] = obj.emitter_ref_id
        if obj.emitter_ref_id >= 0:
            em[obj.emitter_ref_id]["obj_ref_id"] = i
    if acc != n_prims:
        raise ValueError(f"objects hold {acc} primitives but array_info has {n_prims}")
    d.obj_info = ps._ptr("obj_info", _i32(obj_info), C.c_int32)
    d.obj_aabb = ps._ptr("obj_aabb", aabbs, C.c_float)
    d.emitter_id = ps._ptr("emitter_id", emitter_id, C.c_int32)
    ps.keep["bxdfs"] = bx
    d.bxdfs = C.cast(bx.ctypes.data, C.POINTER(adapt_bxdf))
    d.n_emitters = len(emitters)
    ps.keep["emitters"] = em
    d.emitters = C.cast(em.ctypes.data, C.POINTER(adapt_emitter))
    # ---- textures (path_tracer.py:83-123, 262-266): per-object descriptors for the albedo / normal / bump maps + atlases ----
    packed = prop.get("packed_textures", None)
    d.textures = None
    if packed is not None and any(packed.get(k) is not None for k in ("albedo", "normal", "bump")):
        from .bxdf.texture import TEXTURE_DTYPE, Texture_np
        tx = np.zeros((3, len(objects)), dtype=TEXTURE_DTYPE)
        for m, key in enumerate(("albedo", "normal", "bump")):
            img = packed.get(key)
            for i, obj in enumerate(objects):
                t = (obj.texture_group or {}).get(key) if img is not None else None
                tx[m, i] = t.export() if t is not None else Texture_np.default()
            if img is not None:
                img = _f32(img)
                if img.ndim != 3 or img.shape[0] != img.shape[1] or img.shape[2] != 3:
                    raise ValueError(f"packed '{key}' texture must be a square (size, size, 3) image")
                d.tex_image[m] = ps._ptr("tex_" + key, img, C.c_float)
                d.tex_size[m] = img.shape[0]
        ps.keep["textures"] = tx
        d.textures = tx.ctypes.data
    # ---- participating media (renderer/vpt.py): the medium attached to every BSDF object + the world's free-space medium ----
    if integrator not in ("pt", "vpt"):
        raise NotImplementedError(f"integrator '{integrator}': this path covers 'pt' (and 'vpt' in the CPU oracle)")
    d.integrator = 1 if integrator == "vpt" else 0
    if prop.get("volume"):
        if integrator == "vpt":
            raise NotImplementedError("grid volumes (<volume>, bxdf/volume.py) are outside the homogeneous-media slice of vpt")
    md = np.zeros(len(objects) + 1, dtype=MEDIUM_DTYPE)
    for i, obj in enumerate(objects):
        md[i] = medium_record(getattr(obj.bsdf, "medium", None))
    md[len(objects)] = medium_record(world.medium if world is not None else None)
    ps.keep["media"] = md
    d.media = C.cast(md.ctypes.data, C.POINTER(adapt_medium))
    # ---- back-end knobs ----
    d.seed = int(seed)
    d.device_id = int(device_id)
    # several GPUs behind one handle: the library splits the film into interleaved tiles itself (include/adapt_b200.h: n_devices)
    if device_ids is not None and len(device_ids) > 1:
        if pixel_list is not None:
            raise ValueError("device_ids and pixel_list are exclusive: a multi-device handle partitions the film itself")
        ids = _i32(np.asarray(list(device_ids))).reshape(-1)
        d.n_devices = ids.shape[0]
        d.device_ids = ps._ptr("device_ids", ids, C.c_int32)
        d.device_id = int(ids[0])
    else:
        d.n_devices = 0
        d.device_ids = None
        if device_ids is not None and len(device_ids) == 1:
            d.device_id = int(device_ids[0])
    if pixel_list is not None:
        pl = _i32(pixel_list).reshape(-1)
        if pl.shape[0] == 0:
            # n_pixels == 0 means "whole film" to adapt_create: an empty list must never turn into that (a rank that owns no tile would
            # render everything and the framebuffer reduce would count those pixels several times)
            raise ValueError("pixel_list is empty: this handle would own no pixel (film with fewer tiles than ranks? use a smaller tile)")
        d.n_pixels = pl.shape[0]
        d.pixel_list = ps._ptr("pixel_list", pl, C.c_int32)
    else:
        d.n_pixels = 0
        d.pixel_list = None
    d.pool_size = int(pool_size)
    # the reference's accelerator switch. The CUDA path always uses its BVH; the oracle
    # follows the reference and picks brute force unless the XML asked for "bvh".
    d.accelerator = 1 if prop.get("accelerator", "none") == "bvh" else 0
    d.bvh_builder = {"default": 0, "lbvh": 1, "sah_device": 2, "device": 2, "sah": 3, "host": 3}[bvh_builder] if isinstance(bvh_builder, str) else int(bvh_builder)
    ps.host.update(dict(num_objects=len(objects), num_prims=n_prims, src_num=len(emitters)))
    return ps
