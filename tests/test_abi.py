"""C-ABI library checks that need no GPU: it loads, exports every symbol include/adapt_b200.h declares,
refuses to run without a device (no CPU fallback), and its host-side BVH builder (the drop-in for the
reference's bvh_cpp.bvh_build) returns a valid tree in the reference's 4-array layout."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_scene
from adapt_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    from adapt_b200.build import build
    build()
    return _lib.load_library()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "adapt_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(adapt_[a-z_]+)\s*\(", header)))
    assert declared == sorted(_lib.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.adapt_version()


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.adapt_bxdf) == 64 and C.sizeof(_lib.adapt_emitter) == 64 and C.sizeof(_lib.adapt_medium) == 80
    assert _lib.adapt_scene_desc.media.offset % 8 == 0 and _lib.adapt_scene_desc.device_ids.offset + 8 == C.sizeof(_lib.adapt_scene_desc)
    assert _lib.adapt_scene_desc.n_devices.offset == _lib.adapt_scene_desc.media.offset + 8
    assert _lib.adapt_scene_desc.seed.offset % 8 == 0
    assert C.sizeof(_lib.adapt_stats) == 5 * 8 + 4 * 4 + 2 * 8 + 4 * 8


def test_no_cpu_fallback(lib, scene_root):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 16, 16)
    ps = _lib.pack_scene(e, a, o, c)
    h = C.c_void_p()
    rc = lib.adapt_create(C.byref(h), C.byref(ps.desc))
    assert rc == -3 and not h.value                              # ADAPT_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.adapt_last_error()
    from adapt_b200.renderer.vanilla_renderer import Renderer
    with pytest.raises(_lib.AdaptError):
        Renderer(e, a, o, c)


def test_invalid_arguments(lib):
    assert lib.adapt_create(None, None) == -1
    assert lib.adapt_render(None, 1) == -4
    assert lib.adapt_sync(None) == -4
    assert lib.adapt_read_accum(None, None, None) == -1


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_lib.AdaptError):
        _lib.load_library(str(tmp_path / "nope.so"))


def _bvh_build(lib, prims, obj_info, wmin, wmax):
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    prims = np.ascontiguousarray(prims, np.float32).reshape(-1, 9)
    oi = np.ascontiguousarray(obj_info, np.int32)
    wmin = np.ascontiguousarray(wmin, np.float32); wmax = np.ascontiguousarray(wmax, np.float32)
    a, b, c, d = fp(), fp(), ip(), ip()
    nr, nn = C.c_int32(0), C.c_int32(0)
    rc = lib.adapt_bvh_build(prims.ctypes.data_as(fp), prims.shape[0], oi.ctypes.data_as(ip), oi.shape[1],
                             wmin.ctypes.data_as(fp), wmax.ctypes.data_as(fp), C.byref(a), C.byref(b), C.byref(c), C.byref(d),
                             C.byref(nr), C.byref(nn))
    assert rc == 0, lib.adapt_last_error()
    out = (np.ctypeslib.as_array(a, (nr.value, 2, 3)).copy(), np.ctypeslib.as_array(b, (nn.value, 2, 3)).copy(),
           np.ctypeslib.as_array(c, (nr.value, 2)).copy(), np.ctypeslib.as_array(d, (nn.value, 3)).copy())
    for ptr in (a, b, c, d):
        lib.adapt_free(ptr)
    return out


def _skip_traverse(o, d, node_minmax, node_info, bvh_minmax, bvh_info, prims, is_sphere):
    """Reference traversal (tracer/path_tracer.py:338-371) in numpy/float64 over the returned arrays."""
    inv = 1.0 / np.where(np.abs(d) < 1e-30, 1e-30, d)
    best, best_p = 1e7, -1
    i, n = 0, len(node_info)

    def slab(lo, hi):
        t0, t1 = (lo - o) * inv, (hi - o) * inv
        tn, tf = np.minimum(t0, t1).max(), np.maximum(t0, t1).min()
        return tn <= tf and tf > 0, tn
    while i < n:
        hit, tn = slab(node_minmax[i, 0] - 1e-5, node_minmax[i, 1] + 1e-5)
        if not hit or tn > best:
            i += node_info[i, 2]; continue
        if node_info[i, 2] == 1:
            for k in range(node_info[i, 0], node_info[i, 0] + node_info[i, 1]):
                p = bvh_info[k, 1]
                if is_sphere[p]:
                    continue
                v0, v1, v2 = prims[p].astype(np.float64)
                e1, e2 = v1 - v0, v2 - v0
                pv = np.cross(d, e2); det = e1 @ pv
                if abs(det) < 1e-14:
                    continue
                tv = o - v0; u = (tv @ pv) / det; qv = np.cross(tv, e1); v = (d @ qv) / det; t = (e2 @ qv) / det
                if u >= 0 and v >= 0 and u + v <= 1 and 1e-4 < t < best:
                    best, best_p = t, p
        i += 1
    return best, best_p


def test_bvh_build_reference_layout(lib, scene_root):
    e, a, o, c = load_scene(scene_root, "test", "allbxdf.xml", 16, 16)
    prims = a["primitives"]
    n = prims.shape[0]
    obj_info = np.int32([[ob.meshes.shape[0] for ob in o], [ob.type for ob in o]])
    is_sphere = np.concatenate([np.full(ob.meshes.shape[0], ob.type) for ob in o]).astype(bool)
    wmin = np.float32([-1, -1, -9]); wmax = np.float32([7, 7, 7])
    bvh_minmax, node_minmax, bvh_info, node_info = _bvh_build(lib, prims, obj_info, wmin, wmax)
    assert bvh_minmax.shape == (n, 2, 3) and bvh_info.shape == (n, 2)
    assert sorted(bvh_info[:, 1].tolist()) == list(range(n))
    np.testing.assert_array_equal(node_minmax[0], [wmin, wmax])            # root = world AABB (path_tracer.py:137-138)
    assert node_info[0, 2] == len(node_info) and node_info[0, 1] == n
    leaves = node_info[:, 2] == 1
    assert node_info[leaves, 1].sum() == n and (node_info[leaves, 1] == 1).all()     # one primitive per leaf
    assert len(node_info) == 2 * n - 1
    # object ids line up with the per-object primitive ranges
    starts = np.cumsum([0] + [ob.meshes.shape[0] for ob in o])
    for k in range(n):
        ob, p = bvh_info[k]
        assert starts[ob] <= p < starts[ob + 1]
    # children boxes are inside the parent's (except under the widened root) and skip offsets tile the array
    for i in range(1, len(node_info)):
        off = node_info[i, 2]
        if off > 1:
            l = i + 1; r = l + node_info[l, 2]
            assert r + node_info[r, 2] == i + off
            for ch in (l, r):
                assert (node_minmax[ch, 0] >= node_minmax[i, 0] - 1e-6).all() and (node_minmax[ch, 1] <= node_minmax[i, 1] + 1e-6).all()
    # the reference traversal over this tree finds the brute-force closest triangle
    rng = np.random.default_rng(2)
    tri_ids = np.where(~is_sphere)[0]
    for _ in range(40):
        ro = rng.uniform([0.5, 0.5, 0.5], [5, 5, 5]).astype(np.float64)
        rd = rng.normal(size=3); rd /= np.linalg.norm(rd)
        t, p = _skip_traverse(ro, rd, node_minmax, node_info, bvh_minmax, bvh_info, prims, is_sphere)
        best, best_p = 1e7, -1
        for q in tri_ids:
            v0, v1, v2 = prims[q].astype(np.float64)
            e1, e2 = v1 - v0, v2 - v0
            pv = np.cross(rd, e2); det = e1 @ pv
            if abs(det) < 1e-14:
                continue
            tv = ro - v0; u = (tv @ pv) / det; qv = np.cross(tv, e1); v = (rd @ qv) / det; tt = (e2 @ qv) / det
            if u >= 0 and v >= 0 and u + v <= 1 and 1e-4 < tt < best:
                best, best_p = tt, q
        assert p == best_p and abs(t - best) < 1e-9


def test_bvh_build_rejects_bad_input(lib):
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    prims = np.zeros((2, 9), np.float32); oi = np.int32([[3], [0]])         # counts say 3, only 2 primitives
    a, b, c, d = fp(), fp(), ip(), ip(); nr, nn = C.c_int32(0), C.c_int32(0)
    w = np.zeros(3, np.float32)
    rc = lib.adapt_bvh_build(prims.ctypes.data_as(fp), 2, oi.ctypes.data_as(ip), 1, w.ctypes.data_as(fp), w.ctypes.data_as(fp),
                             C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(nr), C.byref(nn))
    assert rc == -1 and b"obj_info" in lib.adapt_last_error()


def test_tile_partition_of_the_library_equals_the_python_one(lib):
    """A multi-device handle (adapt_scene_desc.n_devices > 1) splits the film with the same rule as adapt_b200/dist.py::tile_partition."""
    from adapt_b200.dist import tile_partition
    ip = C.POINTER(C.c_int32)
    for (w, h, world, tile, window) in [(64, 48, 2, 32, None), (130, 37, 3, 16, None), (1920, 1080, 8, 32, None), (200, 120, 4, 32, (40, 150, 10, 90))]:
        seen = []
        for rank in range(world):
            want = tile_partition(w, h, rank, world, tile=tile, window=window)
            win = None if window is None else (C.c_int32 * 4)(*window)
            n = lib.adapt_tile_partition(w, h, rank, world, tile, win, None, 0)
            assert n == want.size
            got = np.zeros(max(n, 1), np.int32)
            assert lib.adapt_tile_partition(w, h, rank, world, tile, win, got.ctypes.data_as(ip), n) == n
            np.testing.assert_array_equal(got[:n], want)
            seen.append(got[:n])
        allpix = np.sort(np.concatenate(seen))
        assert np.unique(allpix).size == allpix.size                # disjoint ownership
    assert lib.adapt_tile_partition(8, 8, 2, 2, 32, None, None, 0) < 0


def test_multi_device_descriptor_needs_a_gpu_too(lib, scene_root):
    """n_devices > 1 goes through the same loud failure without a device: no CPU fallback for the group handle either."""
    from adapt_b200._lib import pack_scene
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 64, 64)
    ps = pack_scene(e, a, o, c, device_ids=[0, 1])
    assert ps.desc.n_devices == 2 and ps.desc.device_ids[1] == 1
    with pytest.raises(ValueError):
        pack_scene(e, a, o, c, device_ids=[0, 1], pixel_list=np.arange(16, dtype=np.int32))
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert lib.adapt_create(C.byref(h), C.byref(ps.desc)) == -3 and not h           # ADAPT_ERR_NO_DEVICE
