"""Drop-in for the reference's pybind11 module ``bvh_cpp`` (tracer/bvh/bvh.cpp:274-296), the native boundary #2 of
DESIGN.md: ``from bvh_cpp import bvh_build`` (tracer/path_tracer.py:146) can be served by

    import sys, adapt_b200.bvh_cpp as m; sys.modules["bvh_cpp"] = m

Same signature, same four flat arrays, ownership handed to numpy; errors become ``RuntimeError`` like pybind11 turns a
C++ exception into one.  The tree is built by the C++ SAH builder behind ``adapt_bvh_build`` (include/adapt_b200.h);
this path needs no GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import load_library


def bvh_build(primitives, obj_info, world_min, world_max):
    """primitives (N,3,3) f32; obj_info (2,n_obj) i32 = [primitive count; is_sphere] (path_tracer.py:213-220);
    world_min / world_max (3,) f32.  Returns (bvh_minmax[Nref*6] f32, node_minmax[Nnode*6] f32, bvh_info[Nref*2] i32,
    node_info[Nnode*3] i32), to be reshaped as path_tracer.py:157-160 does."""
    lib = load_library()
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    p = np.ascontiguousarray(primitives, np.float32).reshape(-1, 9)
    oi = np.ascontiguousarray(obj_info, np.int32)
    if oi.ndim != 2 or oi.shape[0] != 2:
        raise RuntimeError("bvh_build: obj_info must have shape (2, n_obj)")
    wmin = np.ascontiguousarray(world_min, np.float32)
    wmax = np.ascontiguousarray(world_max, np.float32)
    a, b, c, d = fp(), fp(), ip(), ip()
    nr, nn = C.c_int32(0), C.c_int32(0)
    rc = lib.adapt_bvh_build(p.ctypes.data_as(fp), p.shape[0], oi.ctypes.data_as(ip), oi.shape[1], wmin.ctypes.data_as(fp),
                             wmax.ctypes.data_as(fp), C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(nr), C.byref(nn))
    if rc != 0:
        raise RuntimeError(f"bvh_build failed ({rc}): {lib.adapt_last_error().decode()}")
    try:
        return (np.ctypeslib.as_array(a, (nr.value * 6,)).copy(), np.ctypeslib.as_array(b, (nn.value * 6,)).copy(),
                np.ctypeslib.as_array(c, (nr.value * 2,)).copy(), np.ctypeslib.as_array(d, (nn.value * 3,)).copy())
    finally:
        for ptr in (a, b, c, d):
            lib.adapt_free(ptr)
