"""ctypes wrapper of the CPU oracle (oracle/_ref/liboracle.so, built by oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  Nothing under adapt_b200/ imports this module.  Parity unpinned by the
reference (see the header of pt_oracle.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from adapt_b200._lib import PackedScene, adapt_bxdf, adapt_scene_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle.so")

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_lib = None


def build(force: bool = False):
    src = os.path.join(_HERE, "pt_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.oracle_create.argtypes = [C.POINTER(adapt_scene_desc)]
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_render.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, _ip, C.c_int, C.c_int, _u64p]
    lib.oracle_render_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp, _u64p]
    lib.oracle_intersect_batch.argtypes = [C.c_void_p, _fp, _fp, _fp, C.c_int, C.c_int, _ip, _ip, _fp, _fp, _fp, _fp, _u64p]
    lib.oracle_bvh_build.argtypes = [_fp, C.c_int32, _ip, C.c_int32, _fp, _fp,
                                     C.POINTER(_fp), C.POINTER(_fp), C.POINTER(_ip), C.POINTER(_ip), _ip, _ip]
    lib.oracle_bvh_build.restype = C.c_int
    lib.oracle_free.argtypes = [C.c_void_p]
    lib.oracle_fresnel_equation.argtypes = [C.c_float] * 4
    lib.oracle_fresnel_equation.restype = C.c_float
    lib.oracle_rotation_between.argtypes = [_fp, _fp, _fp]
    lib.oracle_rng_stream.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
    lib.oracle_bxdf_eval.argtypes = [C.POINTER(adapt_bxdf), _fp, _fp, _fp, C.c_float, _fp, _fp]
    lib.oracle_bxdf_sample.argtypes = [C.POINTER(adapt_bxdf), _fp, _fp, C.c_float, C.c_uint64, C.c_uint32, _fp, _fp, _fp, _ip]
    _lib = lib
    return lib


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


COUNTER_NAMES = ["paths", "rays_closest", "rays_shadow", "nodes_visited", "prims_tested", "rng_draws", "rays_closest_useful"]


class OracleScene:
    """CPU restatement of PathTracer + Renderer for one packed scene."""

    def __init__(self, packed: PackedScene, force_bvh: Optional[bool] = None):
        self.lib = load()
        self.packed = packed
        if force_bvh is not None:
            packed.desc.reserved[0] = 1 if force_bvh else 0
        self.h = self.lib.oracle_create(C.byref(packed.desc))
        self.w_, self.h_ = packed.desc.width, packed.desc.height

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def render(self, n_spp: int, cnt_start: int = 0, accum: Optional[np.ndarray] = None,
               pixel_list: Optional[np.ndarray] = None, n_threads: int = 0):
        """Adds samples cnt_start+1..cnt_start+n_spp to ``accum`` (w,h,3) and returns (accum, counters)."""
        if accum is None:
            accum = np.zeros((self.w_, self.h_, 3), np.float32)
        assert accum.dtype == np.float32 and accum.flags.c_contiguous
        counters = np.zeros(8, np.uint64)
        if pixel_list is not None:
            pl = np.ascontiguousarray(pixel_list, np.int32)
            self.lib.oracle_render(self.h, cnt_start, n_spp, _p(accum), _p(pl, C.c_int32), pl.size, n_threads, _p(counters, C.c_uint64))
        else:
            self.lib.oracle_render(self.h, cnt_start, n_spp, _p(accum), None, 0, n_threads, _p(counters, C.c_uint64))
        return accum, dict(zip(COUNTER_NAMES, (int(c) for c in counters[:7])))

    def render_sample(self, i: int, j: int, cnt: int):
        rgb = np.zeros(3, np.float32)
        draws = C.c_uint64(0)
        self.lib.oracle_render_sample(self.h, i, j, cnt, _p(rgb), C.byref(draws))
        return rgb, draws.value

    def intersect_batch(self, ro, rd, tmax=None, any_hit=False):
        ro = np.ascontiguousarray(ro, np.float32)
        rd = np.ascontiguousarray(rd, np.float32)
        n = ro.shape[0]
        tm = None if tmax is None else np.ascontiguousarray(tmax, np.float32)
        obj = np.zeros(n, np.int32); prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        ns = np.zeros((n, 3), np.float32)
        counters = np.zeros(4, np.uint64)
        self.lib.oracle_intersect_batch(self.h, _p(ro), _p(rd), None if tm is None else _p(tm), n, int(any_hit),
                                        _p(obj, C.c_int32), _p(prim, C.c_int32), _p(t), _p(u), _p(v), _p(ns),
                                        _p(counters, C.c_uint64))
        return dict(obj=obj, prim=prim, t=t, u=u, v=v, n_s=ns,
                    nodes_visited=int(counters[2]), prims_tested=int(counters[3]))


def bvh_build(primitives: np.ndarray, obj_info: np.ndarray, world_min: np.ndarray, world_max: np.ndarray):
    """Oracle twin of the reference's bvh_cpp.bvh_build: returns the 4 flat arrays."""
    lib = load()
    prims = np.ascontiguousarray(primitives, np.float32).reshape(-1, 9)
    oi = np.ascontiguousarray(obj_info, np.int32)
    wmin = np.ascontiguousarray(world_min, np.float32); wmax = np.ascontiguousarray(world_max, np.float32)
    a, b = _fp(), _fp()
    c, d = _ip(), _ip()
    nr, nn = C.c_int32(0), C.c_int32(0)
    lib.oracle_bvh_build(_p(prims), prims.shape[0], _p(oi, C.c_int32), oi.shape[1], _p(wmin), _p(wmax),
                         C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(nr), C.byref(nn))
    out = (np.ctypeslib.as_array(a, (nr.value * 6,)).copy(), np.ctypeslib.as_array(b, (nn.value * 6,)).copy(),
           np.ctypeslib.as_array(c, (nr.value * 2,)).copy(), np.ctypeslib.as_array(d, (nn.value * 3,)).copy())
    for ptr in (a, b, c, d):
        lib.oracle_free(ptr)
    return out
