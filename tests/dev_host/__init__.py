"""TEST INFRASTRUCTURE: the device code of the volumetric integrator compiled as host C++ (see dev_host.cpp, cuda_host_shim.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "adapt_b200", "csrc")
SRC = os.path.join(_HERE, "dev_host.cpp")
DEPS = [SRC, os.path.join(_HERE, "cuda_host_shim.h"), os.path.join(_HERE, "dev_scene.h")] + [os.path.join(_CSRC, f) for f in (
    "pt_common.cuh", "pt_shade.cuh", "pt_trace.cuh", "pt_path.cuh", "pt_volume.cuh", "scene_pack.h", "bvh_build.cpp", "bvh_build.h")]
LIB = os.path.join(_HERE, "_build", "libdev_host.so")
WF_SRC = os.path.join(_HERE, "wavefront_host.cpp")
WF_LIB = os.path.join(_HERE, "_build", "libwavefront_host.so")
WF_DEPS = [WF_SRC, os.path.join(_HERE, "simt_emu.h"), os.path.join(_HERE, "cuda_host_shim.h"), os.path.join(_HERE, "dev_scene.h"),
           os.path.join(_CSRC, "pt_kernels.cuh")]
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # same contraction / ISA choices as the oracle build (nvcc contracts a*b+c into FMA by default as well)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-fPIC", "-fopenmp", "-ffp-contract=fast", "-march=x86-64-v3", "-I" + CUDA_INC,
                           "-shared", "-o", LIB, SRC, os.path.join(_CSRC, "bvh_build.cpp")])
    return LIB


def build_wavefront(force: bool = False) -> str:
    """The kernels themselves (pt_kernels.cuh) as host C++ under the SIMT emulator (simt_emu.h); C++20 for nothing but designated habits of
    the headers, -O2 because the emulated kernels are the hot loop of these tests."""
    deps = WF_DEPS + DEPS[3:]
    # DEV_HOST_CXXFLAGS="-DTRACE_LEAF_ONE=1 ...": a compile-time variant of the kernels (own library file per flag set)
    extra = os.environ.get("DEV_HOST_CXXFLAGS", "").split()
    lib = WF_LIB if not extra else WF_LIB[:-3] + "_" + "".join(c if c.isalnum() else "_" for c in "".join(extra)) + ".so"
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-w", "-fPIC", "-ffp-contract=fast", "-march=x86-64-v3", "-I" + CUDA_INC] + extra +
                          ["-shared", "-o", lib, WF_SRC, os.path.join(_CSRC, "bvh_build.cpp")])
    return lib


_wf = None


def wavefront_render(packed, n_spp: int, pool_slots: int = 256, trace_grid: int = 2, cnt_start: int = 0):
    """Run the library's kernels under the SIMT emulator on one packed scene -> (film sums (w,h,3), stats dict)."""
    global _wf
    if _wf is None:
        _wf = C.CDLL(build_wavefront())
        _wf.wavefront_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    w, h = packed.desc.width, packed.desc.height
    acc = np.zeros((w, h, 3), np.float32)
    st = np.zeros(6, np.uint64)
    rc = _wf.wavefront_render(C.addressof(packed.desc), int(n_spp), int(pool_slots), int(trace_grid), int(cnt_start), acc.ctypes.data_as(C.POINTER(C.c_float)),
                              st.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc != 0:
        raise RuntimeError(f"emulated wavefront failed ({rc}): no progress" if rc == -1 else f"emulated wavefront failed ({rc})")
    return acc, dict(paths=int(st[0]), rays_closest=int(st[1]), rays_shadow=int(st[2]), iterations=int(st[3]), launches=int(st[4]), rays_culled=int(st[5]))


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        lib.dev_host_create.restype = C.c_void_p
        lib.dev_host_create.argtypes = [C.c_void_p, C.c_int]
        lib.dev_host_bxdf_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, fp, fp, fp, C.c_int, C.c_uint64, fp, fp, fp, fp, fp, ip]
        lib.dev_host_intersect_batch.argtypes = [C.c_void_p, fp, fp, fp, C.c_int, C.c_int, ip, ip, fp, fp, fp]
        lib.dev_host_destroy.argtypes = [C.c_void_p]
        lib.dev_host_render_vpt.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, C.POINTER(C.c_uint64)]
        lib.dev_host_phase_eval.argtypes = [C.c_void_p, fp, fp, C.c_int, fp]
        lib.dev_host_phase_sample.argtypes = [C.c_void_p, fp, C.c_uint64, C.c_int, fp, fp]
        lib.dev_host_medium_sample_mfp.argtypes = [C.c_void_p, C.c_float, C.c_uint64, C.c_int, ip, fp, fp]
        lib.dev_host_index_arith_check.argtypes = [C.c_uint64, C.c_int]
        _lib = lib
    return _lib


class DevHostScene:
    """vpt through the device functions (vol_shade_step / vol_transmit_step / trace) on the CPU, for one packed scene."""

    def __init__(self, packed, for_vpt: bool = True):
        self.lib = load()
        self.packed = packed
        self.h = self.lib.dev_host_create(C.addressof(packed.desc), int(for_vpt))
        if not self.h:
            raise RuntimeError("dev_host_create failed")
        self.w, self.hh = packed.desc.width, packed.desc.height

    def render(self, n_spp: int, cnt_start: int = 0):
        acc = np.zeros((self.w, self.hh, 3), np.float32)
        st = np.zeros(3, np.uint64)
        self.lib.dev_host_render_vpt(self.h, cnt_start, n_spp, acc.ctypes.data_as(C.POINTER(C.c_float)), st.ctypes.data_as(C.POINTER(C.c_uint64)))
        return acc, dict(paths=int(st[0]), traces=int(st[1]), segments=int(st[2]))

    def bxdf_batch(self, obj, n_s, n_g, incid, out, two_sides=False, seed=0):
        """Host twin of Renderer.bxdf_batch / k_bxdf_batch: the device surface models of object ``obj``."""
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        arrs = [np.ascontiguousarray(x, np.float32).reshape(-1, 3) for x in (n_s, n_g, incid, out)]
        n = arrs[0].shape[0]
        ev = np.zeros((n, 3), np.float32); sd = np.zeros((n, 3), np.float32); ss = np.zeros((n, 3), np.float32)
        pdf = np.zeros(n, np.float32); sp = np.zeros(n, np.float32); fl = np.zeros(n, np.int32)
        self.lib.dev_host_bxdf_batch(self.h, int(obj), n, *(x.ctypes.data_as(fp) for x in arrs), int(bool(two_sides)), int(seed),
                                     ev.ctypes.data_as(fp), pdf.ctypes.data_as(fp), sd.ctypes.data_as(fp), ss.ctypes.data_as(fp),
                                     sp.ctypes.data_as(fp), fl.ctypes.data_as(ip))
        return dict(eval=ev, pdf=pdf, s_dir=sd, s_spec=ss, s_pdf=sp, s_flag=fl)

    def intersect_batch(self, rays_o, rays_d, tmax=None, any_hit=False):
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        ro = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3); rd = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
        n = ro.shape[0]
        tm = None if tmax is None else np.ascontiguousarray(tmax, np.float32)
        obj = np.zeros(n, np.int32); prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        self.lib.dev_host_intersect_batch(self.h, ro.ctypes.data_as(fp), rd.ctypes.data_as(fp), None if tm is None else tm.ctypes.data_as(fp), n,
                                          int(any_hit), obj.ctypes.data_as(ip), prim.ctypes.data_as(ip), t.ctypes.data_as(fp),
                                          u.ctypes.data_as(fp), v.ctypes.data_as(fp))
        return dict(obj=obj, prim=prim, t=t, u=u, v=v)

    def __del__(self):
        try:
            if self.h:
                self.lib.dev_host_destroy(self.h)
                self.h = None
        except Exception:
            pass
