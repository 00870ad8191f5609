"""``Renderer`` -- the class the reference registers as ``rdr_mapping["pt"]`` (render.py:33), backed by
the sm_100a wavefront path tracer behind the C ABI (include/adapt_b200.h).

Same constructor and driver-facing surface as the reference (renderer/vanilla_renderer.py:26-30,
tracer/path_tracer.py:54-141,181-211, tracer/tracer_base.py:36-102):
``Renderer(emitters, array_info, objects, prop)``; ``render(*six_ints)`` renders exactly one spp and
bumps ``cnt``; ``reset()``; ``pixels`` / ``color`` with ``.to_numpy() -> (w, h, 3) float32`` indexed
``[x, y]`` with y up; ``cnt[None]``; ``w, h, do_crop, start_x, end_x, start_y, end_y``;
``get_check_point()`` / ``load_check_point()`` with the reference's dict keys; ``summary()``.
Extra: ``render_batch(n)`` enqueues n spp in one call, ``stats()`` returns device counters.

No CPU fallback: constructing a Renderer without the CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from .._lib import AdaptError, adapt_stats, check, load_library, pack_scene
from ..utils.tools import CONSOLE, TicToc

__all__ = ["Renderer"]


class _FieldView:
    """Stand-in for the Taichi vector fields ``pixels`` / ``color``: only ``to_numpy`` and ``shape``."""

    def __init__(self, rdr: "Renderer", mean: bool):
        self._rdr = rdr
        self._mean = mean

    @property
    def shape(self):
        return (self._rdr.w, self._rdr.h)

    def to_numpy(self, copy: bool = True) -> np.ndarray:
        """(w, h, 3) float32.  ``copy=False`` returns a view of the renderer's page-locked staging buffer (valid until the
        next read) and saves one host-side pass over the image."""
        return self._rdr._read_film(self._mean, copy)[0]

    def from_numpy(self, arr: np.ndarray):
        if self._mean:
            raise AdaptError("pixels is derived (color / cnt); load `color` instead")
        self._rdr._load_accum(arr, self._rdr.cnt[None])


class _Counter:
    """``cnt[None]`` of the reference (0-d Taichi field)."""

    def __init__(self, rdr: "Renderer"):
        self._rdr = rdr

    def __getitem__(self, _key) -> int:
        return self._rdr._cnt

    def __setitem__(self, _key, value: int):
        acc, _ = self._rdr._read_accum()
        self._rdr._load_accum(acc, int(value))


class Renderer:
    def __init__(self, emitters: List, array_info: dict, objects: List, prop: dict, *, seed: int = 0,
                 device_id: int = 0, pixel_list: Optional[np.ndarray] = None, pool_size: int = 0,
                 max_bounce: Optional[int] = None, bvh_builder=0, integrator: str = "pt", device_ids=None):
        """bvh_builder: 0 = default (the binned-SAH tree built on the device), 1 / "lbvh" = linear BVH built on the device,
        2 / "sah_device" = binned SAH on the device, 3 / "sah" = the host (OpenMP) binned-SAH build.
        integrator: "pt" (renderer/vanilla_renderer.py) or "vpt" (renderer/vpt.py over homogeneous media; see VolumeRenderer).
        device_ids: several CUDA ordinals -> ONE renderer over all of them (single process): the library replicates the scene, splits the
        film into interleaved tiles and gathers the film over NVLink peer loads when it is read (include/adapt_b200.h: n_devices)."""
        self.clock = TicToc()
        self._lib = load_library()
        self._packed = pack_scene(emitters, array_info, objects, prop, seed=seed, device_id=device_id,
                                  pixel_list=pixel_list, pool_size=pool_size, max_bounce=max_bounce,
                                  bvh_builder=bvh_builder, integrator=integrator, device_ids=device_ids)
        host = self._packed.host
        # attributes the reference driver / watermark / checkpoint code read
        for key in ("w", "h", "crop_x", "crop_y", "crop_rx", "crop_ry", "do_crop", "start_x", "end_x", "start_y",
                    "end_y", "focal", "cam_orient", "cam_t", "cam_r", "num_objects", "num_prims", "src_num"):
            setattr(self, key, host[key])
        d = self._packed.desc
        self.max_bounce = d.max_bounce
        self.use_rr = bool(d.use_rr)
        self.use_mis = bool(d.use_mis)
        self.num_shadow_ray = d.num_shadow_ray
        self.anti_alias = bool(d.anti_alias)
        self.stratified_sample = bool(d.stratified_sampling)
        self.brdf_two_sides = bool(d.brdf_two_sides)
        self.rr_threshold = d.rr_threshold
        self.rr_bounce_th = d.rr_bounce_th
        self.has_v_normal = bool(d.has_v_normal)
        self.inv_focal = 1.0 / self.focal
        self._handle = C.c_void_p()
        check(self._lib, self._lib.adapt_create(C.byref(self._handle), C.byref(self._packed.desc)), "adapt_create")
        self._cnt = 0
        self._pinned = None
        self._pinned_ptr = None
        self.pixels = _FieldView(self, mean=True)
        self.color = _FieldView(self, mean=False)
        self.cnt = _Counter(self)
        CONSOLE.log(f"Path tracer (sm_100a wavefront) initialised in {self.clock.toc_tic():.4f} s: "
                    f"{self.num_prims} primitives, {self.num_objects} objects, {self.src_num} emitters")

    # ------------------------------------------------------------------ driver surface
    def render(self, _t_start: int = 0, _t_end: int = 0, _s_start: int = 0, _s_end: int = 0, _a: int = 0, _b: int = 0):
        """One spp, like the reference kernel (the six ints are ignored there too, vanilla_renderer.py:33)."""
        self.render_batch(1)

    def render_batch(self, n_spp: int):
        """Enqueue n_spp samples per pixel; returns at once (adapt_render is asynchronous), `synchronize` / any film read waits."""
        check(self._lib, self._lib.adapt_render(self._handle, int(n_spp)), "adapt_render")
        self._cnt += int(n_spp)

    def wait(self):
        """Block until every sample enqueued so far has been handed to a path slot (paths in flight keep going)."""
        check(self._lib, self._lib.adapt_wait_enqueued(self._handle), "adapt_wait_enqueued")

    def synchronize(self):
        check(self._lib, self._lib.adapt_sync(self._handle), "adapt_sync")

    def reset(self):
        """No-op in the reference as well (tracer_base.py:284-286)."""

    def summary(self):
        self.synchronize()
        CONSOLE.rule()
        CONSOLE.print("[bold blue]:tada: :tada: :tada: Rendering Finished :tada: :tada: :tada:", justify="center")
        CONSOLE.print(f"PT SPP = {self._cnt}. Rendering time: {self.clock.toc():.3f} s", justify="center")

    # ------------------------------------------------------------------ framebuffer / checkpoint
    def _staging(self) -> np.ndarray:
        """Page-locked (w, h, 3) host buffer the device copies land in (a pageable target is several times slower)."""
        if self._pinned is None:
            nbytes = self.w * self.h * 3 * 4
            ptr = self._lib.adapt_host_alloc(nbytes)
            if not ptr:
                raise AdaptError("adapt_host_alloc failed")
            self._pinned_ptr = ptr
            self._pinned = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), (self.w, self.h, 3))
        return self._pinned

    def _read_film(self, mean: bool, copy: bool = True):
        buf = self._staging()
        spp = C.c_int32(0)
        fn, name = (self._lib.adapt_read_pixels, "adapt_read_pixels") if mean else (self._lib.adapt_read_accum, "adapt_read_accum")
        check(self._lib, fn(self._handle, buf.ctypes.data_as(C.POINTER(C.c_float)), C.byref(spp)), name)
        return (buf.copy() if copy else buf), spp.value

    def _read_accum(self):
        return self._read_film(False, True)

    def _load_accum(self, acc: np.ndarray, spp: int):
        acc = np.ascontiguousarray(acc, np.float32)
        if acc.shape != (self.w, self.h, 3):
            raise ValueError(f"accumulation buffer has shape {acc.shape}, expected {(self.w, self.h, 3)}")
        check(self._lib, self._lib.adapt_load_accum(self._handle, acc.ctypes.data_as(C.POINTER(C.c_float)), int(spp)),
              "adapt_load_accum")
        self._cnt = int(spp)

    def set_stream(self, cuda_stream: Optional[int]):
        """Run on a caller-owned CUDA stream (pass ``torch.cuda.current_stream().cuda_stream``); None restores the own stream."""
        check(self._lib, self._lib.adapt_set_stream(self._handle, C.c_void_p(cuda_stream) if cuda_stream else None), "adapt_set_stream")

    def accum_device_ptr(self):
        """(device pointer, n_floats) of the (w,h,3) sum -- used for the multi-GPU framebuffer reduce."""
        ptr = C.c_void_p()
        n = C.c_uint64(0)
        check(self._lib, self._lib.adapt_accum_device_ptr(self._handle, C.byref(ptr), C.byref(n)), "adapt_accum_device_ptr")
        return ptr.value, n.value

    def get_check_point(self) -> dict:
        """Same keys as tracer/path_tracer.py:181-193 so checkpoints interchange with the reference."""
        items = ["w", "h", "crop_x", "crop_y", "crop_rx", "crop_ry", "focal", "num_objects", "num_prims", "cam_orient", "src_num"]
        check_point = {item: getattr(self, item) for item in items}
        check_point["cam_t"] = np.asarray(self.cam_t, np.float32)
        acc, cnt = self._read_accum()
        check_point["accumulation"] = acc
        check_point["counter"] = cnt
        return check_point

    def load_check_point(self, check_point: dict):
        """Consistency check + restore (path_tracer.py:195-211); a mismatch raises instead of exit(1)."""
        for key, val in check_point.items():
            if key in ("accumulation", "counter"):
                continue
            if key == "cam_t":
                ok = np.abs(np.asarray(val) - np.asarray(self.cam_t)).max() < 1e-4
            elif key == "cam_orient":
                ok = np.abs(np.asarray(val) - np.asarray(self.cam_orient)).max() < 1e-4
            else:
                ok = val == getattr(self, key)
            if not ok:
                CONSOLE.log(f"[bold red]:skull: Error: '{key}' from the checkpoint is different.")
                raise ValueError(f"checkpoint mismatch on '{key}'")
        CONSOLE.log(f"[bold green]Recovered from check-point, elapsed counter: {check_point['counter']}")
        self._load_accum(check_point["accumulation"], int(check_point["counter"]))

    # ------------------------------------------------------------------ extras
    def stats(self, reset: bool = False) -> dict:
        st = adapt_stats()
        check(self._lib, self._lib.adapt_get_stats(self._handle, C.byref(st)), "adapt_get_stats")
        out = st.as_dict()
        out["rays_culled"], out["fused_trace"], out["pool_slots"] = int(st.reserved[0]), bool(st.reserved[1]), int(st.reserved[2])
        out["lanes"] = max(1, int(st.reserved[3]))
        if reset:
            check(self._lib, self._lib.adapt_reset_stats(self._handle), "adapt_reset_stats")
        return out

    def intersect_batch(self, rays_o, rays_d, tmax=None, any_hit: bool = False):
        """Stage-level hook: trace rays through the device BVH (closest hit or any hit)."""
        ro = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3)
        rd = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
        n = ro.shape[0]
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int32)
        tm = None if tmax is None else np.ascontiguousarray(tmax, np.float32)
        obj = np.zeros(n, np.int32); prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        check(self._lib, self._lib.adapt_intersect_batch(
            self._handle, ro.ctypes.data_as(fp), rd.ctypes.data_as(fp), None if tm is None else tm.ctypes.data_as(fp), n,
            int(any_hit), obj.ctypes.data_as(ip), prim.ctypes.data_as(ip), t.ctypes.data_as(fp), u.ctypes.data_as(fp),
            v.ctypes.data_as(fp)), "adapt_intersect_batch")
        return dict(obj=obj, prim=prim, t=t, u=u, v=v)

    def update_geometry(self, primitives, n_g, n_s=None, refit: bool = False):
        """New vertex positions for the same topology (animated meshes): (N,3,3) primitives, (N,3) geometric normals and, when
        the scene has vertex normals, (N,3,3) shading normals -- the arrays of ``array_info``.  Rebuilds the acceleration
        structure with this renderer's builder, or with ``refit=True`` keeps the tree and only recomputes its boxes on the device
        (adapt_refit_geometry); the accumulation buffer is kept (``reset_accumulation()`` starts over)."""
        fp = C.POINTER(C.c_float)
        n = self.num_prims
        pr = np.ascontiguousarray(primitives, np.float32).reshape(-1)
        ng = np.ascontiguousarray(n_g, np.float32).reshape(-1)
        if pr.size != n * 9 or ng.size != n * 3:
            raise ValueError(f"update_geometry: expected {n} primitives")
        ns = None
        if n_s is not None:
            ns = np.ascontiguousarray(n_s, np.float32).reshape(-1)
            if ns.size != n * 9:
                raise ValueError(f"update_geometry: expected {n} x 3 shading normals")
        fn, name = (self._lib.adapt_refit_geometry, "adapt_refit_geometry") if refit else (self._lib.adapt_update_geometry, "adapt_update_geometry")
        check(self._lib, fn(self._handle, pr.ctypes.data_as(fp), ng.ctypes.data_as(fp), None if ns is None else ns.ctypes.data_as(fp)), name)

    def reset_accumulation(self, spp: int = 0):
        """Empty film (cleared on the device); the next sample rendered is number ``spp + 1``."""
        check(self._lib, self._lib.adapt_load_accum(self._handle, None, int(spp)), "adapt_load_accum")
        self._cnt = int(spp)

    def bvh_export(self, arrays: bool = True) -> dict:
        """Stage-level hook: the device acceleration structure (64-byte nodes, 48-byte leaf records; the 80-byte nodes of the compressed
        8-wide tree when the handle traces through one), its builder and build time."""
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        nn, npr, dep, bld = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        ms = C.c_float()
        check(self._lib, self._lib.adapt_bvh_export(self._handle, C.byref(nn), C.byref(npr), C.byref(dep), C.byref(bld), C.byref(ms),
                                                    None, None), "adapt_bvh_export")
        out = dict(n_nodes=nn.value, n_prims=npr.value, depth=dep.value, builder=bld.value, build_ms=ms.value)
        nn8, dep8 = C.c_int32(), C.c_int32()
        check(self._lib, self._lib.adapt_bvh_export_wide(self._handle, C.byref(nn8), C.byref(dep8), None), "adapt_bvh_export_wide")
        out.update(n_nodes8=nn8.value, depth8=dep8.value)           # 0: the handle traces through the binary tree
        if arrays and nn8.value > 0:
            nodes8 = np.zeros((nn8.value, 20), np.uint32)
            check(self._lib, self._lib.adapt_bvh_export_wide(self._handle, None, None, nodes8.ctypes.data_as(C.POINTER(C.c_uint32))), "adapt_bvh_export_wide")
            out.update(nodes8=nodes8)
        if arrays:
            nodes = np.zeros((nn.value, 16), np.float32); prims = np.zeros((npr.value, 12), np.float32)
            check(self._lib, self._lib.adapt_bvh_export(self._handle, None, None, None, None, None, nodes.ctypes.data_as(fp),
                                                        prims.ctypes.data_as(fp)), "adapt_bvh_export")
            out.update(nodes=nodes, prims=prims)
        return out

    def bxdf_batch(self, obj: int, n_s, n_g, incid, out, two_sides: bool = False, seed: int = 0):
        """Stage-level hook: eval / pdf / sample of object ``obj``'s surface model on the device (sample k draws from (seed, k, 0))."""
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        arrs = [np.ascontiguousarray(x, np.float32).reshape(-1, 3) for x in (n_s, n_g, incid, out)]
        n = arrs[0].shape[0]
        ev = np.zeros((n, 3), np.float32); sd = np.zeros((n, 3), np.float32); ss = np.zeros((n, 3), np.float32)
        pdf = np.zeros(n, np.float32); sp = np.zeros(n, np.float32); fl = np.zeros(n, np.int32)
        check(self._lib, self._lib.adapt_bxdf_batch(
            self._handle, int(obj), n, *(x.ctypes.data_as(fp) for x in arrs), int(bool(two_sides)), int(seed),
            ev.ctypes.data_as(fp), pdf.ctypes.data_as(fp), sd.ctypes.data_as(fp), ss.ctypes.data_as(fp), sp.ctypes.data_as(fp),
            fl.ctypes.data_as(ip)), "adapt_bxdf_batch")
        return dict(eval=ev, pdf=pdf, s_dir=sd, s_spec=ss, s_pdf=sp, s_flag=fl)

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle:
            self._lib.adapt_destroy(self._handle)
            self._handle = C.c_void_p()
        if getattr(self, "_pinned_ptr", None):
            self._pinned = None
            self._lib.adapt_host_free(self._pinned_ptr)
            self._pinned_ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
