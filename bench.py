#!/usr/bin/env python
"""bench.py -- headline benchmark of the unidirectional path-tracing hot path (BASELINE.json metric:
Mrays/s + spp/s at 1080p, 16 bounces; 1/2/4/8 B200 vs the CPU path).

    python bench.py --gpus 1 --steps K --warmup W                      # this repo's sm_100a wavefront tracer
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W      # the reference estimator on the host cores (CPU oracle)

One "step" renders `spp_per_step` samples per pixel of the workload scene (default: BASELINE config 3,
the 89 888-triangle "bunny90k" stand-in in the Cornell box at 1920x1080, 16 bounces, 1 shadow ray).
With N ranks the film is tile-partitioned, every rank renders N * spp_per_step samples of its own
pixels (weak scaling: per-GPU work is fixed) and each step ends with the NCCL sum-reduce of the HDR
framebuffer.  `value` = closest-hit rays (primary + secondary, as counted on the device) per second,
whole job, with the scene resident in HBM; `e2e` = the same metric through the checkpoint-style public
API with host buffers: load_check_point (H2D of the pinned (w,h,3) accumulation) -> render -> read back.

Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("ADAPT_QUIET", "1")

WORKLOADS = {
    # name: (scene dir, xml, big meshes needed, BASELINE config)
    "bunny90k": ("cbox", "bunny90k.xml", ("bunny90k",), "configs[2]: bunny ~90k tris, 1920x1080, 16 bounces"),
    "orb500k": ("cbox", "orb500k.xml", ("orb500k",), "configs[3]: material-orb ~500k tris, 1920x1080, 24 bounces"),
    "car290k": ("cbox", "car290k.xml", ("car290k",), "configs[4]: sports-car ~290k tris, 3840x2160, 16 bounces"),
    "balls-mono": ("csphere", "balls-mono.xml", (), "configs[1]: cornell-spheres (film as in the XML unless --width)"),
    "cbox": ("cbox", "cbox.xml", (), "configs[0]: cornell box"),
    # for --integrator vpt (not a BASELINE config): world fog + media-filled objects (adapt_b200/scenes.py::write_media)
    "media": ("test", "media.xml", (), "volumetric coverage scene"),
}


def load_workload(name, width=None, height=None, max_bounce=None):
    from adapt_b200.parsers.xml_parser import scene_parsing
    from adapt_b200.scenes import DEFAULT_ROOT, ensure_big_meshes, ensure_small_scenes
    scene, xml, big, _ = WORKLOADS[name]
    root = ensure_small_scenes(DEFAULT_ROOT)
    if big:
        ensure_big_meshes(root, big)
    e, a, o, c = scene_parsing(os.path.join(root, scene), xml)
    if width:
        c["film"]["width"] = int(width)
        c["film"]["height"] = int(height or width)
    if max_bounce is not None and max_bounce >= 0:
        c["max_bounce"] = int(max_bounce)
    return e, a, o, c


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU every 200 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, threading.Event(), [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def result(self):
        self.stop_flag.set()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_leg(args, e, a, o, c, budget_s=15.0, n_threads=0):
    """Times the CPU oracle (restatement of the reference estimator; the reference itself needs Taichi,
    which is not installable here) on a bounded sample of the workload: a centred block of film tiles."""
    from adapt_b200._lib import pack_scene
    from adapt_b200.dist import tile_partition
    from oracle.pt_oracle import OracleScene
    w, h = c["film"]["width"], c["film"]["height"]
    osc = OracleScene(pack_scene(e, a, o, c, seed=args.seed, integrator=args.integrator))
    cores = n_threads or (os.cpu_count() or 1)
    # pilot: a 192x192 block in the image centre (run twice: the first call warms caches / the OpenMP pool)
    cw, ch = max(0, (w // 2 - 96)) // 32 * 32, max(0, (h // 2 - 96)) // 32 * 32
    pilot = tile_partition(w, h, 0, 1, window=(cw, min(cw + 192, w), ch, min(ch + 192, h)))
    osc.render(1, pixel_list=pilot, n_threads=cores)
    t0 = time.time(); _, cn = osc.render(1, cnt_start=1, pixel_list=pilot, n_threads=cores); dt = max(time.time() - t0, 1e-4)
    rate = len(pilot) / dt
    n_tiles = int(max(1, min((w // 32) * (h // 32), budget_s * rate / 1024)))
    side = int(max(1, np.floor(np.sqrt(n_tiles))))
    x0 = max(0, (w // 2) - side * 16) // 32 * 32; y0 = max(0, (h // 2) - side * 16) // 32 * 32
    window = (x0, min(w, x0 + side * 32), y0, min(h, y0 + side * 32))
    sample = tile_partition(w, h, 0, 1, window=window)
    spp = int(max(1, min(64, budget_s * rate / max(len(sample), 1))))      # whole film is cheap: take several spp
    return osc, sample, window, cores, spp


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (here: the CPU oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    e, a, o, c = load_workload(args.workload, args.width, args.height, args.max_bounce)
    osc, sample, window, cores, spp = cpu_reference_leg(args, e, a, o, c, budget_s=args.cpu_budget / max(args.steps, 1))
    for _ in range(args.warmup):
        osc.render(1, pixel_list=sample[: max(256, len(sample) // 16)], n_threads=cores)
    rays = 0; paths = 0
    t0 = time.time()
    for k in range(args.steps):
        _, cn = osc.render(spp, cnt_start=k * spp, pixel_list=sample, n_threads=cores)
        rays += cn["rays_closest"]; paths += cn["paths"]
    dt = time.time() - t0
    w, h = c["film"]["width"], c["film"]["height"]
    value = rays / dt / 1e6
    sample_desc = f"{len(sample)} pixels (window x[{window[0]},{window[1]}) y[{window[2]},{window[3]})) x {spp} spp per step"
    line = {
        "impl": "reference", "metric": "Mrays/s (closest-hit rays: primary + secondary)", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} {w}x{h}, max_bounce {c['max_bounce']}, nsr {c['num_shadow_ray']} ({WORKLOADS[args.workload][3]})" + ("" if args.integrator == "pt" else f", integrator {args.integrator}"),
                   "spp_per_step": spp, "sample": sample_desc},
        "spp_per_s": paths / dt / (w * h),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from adapt_b200.dist import device_tensor_view, init_process_group, reduce_framebuffer, tile_partition
    from adapt_b200.renderer.vanilla_renderer import Renderer
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the adapt_b200 path has no CPU fallback")
    rank, local_rank, world = init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    e, a, o, c = load_workload(args.workload, args.width, args.height, args.max_bounce)
    w, h = c["film"]["width"], c["film"]["height"]
    pixel_list = tile_partition(w, h, rank, world, tile=32) if world > 1 else None
    rdr = Renderer(e, a, o, c, seed=args.seed, device_id=local_rank, pixel_list=pixel_list, pool_size=args.pool, integrator=args.integrator)
    stream = torch.cuda.current_stream()
    rdr.set_stream(stream.cuda_stream)
    ptr, nfl = rdr.accum_device_ptr()
    fb = device_tensor_view(ptr, nfl, local_rank)
    spp_step = args.spp_per_step * world          # weak scaling: per-GPU work fixed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the reduce works on a copy: the renderer's own buffer keeps accumulating across steps, and summing it in place on
    # rank 0 would count the other ranks' pixels once per step
    fb_out = torch.empty_like(fb) if world > 1 else None

    def step():
        rdr.render_batch(spp_step)
        rdr.synchronize()
        if world > 1:
            fb_out.copy_(fb)
            reduce_framebuffer(fb_out, dst=0)

    for _ in range(args.warmup):
        step()
    barrier()
    rdr.stats(reset=True)
    sampler = ClockSampler(local_rank); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.result()
    st = rdr.stats()
    tot = torch.tensor([ms, float(st["rays_closest"]), float(st["rays_shadow"]), float(st["paths"]), float(st["kernel_launches"]),
                        float(st["ms_closest"]), float(st["iterations"]), float(st["ms_logic"]), float(st["ms_shadow"])],
                       dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        ms = float(mx[0])
    tot = tot.cpu().numpy()
    rays_closest, rays_shadow, paths, launches = tot[1], tot[2], tot[3], tot[4]
    value = rays_closest / (ms * 1e-3) / 1e6

    # ---- e2e: checkpoint-style public API with host buffers (H2D + render + D2H inside the timed region)
    nbytes = w * h * 3 * 4
    host_in = torch.zeros((w, h, 3), dtype=torch.float32).pin_memory()
    host_np = host_in.numpy()
    ck = rdr.get_check_point()
    # one untimed pass through the same calls: the page-locked staging buffer of to_numpy() is allocated on first use
    ck["accumulation"] = host_np; ck["counter"] = 0
    rdr.load_check_point(ck); rdr.render_batch(1); rdr.pixels.to_numpy(copy=False)
    e2e_rays0 = rdr.stats(reset=True)["rays_closest"]
    barrier()
    ev0.record(stream)
    for k in range(args.steps):
        ck["accumulation"] = host_np
        ck["counter"] = k * spp_step
        rdr.load_check_point(ck)                       # H2D of the pinned accumulation buffer
        rdr.render_batch(spp_step)
        if world > 1:
            rdr.synchronize(); fb_out.copy_(fb); reduce_framebuffer(fb_out, dst=0)
            out = (fb_out.cpu().numpy() if rank == 0 else rdr.pixels.to_numpy(copy=False))   # D2H of the reduced film (sync point)
        else:
            out = rdr.pixels.to_numpy(copy=False)      # D2H of the mean buffer into page-locked memory (sync point)
    ev1.record(stream)
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    e2e_st = rdr.stats()
    e2e_t = torch.tensor([e2e_ms, float(e2e_st["rays_closest"])], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        mx = e2e_t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(e2e_t, op=dist.ReduceOp.SUM)
        e2e_ms = float(mx[0])
    e2e_value = float(e2e_t[1]) / (e2e_ms * 1e-3) / 1e6
    assert np.isfinite(out).all()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        # CPU baseline + reference-traversal statistics on a bounded sample (rank 0, N=1 only)
        cpu = None; nbar_node = nbar_prim = nbar_node_s = nbar_prim_s = None
        if world == 1 and not args.no_cpu:
            osc, sample, window, cores, cspp = cpu_reference_leg(args, e, a, o, c, budget_s=args.cpu_budget)
            t0 = time.time(); _, cn = osc.render(cspp, pixel_list=sample, n_threads=cores); dt = time.time() - t0
            cpu = {"value": cn["rays_closest"] / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": f"{len(sample)} pixels (window x[{window[0]},{window[1]}) y[{window[2]},{window[3]})) x {cspp} spp, {dt:.1f} s"}
            if cn["nodes_visited"]:
                nbar_node = cn["nodes_visited"] / cn["rays_closest"]; nbar_prim = cn["prims_tested"] / cn["rays_closest"]
                nbar_node_s = cn["nodes_shadow"] / max(cn["rays_shadow"], 1); nbar_prim_s = cn["prims_shadow"] / max(cn["rays_shadow"], 1)
        # roofline of the dominant kernel, SURVEY 8(d): per closest-hit ray B = 48 B queue + 36 B * nodes + 104 B * prims under the
        # reference traversal order (oracle counters); per shadow ray 60 B queue + the same with the any-hit statistics.  The
        # fused k_trace launch processes both streams, so its algorithmic bytes are the sum; avg launch duration from the
        # CUDA events the library records around every launch on its stream
        fused = bool(st.get("fused_trace"))
        iters = max(tot[6], 1.0)
        avg_ms = (tot[5] + (tot[8] if fused else 0.0)) / iters / world
        rays_per_launch = rays_closest / iters / world
        shadow_per_launch = rays_shadow / iters / world
        b_queue, b_queue_s = 48.0, 60.0
        b_bvh = (36.0 * nbar_node + 104.0 * nbar_prim) if nbar_node else None
        b_ray = b_queue + (b_bvh or 0.0)
        b_ray_s = b_queue_s + ((36.0 * nbar_node_s + 104.0 * nbar_prim_s) if nbar_node else 0.0)
        bytes_per_launch = rays_per_launch * b_ray + (shadow_per_launch * b_ray_s if fused else 0.0)
        queue_bytes_per_launch = rays_per_launch * b_queue + (shadow_per_launch * b_queue_s if fused else 0.0)
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        kern = "k_trace" if fused else "k_closest"
        # k_logic: 104 B state read + 88 B state write per live slot, 48 B per shadow ray written
        pool_slots = int(st.get("pool_slots") or args.pool or int(os.environ.get("ADAPT_POOL", 0)) or 0)
        logic_bytes = pool_slots * 192.0 + shadow_per_launch * 48.0
        traffic = logic_traffic = None
        # ncu DRAM bytes per launch (profiles/ncu_summary.json) were captured on the default workload only
        summary_ok = args.workload == "bunny90k" and not args.width
        try:
            if not summary_ok:
                raise KeyError("no ncu capture for this workload")
            logic_traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get("k_logic", {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        try:
            if not summary_ok:
                raise KeyError("no ncu capture for this workload")
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get(kern, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "Mrays/s (closest-hit rays: primary + secondary)", "value": value, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} {w}x{h}, max_bounce {c['max_bounce']}, nsr {c['num_shadow_ray']} ({WORKLOADS[args.workload][3]})" + ("" if args.integrator == "pt" else f", integrator {args.integrator}"),
                       "spp_per_step": spp_step, "pool_slots": pool_slots,
                       "bvh": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in rdr.bvh_export(arrays=False).items()
                               if k in ("builder", "n_nodes", "depth", "build_ms")},
                       "parallelism": f"tile-split x{world}" if world > 1 else "single GPU",
                       "l2": "path pool + queues (>200 MB) stream through HBM every wavefront iteration (> 126 MB L2); the BVH stays L2-resident by design"},
            "spp_per_s": paths / (w * h) / (ms * 1e-3),
            "mrays_shadow_per_s": rays_shadow / (ms * 1e-3) / 1e6,
            "paths_per_s": paths / (ms * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "stage_ms_per_step": {"logic": tot[7] / world / args.steps, "shadow": tot[8] / world / args.steps, "closest": tot[5] / world / args.steps},
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_ray": b_ray, "bytes_per_shadow_ray": b_ray_s if fused else None,
                         "bytes_per_ray_queue_only": b_queue,
                         "frac_queue_only": queue_bytes_per_launch / (avg_ms * 1e-3) / 1e9 / peak_gbs,
                         "ref_nodes_per_ray": nbar_node, "ref_prims_per_ray": nbar_prim,
                         "ref_nodes_per_shadow_ray": nbar_node_s, "ref_prims_per_shadow_ray": nbar_prim_s, "avg_launch_ms": avg_ms,
                         "rays_per_launch": rays_per_launch, "shadow_rays_per_launch": shadow_per_launch if fused else None,
                         "note": "achieved = the logical bytes the REFERENCE traversal would touch for the rays of one launch (SURVEY 8(d)) / launch time; "
                                 "it can exceed the HBM peak because the BVH is served from L2/L1 -- frac_queue_only (compulsory HBM bytes) and traffic (ncu dram bytes) are the DRAM-side figures",
                         # second kernel of the iteration: streams the whole path pool (HBM-bound by construction)
                         "k_logic": {"achieved": logic_bytes / max(tot[7] / iters / world * 1e-3, 1e-9) / 1e9,
                                     "frac": logic_bytes / max(tot[7] / iters / world * 1e-3, 1e-9) / 1e9 / peak_gbs,
                                     "bytes_per_launch": logic_bytes, "avg_launch_ms": tot[7] / iters / world,
                                     "traffic": logic_traffic}},
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything libraries print meanwhile (NCCL's version banner,
    torchrun chatter) has been diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="bunny90k", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=32)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--max-bounce", type=int, default=None)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU-oracle work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--integrator", default="pt", choices=["pt", "vpt"],
                    help="vpt: the volumetric integrator over homogeneous media (use with --workload cbox or media)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
