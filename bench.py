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


def workload_config(args, c, world):
    """The `config` dict of the JSON line: what is rendered, identical for both arms (the reference arm times a bounded sample of it)."""
    w, h = c["film"]["width"], c["film"]["height"]
    return {"workload": f"{args.workload} {w}x{h}, max_bounce {c['max_bounce']}, nsr {c['num_shadow_ray']} ({WORKLOADS[args.workload][3]})"
                        + ("" if args.integrator == "pt" else f", integrator {args.integrator}"),
            "scene": WORKLOADS[args.workload][1], "width": w, "height": h, "max_bounce": int(c["max_bounce"]),
            "num_shadow_ray": int(c["num_shadow_ray"]), "integrator": args.integrator, "seed": args.seed,
            "parallelism": f"film split into interleaved 32x32 tiles over {world} GPUs" if world > 1 else "single GPU",
            "l2": "inputs larger than L2: path pool + queues (>200 MB) stream through HBM every wavefront iteration (126 MB L2); the BVH stays L2-resident by design"}


def cpu_reference_leg(args, e, a, o, c, budget_s=15.0, n_threads=0):
    """Times the CPU oracle (restatement of the reference estimator; the reference itself needs Taichi, which is not installable
    here) on a bounded sample of the SAME workload: every K-th 32x32 tile of the whole film (tile_partition(w, h, 0, K)), so the
    mix of rays -- sky / wall / mesh pixels -- is the film's own, not that of a window around the geometry."""
    from adapt_b200._lib import pack_scene
    from adapt_b200.dist import tile_partition
    from oracle.pt_oracle import OracleScene
    w, h = c["film"]["width"], c["film"]["height"]
    osc = OracleScene(pack_scene(e, a, o, c, seed=args.seed, integrator=args.integrator))
    cores = n_threads or (os.cpu_count() or 1)
    n_tiles = ((w + 31) // 32) * ((h + 31) // 32)
    # pilot: about 48 tiles spread over the film (run twice: the first call warms caches / the OpenMP pool)
    k_pilot = max(1, n_tiles // 48)
    pilot = tile_partition(w, h, 0, k_pilot)
    osc.render(1, pixel_list=pilot, n_threads=cores)
    t0 = time.time(); osc.render(1, cnt_start=1, pixel_list=pilot, n_threads=cores); dt = max(time.time() - t0, 1e-4)
    rate = len(pilot) / dt                                              # pixel-samples per second
    want = budget_s * rate                                             # pixel-samples the budget buys
    k = int(max(1, np.ceil(w * h / max(want, 1.0))))                   # every k-th tile at 1 spp ...
    k = min(k, max(1, n_tiles // 16))
    sample = tile_partition(w, h, 0, k)
    spp = int(max(1, min(64, want / max(len(sample), 1))))             # ... or several spp when the whole film is cheap
    desc = f"every {k}-th 32x32 tile of the whole film ({len(sample)} of {w * h} pixels)" if k > 1 else f"whole film ({w * h} pixels)"
    return osc, sample, desc, cores, spp


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (here: the CPU oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    e, a, o, c = load_workload(args.workload, args.width, args.height, args.max_bounce)
    osc, sample, desc, cores, spp = cpu_reference_leg(args, e, a, o, c, budget_s=args.cpu_budget / max(args.steps, 1))
    for _ in range(args.warmup):
        osc.render(1, pixel_list=sample[: max(1024, len(sample) // 16)], n_threads=cores)
    rays = 0; paths = 0
    t0 = time.time()
    for k in range(args.steps):
        _, cn = osc.render(spp, cnt_start=k * spp, pixel_list=sample, n_threads=cores)
        rays += cn["rays_closest"]; paths += cn["paths"]
    dt = time.time() - t0
    w, h = c["film"]["width"], c["film"]["height"]
    value = rays / dt / 1e6
    sample_desc = f"{desc} x {spp} spp per step"
    line = {
        "impl": "reference", "metric": "Mrays/s (closest-hit rays: primary + secondary)", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, c, world),
        "run": {"spp_per_step": spp, "sample": sample_desc},
        "spp_per_s": paths / dt / (w * h),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def time_renderer(rdr, stream, fb, rank, local_rank, world, spp_step, steps, warmup, sample_clocks=True):
    """Device-timed loop (scene and film resident in HBM) and end-to-end loop (host buffers, H2D + D2H inside the timed region) of one
    renderer; CUDA events on the stream the library launches on, barrier + synchronize on both sides, max over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from adapt_b200.dist import reduce_framebuffer

    def step(k):
        if world > 1:
            rdr.reset_accumulation(k * spp_step)       # per-step film, cleared on the device: the reduce below is in place
        rdr.render_batch(spp_step)
        rdr.synchronize()
        if world > 1:
            reduce_framebuffer(fb, dst=0)              # NCCL sum over NVLink; pixels a rank does not own are exactly zero

    for k in range(warmup):
        step(k)
    barrier()
    rdr.stats(reset=True)
    sampler = ClockSampler(local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(steps):
        step(warmup + k)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.result() if sampler else None
    st = rdr.stats()
    tot = torch.tensor([ms, float(st["rays_closest"]), float(st["rays_shadow"]), float(st["paths"]), float(st["kernel_launches"]),
                        float(st["ms_closest"]), float(st["iterations"]), float(st["ms_logic"]), float(st["ms_shadow"])],
                       dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        ms = float(mx[0])
    tot = tot.cpu().numpy()

    # ---- e2e: checkpoint-style public API with host buffers.  Every step: H2D of a pinned (w,h,3) accumulation (load_check_point),
    # render, [NCCL reduce], D2H of the film into page-locked memory (pixels.to_numpy: mean formed on the device).
    w, h = rdr.w, rdr.h
    nbytes = w * h * 3 * 4
    host_in = torch.zeros((w, h, 3), dtype=torch.float32).pin_memory()
    host_np = host_in.numpy()
    ck = rdr.get_check_point()
    # one untimed pass through the same calls: the page-locked staging buffer of to_numpy() is allocated on first use
    ck["accumulation"] = host_np; ck["counter"] = 0
    rdr.load_check_point(ck); rdr.render_batch(1); rdr.pixels.to_numpy(copy=False)
    rdr.stats(reset=True)
    barrier()
    ev0.record(stream)
    for k in range(steps):
        ck["accumulation"] = host_np
        ck["counter"] = k * spp_step
        rdr.load_check_point(ck)                       # H2D of the pinned accumulation buffer
        rdr.render_batch(spp_step)
        if world > 1:
            rdr.synchronize(); reduce_framebuffer(fb, dst=0)       # in place: the accumulation is reloaded at the top of every step
        out = rdr.pixels.to_numpy(copy=False)          # D2H into page-locked memory (sync point); rank 0 holds the whole film
    ev1.record(stream)
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    e2e_st = rdr.stats()
    e2e_t = torch.tensor([e2e_ms, float(e2e_st["rays_closest"])], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        mx = e2e_t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(e2e_t, op=dist.ReduceOp.SUM)
        e2e_ms = float(mx[0])
    assert np.isfinite(out).all()
    return dict(ms=ms, tot=tot, st=st, clocks=clocks, e2e_ms=e2e_ms, e2e_rays=float(e2e_t[1]), nbytes=nbytes)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from adapt_b200.dist import device_tensor_view, init_process_group, tile_partition
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
    # weak scaling (default): every GPU renders spp_per_step x N samples of its 1/N of the film -- per-GPU work is fixed;
    # strong scaling: the step is spp_per_step samples of the whole film whatever N is
    spp_step = args.spp_per_step * (world if args.scaling == "weak" else 1)
    m = time_renderer(rdr, stream, fb, rank, local_rank, world, spp_step, args.steps, args.warmup)
    ms, tot, st, clocks, e2e_ms = m["ms"], m["tot"], m["st"], m["clocks"], m["e2e_ms"]
    rays_closest, rays_shadow, paths, launches = tot[1], tot[2], tot[3], tot[4]
    value = rays_closest / (ms * 1e-3) / 1e6
    e2e_value = m["e2e_rays"] / (e2e_ms * 1e-3) / 1e6
    nbytes = m["nbytes"]
    bvh_info = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in rdr.bvh_export(arrays=False).items()
                if k in ("builder", "n_nodes", "depth", "build_ms")}

    # ---- a second workload in the same process (N = 1 only): the 500k-triangle material-orb scene of configs[3], the scene class the
    # reference quotes its interactive frame rate on (README.md:34) -- a driver-side number for the ">= 1 Gray/s on 500k triangles" target
    also = None
    if world == 1 and args.also and args.also != args.workload and not args.no_cpu:
        try:
            rdr.close()
            e2, a2, o2, c2 = load_workload(args.also)
            rdr2 = Renderer(e2, a2, o2, c2, seed=args.seed, device_id=local_rank, pool_size=args.pool)
            rdr2.set_stream(stream.cuda_stream)
            p2, n2 = rdr2.accum_device_ptr()
            m2 = time_renderer(rdr2, stream, device_tensor_view(p2, n2, local_rank), 0, local_rank, 1, args.also_spp, 3, 3, sample_clocks=False)
            also = {args.also: {"config": workload_config(argparse.Namespace(**{**vars(args), "workload": args.also}), c2, 1)["workload"],
                                "value": m2["tot"][1] / (m2["ms"] * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": m2["ms"] / 3,
                                "spp_per_step": args.also_spp, "steps": 3, "warmup": 3,
                                "e2e": {"value": m2["e2e_rays"] / (m2["e2e_ms"] * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": m2["e2e_ms"] / 3},
                                "stage_ms_per_step": {"logic": m2["tot"][7] / 3, "trace": (m2["tot"][5] + m2["tot"][8]) / 3}}}
            rdr2.close()
        except Exception as ex:                        # the headline line must not depend on the extra workload
            also = {args.also: {"error": repr(ex)}}

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
            osc, sample, desc, cores, cspp = cpu_reference_leg(args, e, a, o, c, budget_s=args.cpu_budget)
            t0 = time.time(); _, cn = osc.render(cspp, pixel_list=sample, n_threads=cores); dt = time.time() - t0
            cpu = {"value": cn["rays_closest"] / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": f"{desc} x {cspp} spp, {dt:.1f} s"}
            if cn["nodes_visited"]:
                nbar_node = cn["nodes_visited"] / cn["rays_closest"]; nbar_prim = cn["prims_tested"] / cn["rays_closest"]
                nbar_node_s = cn["nodes_shadow"] / max(cn["rays_shadow"], 1); nbar_prim_s = cn["prims_shadow"] / max(cn["rays_shadow"], 1)
        # ---- roofline of the dominant kernel (k_trace: both ray streams of an iteration in one launch).  Three figures:
        #   frac         DRAM side: ncu dram bytes per launch / that capture's own duration / measured HBM peak (profiles/ncu_summary.json,
        #                captured on this workload with the shipped defaults); without a capture: the compulsory queue bytes / live duration
        #   frac_queue_only  compulsory HBM bytes (48 B per closest-hit ray: 32 B ray in + 16 B hit out; 60 B per shadow ray) / live duration
        #   frac_logical     SURVEY 8(d)'s logical figure: the bytes the REFERENCE's unordered traversal would touch (36 B per node, 104 B per
        #                    primitive, oracle counters) / live duration -- served by L1/L2 here, so it is not an HBM fraction
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
        logical = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        queue_gbs = queue_bytes_per_launch / (avg_ms * 1e-3) / 1e9
        kern = "k_trace" if fused else "k_closest"
        # k_logic: 80 B state read + 64 B state write per live slot (the colour word stays in HBM, the RNG state rides in two spare words),
        # 48 B per shadow ray written
        pool_slots = int(st.get("pool_slots") or args.pool or int(os.environ.get("ADAPT_POOL", 0)) or 0)
        lanes = int(st.get("lanes") or 1)          # a handle runs `lanes` pools side by side; a launch covers one of them
        logic_bytes = pool_slots / lanes * 144.0 + shadow_per_launch * 48.0
        # ncu captures (tools/profile_summary.py): DRAM bytes, duration and ray counts of the SAME launches
        cap = cap_logic = None
        if args.workload == "bunny90k" and not args.width and args.integrator == "pt":
            try:
                summ = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
                cap, cap_logic = summ.get(kern), summ.get("k_logic")
            except Exception:
                pass
        traffic = cap["dram_bytes_per_launch"] if cap else None
        if cap:
            achieved = cap["dram_bytes_per_launch"] / (cap["duration_us"] * 1e-6) / 1e9
            frac_source = f"ncu dram bytes / capture duration (profiles/ncu_summary.json, session {cap.get('session')}: cold-cache, serialised launches, the kernel alone at nine resident blocks per SM; in the timed region two lanes overlap and k_trace runs six blocks per SM beside the other lane's k_logic, so avg_launch_ms is longer than the capture's duration)"
        else:
            achieved = queue_gbs
            frac_source = "compulsory queue bytes / live launch duration (no ncu capture for this workload)"
        logic_ms = tot[7] / iters / world
        line = {
            "metric": "Mrays/s (closest-hit rays: primary + secondary)", "value": value, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, c, world),
            "run": {"spp_per_step": spp_step, "pool_slots": pool_slots, "lanes": lanes, "bvh": bvh_info},
            "spp_per_s": paths / (w * h) / (ms * 1e-3),
            "mrays_shadow_per_s": rays_shadow / (ms * 1e-3) / 1e6,
            "paths_per_s": paths / (ms * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            # event-timed per launch and summed: with two lanes the launches of one lane run WHILE the other lane's do, so the stages add up
            # to more than the step (and avg_launch_ms below is the wall duration of a launch that shares the SMs with the other lane's)
            "stage_ms_per_step": {"logic": tot[7] / world / args.steps, "shadow": tot[8] / world / args.steps, "closest": tot[5] / world / args.steps,
                                  "overlapping_lanes": lanes},
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "frac_source": frac_source, "traffic": traffic,
                         "traffic_capture": ({k: cap.get(k) for k in ("duration_us", "launches_captured", "rays_in_launch", "shadow_rays_in_launch", "session")} if cap else None),
                         "peak_source": peak_src,
                         "achieved_queue_only": queue_gbs, "frac_queue_only": queue_gbs / peak_gbs, "bytes_per_ray_queue_only": b_queue,
                         "achieved_logical": logical, "frac_logical": logical / peak_gbs,
                         "bytes_per_ray": b_ray, "bytes_per_shadow_ray": b_ray_s if fused else None,
                         "ref_nodes_per_ray": nbar_node, "ref_prims_per_ray": nbar_prim,
                         "ref_nodes_per_shadow_ray": nbar_node_s, "ref_prims_per_shadow_ray": nbar_prim_s, "avg_launch_ms": avg_ms,
                         "rays_per_launch": rays_per_launch, "shadow_rays_per_launch": shadow_per_launch if fused else None,
                         "note": "k_trace is latency / issue bound, not HBM bound: the BVH is served by L1/L2, only the ray and hit queues stream through HBM. "
                                 "frac = DRAM side (ncu bytes over the capture's own duration); frac_logical = bytes the reference's unordered traversal would touch "
                                 "(SURVEY 8(d)), not an HBM figure; frac_queue_only = compulsory queue bytes over the live launch duration",
                         # the whole step: compulsory HBM bytes of both kernels (pool state + queues) over the device-timed step
                         # (`iters` counts the launches of all ranks; achieved / frac are per GPU)
                         "step": {"compulsory_bytes_per_step": (logic_bytes + queue_bytes_per_launch) * iters / args.steps,
                                  "achieved": (logic_bytes + queue_bytes_per_launch) * iters / world / (ms * 1e-3) / 1e9,
                                  "frac": (logic_bytes + queue_bytes_per_launch) * iters / world / (ms * 1e-3) / 1e9 / peak_gbs},
                         # second kernel of the iteration: streams the whole path pool (HBM-bound by construction)
                         "k_logic": {"achieved": logic_bytes / max(logic_ms * 1e-3, 1e-9) / 1e9,
                                     "frac": logic_bytes / max(logic_ms * 1e-3, 1e-9) / 1e9 / peak_gbs,
                                     "bytes_per_launch": logic_bytes, "avg_launch_ms": logic_ms,
                                     "traffic": cap_logic["dram_bytes_per_launch"] if cap_logic else None,
                                     "frac_dram": (cap_logic["dram_bytes_per_launch"] / (cap_logic["duration_us"] * 1e-6) / 1e9 / peak_gbs) if cap_logic else None}},
            "cpu_baseline": cpu,
        }
        if also:
            line["also"] = also
        emit(line)
    try:
        rdr.close()
    except Exception:
        pass
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
    ap.add_argument("--spp-per-step", type=int, default=256,
                    help="samples per pixel one step enqueues (a step ends with adapt_sync, i.e. drains the path pool: ~4 ms of ragged tail, 7 %% of a 32-spp step on bunny90k, 2 %% of a 128-spp one, 1 %% of a 256-spp one)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--max-bounce", type=int, default=None)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU-oracle work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the --also workload (A/B and profiling runs)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = spp_per_step x N samples per step (per-GPU work fixed), strong = spp_per_step samples per step whatever N")
    ap.add_argument("--also", default="orb500k", help="second workload timed in the same process at N = 1 ('' = none)")
    ap.add_argument("--also-spp", type=int, default=256)
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
