#!/usr/bin/env python
"""Rendering main executable -- same flags and flow as the reference's render.py (65-166), driving the
B200-native `pt` / `vpt` renderers.  Usage (identical to AdaPT):

    python render.py --scene cbox --name cbox.xml --type pt --iter_num 64 --no_gui

Differences: there is no GUI (the loop always runs head-less), `--arch` only accepts the CUDA back end, and
`--gpus N` renders on N GPUs from this one process (film tiles per device, gathered over NVLink peer loads); under torchrun (one
process per GPU) the ranks tile-split the film and one NCCL framebuffer reduce assembles it at the end.
"""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from adapt_b200.parsers.opts import get_options                     # noqa: E402
from adapt_b200.parsers.xml_parser import scene_parsing             # noqa: E402
from adapt_b200.utils.tools import CONSOLE, folder_path             # noqa: E402
from adapt_b200.utils.watermark import apply_watermark              # noqa: E402


def imwrite(image: np.ndarray, path: str):
    """ti.tools.imwrite semantics: (w, h, 3) float image indexed [x, y] with y up -> rows top-down, clipped to [0, 1]."""
    import cv2
    img = np.clip(np.asarray(image, np.float32), 0.0, 1.0)
    img = np.flipud(np.transpose(img, (1, 0, 2)))
    cv2.imwrite(path, (img[..., ::-1] * 255.0 + 0.5).astype(np.uint8))


def save_check_point(chkpt: dict, opts):
    chkpt_path = os.path.join(folder_path(opts.chkpt_path), f"{opts.img_name}-{opts.name[:-4]}-{opts.type}.pkl")
    with open(chkpt_path, "wb") as file:
        pickle.dump(chkpt, file, protocol=pickle.HIGHEST_PROTOCOL)


def _progress_bar():
    """The reference's bar (render.py:103-113), counting samples per pixel."""
    from rich.progress import BarColumn, MofNCompleteColumn, Progress, SpinnerColumn, TextColumn, TimeElapsedColumn, TimeRemainingColumn
    from adapt_b200.utils.rich_utils import ItersPerSecColumn
    return Progress(TextColumn(":movie_camera: Rendering :movie_camera:"), SpinnerColumn(), BarColumn(), MofNCompleteColumn(),
                    ItersPerSecColumn(suffix="spp/s"), TextColumn(" | ETA: "), TimeRemainingColumn(elapsed_when_finished=True),
                    TextColumn(" | elasped: "), TimeElapsedColumn())


def write_metrics(path: str, opts, rdr, stats: dict, seconds: float, spp: int, world: int, n_devices: int = 1):
    """Machine-readable summary of the run next to the image (SURVEY 5 "metrics / logging"): throughput, per-stage device time,
    counters.  Under torchrun the figures are rank 0's share of the film; with several devices behind one handle the counters are sums
    and the stage times those of the slowest device."""
    import json
    pool = stats.get("pool_slots", 0) // max(1, stats.get("lanes", 1)) // max(1, n_devices)
    logic_ms, trace_ms = stats["ms_logic"], stats["ms_closest"] + stats["ms_shadow"]
    out = {
        "scene": opts.scene, "name": opts.name, "type": opts.type, "film": [rdr.w, rdr.h], "max_bounce": rdr.max_bounce, "spp": spp,
        "gpus": world * n_devices, "processes": world, "seconds": seconds, "spp_per_s": spp / max(seconds, 1e-9),
        "mrays_per_s": stats["rays_closest"] / max(seconds, 1e-9) / 1e6, "mrays_shadow_per_s": stats["rays_shadow"] / max(seconds, 1e-9) / 1e6,
        "paths": stats["paths"], "rays_closest": stats["rays_closest"], "rays_shadow": stats["rays_shadow"],
        "iterations": stats["iterations"], "kernel_launches": stats["kernel_launches"],
        "stage_ms": {"logic": logic_ms, "trace": trace_ms},
        # k_logic streams the pool: 80 B read + 64 B written per slot and launch
        "k_logic_hbm_gbs": (stats["iterations"] * pool * 144.0 / (logic_ms * 1e-3) / 1e9) if logic_ms > 0 else None,
        "bvh": {k: v for k, v in rdr.bvh_export(arrays=False).items() if k in ("builder", "n_nodes", "depth", "build_ms")},
    }
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


def main(argv=None):
    opts = get_options(argv=argv)
    from adapt_b200.renderer.vanilla_renderer import Renderer
    from adapt_b200.renderer.vpt import VolumeRenderer
    rdr_mapping = {"pt": Renderer, "vpt": VolumeRenderer}             # render.py:33 (bdpt / ao are outside the hot-path scope)
    if opts.type not in rdr_mapping:
        raise NotImplementedError(f"--type {opts.type}: `pt` and `vpt` (homogeneous media) run on the device in this build; nothing falls back")
    input_folder = os.path.join(opts.input_path, opts.scene)
    emitter_configs, array_info, all_objs, configs = scene_parsing(input_folder, opts.name)
    output_folder = folder_path(opts.output_path)
    # multi-GPU: one process per GPU (torchrun); each rank owns interleaved film tiles
    from adapt_b200.dist import auto_tile, device_tensor_view, init_process_group, reduce_framebuffer, tile_partition
    rank, local_rank, world = init_process_group() if int(os.environ.get("WORLD_SIZE", "1")) > 1 else (0, 0, 1)
    film = configs["film"]
    pixel_list = None
    if world > 1:
        # the ranks split the crop window when the film has one (tracer_base.py:64-75), not the whole film
        w, h = film["width"], film["height"]
        window = None
        if film.get("crop_rx", 0) > 0 and film.get("crop_ry", 0) > 0:
            cx, cy, rx, ry = film.get("crop_x", 0), film.get("crop_y", 0), film["crop_rx"], film["crop_ry"]
            window = (max(0, cx - rx), min(w, cx + rx), max(0, cy - ry), min(h, cy + ry))
        pixel_list = tile_partition(w, h, rank, world, tile=auto_tile(w, h, world, window), window=window)
        if (opts.save_iter > 0 or opts.output_freq > 0) and rank == 0:
            CONSOLE.log("[yellow]--save_iter / --output_freq are ignored under torchrun (the film is only assembled at the end)")
    # --gpus N without torchrun: one process, one renderer over N devices (the library partitions the film and gathers it over NVLink
    # peer loads; checkpoints and --output_freq keep working because this process sees the whole film)
    device_ids = list(range(opts.gpus)) if (world == 1 and opts.gpus > 1) else None
    rdr = rdr_mapping[opts.type](emitter_configs, array_info, all_objs, configs, seed=opts.seed, device_id=local_rank,
                                 pixel_list=pixel_list, max_bounce=opts.max_bounce, device_ids=device_ids)
    max_iter_num = opts.iter_num if opts.iter_num > 0 else configs.get("iter_num", 2000)
    max_iter_num += 1                                    # the reference's head-less loop renders iter_num + 1 spp (render.py:81,118)
    max_bounce = rdr.max_bounce
    CONSOLE.log(f"Path Tracing with {max_bounce} bounce(s)")
    if opts.load:
        chkpt_path = os.path.join(folder_path(opts.chkpt_path), f"{opts.img_name}-{opts.name[:-4]}-{opts.type}.pkl")
        with open(chkpt_path, "rb") as file:
            rdr.load_check_point(pickle.load(file))
    CONSOLE.rule()
    # samples are enqueued in batches (adapt_render is asynchronous; a batch is `n` calls of rdr.render(...) in the reference loop).  The
    # bar advances when a batch has been handed out to the path pool, like the reference's bar advances when render() returns.
    batch = opts.spp_per_launch if opts.spp_per_launch > 0 else (opts.save_iter if opts.save_iter > 0 else max(1, (max_iter_num + 15) // 16))
    done = 0
    progress = _progress_bar() if (rank == 0 and os.environ.get("ADAPT_QUIET", "0") != "1") else None
    import time
    t_start = time.time()
    rdr.stats(reset=True)
    try:
        if progress is not None:
            progress.start()
            task = progress.add_task("", total=max_iter_num)
        while done < max_iter_num:
            if opts.save_iter > 0 and done % opts.save_iter == 0 and world == 1:
                save_check_point(rdr.get_check_point(), opts)
            n = min(batch, max_iter_num - done)
            rdr.render_batch(n)
            rdr.wait()
            done += n
            if progress is not None:
                progress.update(task, advance=n)
            if opts.output_freq > 0 and done % opts.output_freq == 0 and world == 1:
                imwrite(rdr.pixels.to_numpy(), f"{output_folder}img_{done:05d}.{opts.img_ext}")
    except KeyboardInterrupt:
        if opts.save_iter > 0 and world == 1:
            save_check_point(rdr.get_check_point(), opts)
        CONSOLE.log(":ok: Quit on Keyboard interruptions")
    finally:
        if progress is not None:
            progress.stop()
    rdr.summary()
    seconds = time.time() - t_start
    stats = rdr.stats()
    if world > 1:
        import torch
        rdr.synchronize()
        ptr, n = rdr.accum_device_ptr()
        reduce_framebuffer(device_tensor_view(ptr, n, local_rank), dst=0)
        torch.cuda.synchronize()
    if opts.profile:
        CONSOLE.rule()
        CONSOLE.print(stats)
    if rank == 0:
        write_metrics(f"{output_folder}{opts.img_name}-{opts.name[:-4]}-{opts.type}.metrics.json", opts, rdr, stats, seconds, done,
                      world, len(device_ids) if device_ids else 1)
        image = apply_watermark(rdr, opts.normalize, True, not opts.no_watermark)
        if opts.save_hdr:
            np.save(f"{output_folder}{opts.img_name}-{opts.name[:-4]}-{opts.type}.npy", rdr.pixels.to_numpy())
        if not opts.no_save_fig:
            imwrite(image, f"{output_folder}{opts.img_name}-{opts.name[:-4]}-{opts.type}.{opts.img_ext}")
    return rdr


if __name__ == "__main__":
    main()
