#!/usr/bin/env python
"""Rendering main executable -- same flags and flow as the reference's render.py (65-166), driving the
B200-native `pt` / `vpt` renderers.  Usage (identical to AdaPT):

    python render.py --scene cbox --name cbox.xml --type pt --iter_num 64 --no_gui

Differences: there is no GUI (the loop always runs head-less), `--arch` only accepts the CUDA back end, and
`--gpus N` (under torchrun) tile-splits the film with one NCCL framebuffer reduce at the end.
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
    rdr = rdr_mapping[opts.type](emitter_configs, array_info, all_objs, configs, seed=opts.seed, device_id=local_rank,
                                 pixel_list=pixel_list, max_bounce=opts.max_bounce)
    max_iter_num = opts.iter_num if opts.iter_num > 0 else configs.get("iter_num", 2000)
    max_iter_num += 1                                    # the reference's head-less loop renders iter_num + 1 spp (render.py:81,118)
    max_bounce = rdr.max_bounce
    CONSOLE.log(f"Path Tracing with {max_bounce} bounce(s)")
    if opts.load:
        chkpt_path = os.path.join(folder_path(opts.chkpt_path), f"{opts.img_name}-{opts.name[:-4]}-{opts.type}.pkl")
        with open(chkpt_path, "rb") as file:
            rdr.load_check_point(pickle.load(file))
    CONSOLE.rule()
    batch = opts.spp_per_launch if opts.spp_per_launch > 0 else (opts.save_iter if opts.save_iter > 0 else max_iter_num)
    done = 0
    try:
        while done < max_iter_num:
            if opts.save_iter > 0 and done % opts.save_iter == 0 and world == 1:
                save_check_point(rdr.get_check_point(), opts)
            n = min(batch, max_iter_num - done)
            rdr.render_batch(n)                          # == n calls of rdr.render(...) in the reference loop
            done += n
            if opts.output_freq > 0 and done % opts.output_freq == 0 and world == 1:
                imwrite(rdr.pixels.to_numpy(), f"{output_folder}img_{done:05d}.{opts.img_ext}")
    except KeyboardInterrupt:
        if opts.save_iter > 0 and world == 1:
            save_check_point(rdr.get_check_point(), opts)
        CONSOLE.log(":ok: Quit on Keyboard interruptions")
    rdr.summary()
    if world > 1:
        import torch
        rdr.synchronize()
        ptr, n = rdr.accum_device_ptr()
        reduce_framebuffer(device_tensor_view(ptr, n, local_rank), dst=0)
        torch.cuda.synchronize()
    if opts.profile:
        CONSOLE.rule()
        CONSOLE.print(rdr.stats())
    if rank == 0:
        image = apply_watermark(rdr, opts.normalize, True, not opts.no_watermark)
        if opts.save_hdr:
            np.save(f"{output_folder}{opts.img_name}-{opts.name[:-4]}-{opts.type}.npy", rdr.pixels.to_numpy())
        if not opts.no_save_fig:
            imwrite(image, f"{output_folder}{opts.img_name}-{opts.name[:-4]}-{opts.type}.{opts.img_ext}")
    return rdr


if __name__ == "__main__":
    main()
