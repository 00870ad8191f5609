"""Multi-GPU plumbing: image-tile partition + framebuffer reduce over torch.distributed.

The reference is single-device.  Here every pixel-sample is independent and the RNG is
keyed by (pixel, sample), so the film is split into interleaved tiles, each rank renders its own
pixels with a full copy of the scene, and the only exchange is ONE sum-reduce of the (w,h,3) fp32
accumulation buffer (NCCL over NVLink on GPUs, gloo in the CPU tests).  Pixels a rank does not own
stay exactly zero in its buffer, so the sum is a gather and the N-GPU image equals the 1-GPU image.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

__all__ = ["tile_partition", "auto_tile", "dist_env", "init_process_group", "reduce_framebuffer", "device_tensor_view"]


def tile_partition(w: int, h: int, rank: int, world: int, tile: int = 32,
                   window: Optional[Tuple[int, int, int, int]] = None) -> np.ndarray:
    """Film indices ``i * h + j`` owned by ``rank``: tile k (row-major over tiles of ``tile`` x ``tile``
    pixels inside ``window`` = (start_x, end_x, start_y, end_y)) goes to rank ``k % world``.  Inside a
    tile pixels are listed in 4 x 8 blocks so one warp's worth of work items is a compact patch."""
    sx, ex, sy, ey = window if window is not None else (0, w, 0, h)
    out = []
    k = 0
    for ti in range(sx, ex, tile):
        for tj in range(sy, ey, tile):
            if k % world == rank:
                ii = np.arange(ti, min(ti + tile, ex))
                jj = np.arange(tj, min(tj + tile, ey))
                # 4 x 8 blocks
                bi = (ii - ti) // 4
                bj = (jj - tj) // 8
                I, J = np.meshgrid(ii, jj, indexing="ij")
                BI, BJ = np.meshgrid(bi, bj, indexing="ij")
                key = (BI * 1024 + BJ) * 64 + ((I - ti) % 4) * 8 + ((J - tj) % 8)
                order = np.argsort(key.reshape(-1), kind="stable")
                out.append((I.reshape(-1) * h + J.reshape(-1))[order])
            k += 1
    if not out:
        return np.zeros((0,), np.int32)
    return np.concatenate(out).astype(np.int32)


def auto_tile(w: int, h: int, world: int, window: Optional[Tuple[int, int, int, int]] = None, tile: int = 32) -> int:
    """Largest tile edge <= ``tile`` (halving down to 4) that leaves every rank at least one tile of the film / crop window."""
    sx, ex, sy, ey = window if window is not None else (0, w, 0, h)
    while tile > 4 and (-(-(ex - sx) // tile)) * (-(-(ey - sy) // tile)) < world:
        tile //= 2
    return tile


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: Optional[str] = None):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


class _DevArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def device_tensor_view(ptr: int, n_floats: int, device_id: int):
    """Zero-copy torch view of the renderer's device framebuffer (adapt_accum_device_ptr)."""
    import torch
    return torch.as_tensor(_DevArray(ptr, n_floats), device=torch.device("cuda", device_id))


def reduce_framebuffer(buf, dst: int = 0, all_ranks: bool = False):
    """In-place sum of the per-rank accumulation buffers (a torch tensor on cuda -> NCCL, on cpu -> gloo)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return buf
    if all_ranks:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM)
    return buf
