"""Packs texture images of different sizes into one square atlas per map kind (reference parsers/texture_packing.py:32-81).

The reference delegates the rectangle packing to the third-party ``rectpack`` module (absent here); a shelf packer takes its
place.  The layout is free to differ: ``Texture.query`` wraps its coordinates into ``[0, w - 1) x [0, h - 1)`` before adding the
offsets (bxdf/texture.py:117-118), so a lookup never reads outside its own rectangle and the result does not depend on where
the rectangle sits.  Atlas sizes tried are the reference's (720, 1024, 2048, 3072; :26), smallest first.
"""
from typing import Dict, List, Tuple

import numpy as np

from ..bxdf.texture import Texture_np

SIZE2USE = [3072, 2048, 1024, 720]

__all__ = ["image_packer", "shelf_pack"]


def shelf_pack(rects: List[Tuple[int, int, int]], size: int):
    """rects: (w, h, id). Returns {id: (x, y)} or None when they do not fit a size x size bin (no rotation)."""
    placed = {}
    x = y = shelf_h = 0
    for w, h, rid in sorted(rects, key=lambda r: (-r[1], -r[0], r[2])):
        if w > size or h > size:
            return None
        if x + w > size:
            y += shelf_h
            x = shelf_h = 0
        if y + h > size:
            return None
        placed[rid] = (x, y)
        x += w
        shelf_h = max(shelf_h, h)
    return placed


def image_packer(textures: List[Texture_np]) -> Tuple[np.ndarray, Dict[str, Texture_np]]:
    rects = []
    for idx, texture in enumerate(textures):
        if texture.mode == Texture_np.MODE_CHECKER:
            continue
        h, w, _ = texture.texture_img.shape
        rects.append((w, h, idx))
    final_size = None
    for cur_size in reversed(SIZE2USE):
        placed = shelf_pack(rects, cur_size)
        if placed is not None:
            final_size = cur_size
            for rid, (x, y) in placed.items():
                textures[rid].off_x, textures[rid].off_y = x, y
            break
    if final_size is None:
        raise ValueError("Texture image packing failed, max size 3072 can not even satisfy.")
    result_image = np.zeros((final_size, final_size, 3), dtype=np.float32)
    result_dict = {}
    for texture in textures:
        if texture.texture_img is not None:
            sx, sy = texture.off_x, texture.off_y
            result_image[sy:sy + texture.h, sx:sx + texture.w, :] = texture.texture_img
        result_dict[texture.id] = texture
    return result_image, result_dict
