"""Watermark + quantile normalisation of the final image (reference utils/watermark.py:22-33).  The stamp is the
reference's 7 x 92 bitmap reading "RENDERED WITH AdaPT" (bit-packed here), so images carry the same pixels."""
import numpy as np

from .tools import CONSOLE

__all__ = ["apply_watermark", "water_mark"]

# The stamp of the reference (utils/watermark.py:13-20), one integer per row, most significant bit = leftmost column.
# Row 0 is the BOTTOM line of the lettering: the film is indexed [x, y] with y up and imwrite flips it.
_STAMP_WIDTH = 92
_STAMP_ROWS = (0x97a2e7a5ee05222425efa04, 0x94269425090522242529204, 0xa4269429090522242529204, 0xe7aa97b9e905223c3def384,
               0x942a9425090aa2242420244, 0x94329425090aa2242420244, 0xe7b2e7b9ee0aafa4182039f)
water_mark = np.float32([[(row >> (_STAMP_WIDTH - 1 - c)) & 1 for c in range(_STAMP_WIDTH)] for row in _STAMP_ROWS])


def apply_watermark(rdr, normalize: float = 0.0, verbose: bool = False, add_watermark: bool = True):
    """Same call contract as the reference: reads rdr.pixels (w, h, 3), crops when rdr.do_crop, optional
    quantile normalisation, stamp written at img[-w-1:-1, :h] (lands bottom-right after the imwrite transpose/flip)."""
    img = rdr.pixels.to_numpy()
    if rdr.do_crop:
        img = img[rdr.start_y:rdr.end_y, rdr.start_x:rdr.end_x, :]
    if verbose:
        CONSOLE.log(f"Pixel max value = {img.max():.3f}")
    if normalize > 0.9:
        img /= np.quantile(img, normalize)
    if not rdr.do_crop and add_watermark:
        h, w = water_mark.shape
        if img.shape[0] > w + 1 and img.shape[1] > h:
            img[-w - 1:-1, :h, :] += water_mark.T[..., None]
    return img
