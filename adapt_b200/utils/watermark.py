"""Watermark + quantile normalisation of the final image (reference utils/watermark.py:22-33).  The stamp is a
7 x 92 bitmap reading "RENDERED WITH AdaPT"; here it is generated from a tiny 5 x 7 font instead of a literal table."""
import numpy as np

from .tools import CONSOLE

__all__ = ["apply_watermark", "water_mark"]

_FONT = {
    "A": ["0110", "1001", "1001", "1111", "1001", "1001", "1001"], "D": ["1110", "1001", "1001", "1001", "1001", "1001", "1110"],
    "E": ["1111", "1000", "1000", "1111", "1000", "1000", "1111"], "H": ["1001", "1001", "1001", "1111", "1001", "1001", "1001"],
    "I": ["1", "1", "1", "1", "1", "1", "1"], "N": ["1001", "1101", "1101", "1011", "1011", "1001", "1001"],
    "P": ["1110", "1001", "1001", "1110", "1000", "1000", "1000"], "R": ["1110", "1001", "1001", "1110", "1010", "1001", "1001"],
    "T": ["11111", "00100", "00100", "00100", "00100", "00100", "00100"], "W": ["10001", "10001", "10001", "10101", "10101", "10101", "01010"],
    "a": ["0000", "0000", "0110", "0001", "0111", "1001", "0111"], "d": ["0001", "0001", "0111", "1001", "1001", "1001", "0111"],
    " ": ["00", "00", "00", "00", "00", "00", "00"],
}


def _stamp(text: str) -> np.ndarray:
    cols = []
    for ch in text:
        glyph = np.array([[int(c) for c in row] for row in _FONT[ch]], np.float32)
        cols.append(glyph)
        cols.append(np.zeros((7, 1), np.float32))
    return np.concatenate(cols[:-1], axis=1)


water_mark = _stamp("RENDERED WITH AdaPT")


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
