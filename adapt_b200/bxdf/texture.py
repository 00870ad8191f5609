"""Host-side texture descriptor (reference bxdf/texture.py:35-101, `Texture_np`) re-hosted without Taichi.

``export()`` returns one packed ``TEXTURE_DTYPE`` record -- the C-ABI ``adapt_texture`` of include/adapt_b200.h, field for
field the reference's Taichi ``Texture`` struct (bxdf/texture.py:103-112).  The bilinear lookup itself (``Texture.query``,
:114-139) is device code: csrc/pt_shade.cuh ``texture_query``.

Reference behaviour kept on purpose: the image path is used as written (relative to the working directory; here the XML's
directory is tried as a fallback); images larger than ``max_size`` are resized; ``bump`` images get their G and B
channels swapped (y-up convention); checkerboard textures are parsed but the reference never implemented their lookup
(``query`` would read a w = h = 0 image), so a scene that attaches one raises ``NotImplementedError`` here.
"""
import os
import xml.etree.ElementTree as xet

import numpy as np

from ..parsers.general_parser import get, rgb_parse
from ..utils.tools import CONSOLE

__all__ = ["Texture_np", "TEXTURE_DTYPE", "TEX_INVALID"]

TEX_INVALID = -255

# C layout of adapt_texture (8 x 4 bytes)
TEXTURE_DTYPE = np.dtype([
    ("type", np.int32), ("off_x", np.int32), ("off_y", np.int32), ("w", np.int32), ("h", np.int32),
    ("scale_u", np.float32), ("scale_v", np.float32), ("_pad", np.int32),
], align=False)
assert TEXTURE_DTYPE.itemsize == 32


class Texture_np:
    MODE_IMAGE = 0
    MODE_CHECKER = 1

    def __init__(self, elem: xet.Element, max_size=2048, directory: str = ""):
        self.tag = elem.get("tag", "albedo")
        self.max_size = max_size
        self.id = elem.get("id")
        self.type = elem.get("type")
        self.c1 = np.zeros(3)
        self.c2 = np.ones(3)
        self.scale_u = 1.0
        self.scale_v = 1.0
        self.off_x = 0
        self.off_y = 0
        self.h, self.w = 0, 0
        self.texture_img = None
        if self.type == "checkerboard":
            self.mode = Texture_np.MODE_CHECKER
            rgb_nodes = elem.findall("rgb")
            if len(rgb_nodes) > 0:
                self.c1 = rgb_parse(rgb_nodes[0])
                if len(rgb_nodes) > 1:
                    self.c2 = rgb_parse(rgb_nodes[1])
        else:
            import cv2 as cv
            self.mode = Texture_np.MODE_IMAGE
            file_path = elem.find("string").get("value")
            if not os.path.exists(file_path) and os.path.exists(os.path.join(directory, file_path)):
                file_path = os.path.join(directory, file_path)
            if not os.path.exists(file_path):
                raise ValueError(f"Texture image input path '{file_path}' does not exist.")
            self.texture_path = file_path
            texture_img = cv.cvtColor(cv.imread(file_path), cv.COLOR_BGR2RGB)
            self.h, self.w, _ = texture_img.shape
            if self.h > max_size or self.w > max_size:
                self.w = min(self.w, max_size)
                self.h = min(self.h, max_size)
                texture_img = cv.resize(texture_img, (self.w, self.h))
            self.texture_img = texture_img.astype(np.float32) / 255.0
            if self.tag == "bump":
                self.texture_img[..., [1, 2]] = self.texture_img[..., [2, 1]]
        for float_n in elem.findall("float"):
            name = float_n.get("name")
            if name in {"scale_u", "scale_v"}:
                setattr(self, name, get(float_n, "value"))
            else:
                CONSOLE.log(f"[yellow]:warning: Warning: <{name}> not used in loading textures")

    def export(self) -> np.ndarray:
        rec = np.zeros((), dtype=TEXTURE_DTYPE)
        rec["type"] = Texture_np.MODE_CHECKER if self.type == "checkerboard" else Texture_np.MODE_IMAGE
        rec["off_x"], rec["off_y"], rec["w"], rec["h"] = self.off_x, self.off_y, self.w, self.h
        rec["scale_u"], rec["scale_v"] = self.scale_u, self.scale_v
        return rec

    @staticmethod
    def default() -> np.ndarray:
        rec = np.zeros((), dtype=TEXTURE_DTYPE)
        rec["type"] = TEX_INVALID
        rec["scale_u"] = rec["scale_v"] = 1.0
        return rec

    def __repr__(self) -> str:
        return f"<Texture '{self.id}': {self.off_x}, {self.off_y}, {self.w}, {self.h}>"
