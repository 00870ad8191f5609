"""Spot emitter (reference emitters/spot.py:18-52): positional delta, cone given by cos(half-angle)."""
import xml.etree.ElementTree as xet

import numpy as np

from ..parsers.general_parser import get, vec3d_parse
from ..renderer.constants import DEG2RAD, SPOT_SOURCE
from .abtract_source import LightSource


class SpotSource(LightSource):
    def __init__(self, elem: xet.Element = None):
        super().__init__(elem)
        point_elems = elem.findall("point")
        assert len(point_elems) >= 2
        self.dir = np.float32([0, 0, 1])
        self.pos = np.zeros(3, np.float32)
        self.half_cos = np.cos(15.0 * DEG2RAD)
        for point_elem in point_elems:
            name = point_elem.get("name")
            if name in {"position", "pos"}:
                self.pos = vec3d_parse(point_elem)
            elif name in {"direction", "dir"}:
                self.dir = vec3d_parse(point_elem)
                norm = np.linalg.norm(self.dir)
                if norm < 1e-5:
                    raise ValueError(f"Direction of collimated source <{self.id}> is ill-conditioned.")
                self.dir /= norm
        for float_elem in elem.findall("float"):
            if float_elem.get("name") == "half-angle":
                self.half_cos = np.cos(max(1e-3, get(float_elem, "value", float)) * DEG2RAD)
        self.inv_area = 1.0

    def export(self) -> np.ndarray:
        bool_bits = 0x01 + (int(self.in_free_space) << 4)
        return self._record(SPOT_SOURCE, bool_bits, pos=self.pos, dirv=self.dir, r=self.half_cos)
