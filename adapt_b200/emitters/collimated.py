"""Collimated (laser-like) emitter (reference emitters/collimated.py:22-61)."""
import xml.etree.ElementTree as xet

import numpy as np

from ..parsers.general_parser import get, vec3d_parse
from ..renderer.constants import COLLIMATED_SOURCE, INV_PI
from .abtract_source import LightSource


class CollimatedSource(LightSource):
    def __init__(self, elem: xet.Element = None):
        super().__init__(elem)
        point_elems = elem.findall("point")
        assert len(point_elems) >= 2
        self.dir = np.float32([0, 0, 1])
        self.pos = np.zeros(3, np.float32)
        self.radius = 0.0
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
            if float_elem.get("name") == "radius":
                self.radius = max(0.0, get(float_elem, "value", float))
        self.inv_area = 1 if self.radius == 0 else INV_PI / (self.radius * self.radius)

    def export(self) -> np.ndarray:
        bool_bits = int(self.radius == 0) + 0x02 + (int(self.in_free_space) << 4)
        return self._record(COLLIMATED_SOURCE, bool_bits, pos=self.pos, dirv=self.dir, r=self.radius)
