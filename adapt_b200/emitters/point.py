"""Point emitter (reference emitters/point.py:19-28): positional delta, bool_bits = 0x01 | in_free_space<<4."""
import xml.etree.ElementTree as xet

import numpy as np

from ..parsers.general_parser import vec3d_parse
from ..renderer.constants import POINT_SOURCE
from .abtract_source import LightSource


class PointSource(LightSource):
    def __init__(self, elem: xet.Element = None):
        super().__init__(elem)
        pos_elem = elem.find("point")
        assert pos_elem is not None
        self.pos: np.ndarray = vec3d_parse(pos_elem)

    def export(self) -> np.ndarray:
        bool_bits = 0x01 + (int(self.in_free_space) << 4)
        rec = self._record(POINT_SOURCE, bool_bits, pos=self.pos)
        rec["inv_area"] = 0.0      # TaichiSource default for point sources (field left unset, point.py:28)
        return rec
