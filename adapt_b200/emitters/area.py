"""Area emitter attached to a mesh / sphere (reference emitters/area.py:15-27)."""
import xml.etree.ElementTree as xet

import numpy as np

from ..renderer.constants import AREA_SOURCE
from .abtract_source import LightSource


class AreaSource(LightSource):
    def __init__(self, elem: xet.Element):
        super().__init__(elem)
        self.attached = True

    def export(self) -> np.ndarray:
        bool_bits = (int(self.in_free_space) << 4) | 0x04
        return self._record(AREA_SOURCE, bool_bits)
