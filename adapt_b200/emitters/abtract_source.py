"""Host emitter descriptors (reference emitters/abtract_source.py:246-281).

``LightSource.export()`` returns a packed ``EMITTER_DTYPE`` record (C-ABI ``adapt_emitter``) whose
fields are those of the reference ``TaichiSource`` dataclass (:44-54): _type, obj_ref_id, bool_bits
(b0 pos-delta, b1 dir-delta, b2 area, b3 infinite, b4 in-free-space), intensity, dir, pos, inv_area,
r, emit_time.  The device methods (sample_hit :81-158, eval_le :210-218, solid_angle_pdf :220-232)
are in csrc/pt_shade.cuh."""
import xml.etree.ElementTree as xet

import numpy as np

from ..parsers.general_parser import rgb_parse
from ..utils.tools import CONSOLE

__all__ = ["LightSource", "EMITTER_DTYPE"]

EMITTER_DTYPE = np.dtype([
    ("type", np.int32), ("obj_ref_id", np.int32), ("bool_bits", np.int32),
    ("intensity", np.float32, 3), ("dir", np.float32, 3), ("pos", np.float32, 3),
    ("inv_area", np.float32), ("r", np.float32), ("emit_time", np.float32), ("_pad", np.float32),
], align=False)
assert EMITTER_DTYPE.itemsize == 64


class LightSource:
    def __init__(self, base_elem: xet.Element = None):
        self.intensity = np.ones(3, np.float32)
        if base_elem is not None:
            for rgb_elem in base_elem.findall("rgb"):
                name = rgb_elem.get("name")
                if name == "emission":
                    self.intensity = rgb_parse(rgb_elem)
                elif name == "scaler":
                    self.intensity *= rgb_parse(rgb_elem)
        else:
            CONSOLE.log("[yellow]:warning: Warning: default intializer should only be used in testing.")
        self.type: str = base_elem.get("type")
        self.id: str = base_elem.get("id")
        self.inv_area = 1.0
        self.attached = False
        self.in_free_space = True
        self.emit_time = 0.0
        bool_elem = base_elem.find("boolean")
        if bool_elem is not None and bool_elem.get("value").lower() == "false":
            self.in_free_space = False

    def _record(self, _type, bool_bits, pos=None, dirv=None, r=0.0) -> np.ndarray:
        rec = np.zeros((), dtype=EMITTER_DTYPE)
        rec["type"] = _type
        rec["obj_ref_id"] = -1
        rec["bool_bits"] = bool_bits
        rec["intensity"] = self.intensity
        if pos is not None:
            rec["pos"] = pos
        if dirv is not None:
            rec["dir"] = dirv
        rec["inv_area"] = self.inv_area
        rec["r"] = r
        rec["emit_time"] = self.emit_time
        return rec

    def export(self) -> np.ndarray:
        raise NotImplementedError("Can not call virtual method to be overridden.")

    def __repr__(self):
        return (f"<{self.type.capitalize()} light source. Intensity: {self.intensity}. "
                f"Area: {1. / self.inv_area:.5f}. Attached = {self.attached}>")
