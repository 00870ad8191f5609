"""Host BSDF descriptor ``BSDF_np`` (reference bxdf/bsdf.py:29-58): det-refraction (0),
lambertian transmission (1), null (-1), each with an attached medium of which ``pt`` reads the ior."""
import xml.etree.ElementTree as xet

import numpy as np

from .brdf import BRDF_np, BXDF_DTYPE
from .medium import Medium_np

__all__ = ["BSDF_np"]


class BSDF_np(BRDF_np):
    _bsdf_type_mapping = {"det-refraction": 0, "null": -1, "lambertian": 1}

    def __init__(self, elem: xet.Element):
        super().__init__(elem, True)
        self.medium = Medium_np(elem.find("medium"))
        self.is_delta = False
        self.setup()
        if self.type_id == 0:
            self.is_delta = True

    def setup(self):
        if self.type not in BSDF_np._bsdf_type_mapping:
            raise NotImplementedError(f"Unknown BSDF type: {self.type}")
        self.type_id = BSDF_np._bsdf_type_mapping[self.type]

    def export(self) -> np.ndarray:
        rec = np.zeros((), dtype=BXDF_DTYPE)
        rec["kind"] = 1
        rec["type"] = self.type_id
        rec["is_delta"] = int(self.is_delta)
        rec["k_d"] = self.k_d
        rec["k_s"] = self.k_s
        rec["k_g"] = self.k_g
        rec["mean"] = np.float32([self.k_d.mean(), self.k_s.mean(), self.k_g.mean()])
        rec["ior"] = self.medium.ior
        return rec

    def __repr__(self) -> str:
        return f"<{self.type.capitalize()} BSDF with {self.medium.__repr__()} >"
