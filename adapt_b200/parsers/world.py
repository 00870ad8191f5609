"""World block: skybox / ambient / free-space medium (reference parsers/world.py:21-47).
``pt`` only reads ``world.medium.ior`` (tracer/path_tracer.py:456)."""
import xml.etree.ElementTree as xet

import numpy as np

from ..bxdf.medium import Medium_np
from ..utils.tools import CONSOLE
from .general_parser import rgb_parse

__all__ = ["World_np"]


class World_np:
    def __init__(self, elem: xet.Element):
        self.skybox = np.zeros(3, np.float32)
        self.ambient = np.zeros(3, np.float32)
        if elem is not None:
            for rgb_elem in elem.findall("rgb"):
                name = rgb_elem.get("name")
                if hasattr(self, name):
                    setattr(self, name, rgb_parse(rgb_elem))
        self.medium = Medium_np(elem.find("medium") if elem is not None else None, is_world=True)
        self.C = 1.0
        CONSOLE.log(f":earth_asia: World loading completed: \n {self}")

    def export(self):
        return self

    def __repr__(self):
        is_scattering = np.linalg.norm(self.medium.u_e) > 1e-4
        return (f"<World with free space being [{self.medium.type_name.capitalize()}], "
                f"ior: {self.medium.ior:.3f}, scatter: {is_scattering}>")
