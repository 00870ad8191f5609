"""Host medium descriptor.  Only ``ior`` reaches the ``pt`` path (reference bxdf/medium.py:24-66;
tracer/path_tracer.py:456 passes world.medium for its ior only) but all XML fields are parsed the
same way so scene files stay interchangeable."""
import xml.etree.ElementTree as xet

import numpy as np

from ..parsers.general_parser import get, rgb_parse
from ..utils.tools import CONSOLE

__all__ = ["Medium_np"]


class Medium_np:
    _type_mapping = {"hg": 0, "multi-hg": 1, "rayleigh": 2, "mie": 3, "transparent": -1}

    @staticmethod
    def is_supported_type(_type: str):
        return Medium_np._type_mapping.get(_type, None)

    def __init__(self, elem: xet.Element, is_world: bool = False):
        self.ior = 1.0
        self.u_a = np.zeros(3, np.float32)
        self.u_s = np.zeros(3, np.float32)
        self.par = np.zeros(3, np.float32)
        self.pdf = np.float32([1.0, 0.0, 0.0])
        self.type_id = -1
        self.type_name = "transparent"
        elem_to_query = {"rgb": rgb_parse, "float": lambda el: get(el, "value")}
        if elem is not None:
            type_name = elem.get("type")
            if type_name in Medium_np._type_mapping:
                self.type_id = Medium_np._type_mapping[type_name]
            else:
                raise NotImplementedError(f"Medium type '{type_name}' is not supported.")
            self.type_name = type_name
            for tag, query_func in elem_to_query.items():
                for tag_elem in elem.findall(tag):
                    name = tag_elem.get("name")
                    if hasattr(self, name):
                        setattr(self, name, query_func(tag_elem))
        elif not is_world:
            CONSOLE.log("[yellow]:warning: Warning: default initialization yields <transparent>, which is a trivial medium.")
        self.u_e = self.u_a + self.u_s

    def __repr__(self):
        return (f"<Medium {self.type_name.capitalize()} with ior {self.ior:.3f}, "
                f"extinction: {self.u_e}, scattering: {self.u_s}>")
