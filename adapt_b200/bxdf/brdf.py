"""Host BRDF descriptor ``BRDF_np`` (reference bxdf/brdf.py:35-140).

``export()`` returns one packed ``BXDF_DTYPE`` record (the C-ABI ``adapt_bxdf`` struct of
include/adapt_b200.h) instead of a Taichi struct; field meaning follows the reference ``BRDF``
dataclass (brdf.py:152-158): _type, is_delta, k_d, k_s, k_g, mean.  The eval/sample/pdf methods of
that dataclass are device code and live in csrc/pt_shade.cuh.

The reference compiles microfacet support out by default (``__ENABLE_MICROFACET__ = False``,
brdf.py:8) and silently downgrades such BRDFs to Lambertian (:60-65).  Same default here; call
``set_enable_microfacet(True)`` (or put ``<boolean name="enable_microfacet" value="true"/>`` in the
sensor block) to turn the GGX path on.
"""
import xml.etree.ElementTree as xet

import numpy as np

from ..parsers.general_parser import rgb_parse
from ..renderer.constants import BRDFTag, DEG2RAD
from ..utils.tools import CONSOLE

__all__ = ["BRDF_np", "BXDF_DTYPE", "set_enable_microfacet", "microfacet_enabled"]

_ENABLE_MICROFACET = False

# C layout of adapt_bxdf (16 x 4 bytes)
BXDF_DTYPE = np.dtype([
    ("kind", np.int32),        # 0 = BRDF (opaque), 1 = BSDF
    ("type", np.int32),
    ("is_delta", np.int32),
    ("k_d", np.float32, 3), ("k_s", np.float32, 3), ("k_g", np.float32, 3), ("mean", np.float32, 3),
    ("ior", np.float32),
], align=False)
assert BXDF_DTYPE.itemsize == 64


def set_enable_microfacet(flag: bool):
    global _ENABLE_MICROFACET
    _ENABLE_MICROFACET = bool(flag)


def microfacet_enabled() -> bool:
    return _ENABLE_MICROFACET


class BRDF_np:
    _all_albedo_name = {"reflectance", "albedo", "k_d"}
    _all_glossiness_name = {"glossiness", "shininess", "roughness", "sigma", "k_g"}
    _all_specular_name = {"specular", "ref_ior", "k_s"}
    _type_mapping = {"phong": 0, "lambertian": 1, "specular": 2, "microfacet": 3,
                     "mod-phong": 4, "fresnel-blend": 5, "oren-nayar": 6, "thin-coat": 7}

    def __init__(self, elem: xet.Element, no_setup: bool = False):
        self.type: str = elem.get("type")
        self.type_id = BRDF_np._type_mapping.get(self.type, -1)
        self.id: str = elem.get("id")
        self.k_d = np.ones(3, np.float32)
        self.k_s = np.zeros(3, np.float32)
        self.k_g = np.ones(3, np.float32)
        self.is_delta = False
        self.kd_default = True
        self.ks_default = True
        self.kg_default = True
        self.uv_coords = None
        if not _ENABLE_MICROFACET and self.type_id == BRDFTag.MICROFACET:
            CONSOLE.log(f"[yellow]Warning: [/yellow]BRDF <{self.id}> is microfacet while microfacet BRDF is not enabled. "
                        "Falling back to Lambertian.")
            self.type = "lambertian"
            self.type_id = BRDFTag.LAMBERTIAN

        texture_nodes = elem.findall("texture")
        if len(texture_nodes) > 1:
            CONSOLE.log(f"[yellow]Warning: [/yellow]Only one texture is supported in a BR(S)DF <{self.id}>.")
        rgb_nodes = elem.findall("rgb")
        if len(rgb_nodes) == 0 and len(texture_nodes) == 0:
            CONSOLE.log(f"[yellow]Warning: [/yellow]BSDF <{self.id}> has no surface color / textures defined.")
        for rgb_node in rgb_nodes:
            name = rgb_node.get("name")
            if name is None:
                raise ValueError(f"RGB node in BR(S)DF <{elem.get('id')}> has empty name.")
            if name in BRDF_np._all_albedo_name:
                self.k_d = rgb_parse(rgb_node)
                self.kd_default = False
            elif name in BRDF_np._all_specular_name:
                self.k_s = rgb_parse(rgb_node)
                self.ks_default = False
            elif name in BRDF_np._all_glossiness_name:
                self.k_g = rgb_parse(rgb_node)
                self.kg_default = False
                if name == "roughness":
                    # roughness -> GGX alpha (brdf.py:97-103)
                    if (self.k_g > 1).any() or (self.k_g < 0).any():
                        CONSOLE.log(f"[yellow]Warning: [/yellow]roughness of <{self.id}> clamped to [0, 1].")
                        self.k_g = self.k_g.clip(0, 1)
                    self.k_g = BRDF_np.roughness_to_alpha(self.k_g)
                elif name == "sigma":
                    # sigma (degrees) -> Oren-Nayar A, B; k_g[2] = coating IOR >= 1 (brdf.py:104-110)
                    sigma = self.k_g[0] * DEG2RAD
                    sigma2 = sigma * sigma
                    self.k_g[0] = 1 - (sigma2 / (2 * (sigma2 + 0.33)))
                    self.k_g[1] = 0.45 * sigma2 / (sigma2 + 0.09)
                    self.k_g[2] = max(1.0, self.k_g[2])
        if not no_setup:
            self.setup()

    @staticmethod
    def roughness_to_alpha(roughness: np.ndarray) -> np.ndarray:
        """pbrt-v3 TrowbridgeReitzDistribution::RoughnessToAlpha polynomial (brdf.py:115-120)."""
        x = np.log(np.maximum(roughness, 1e-3))
        return 1.62142 + 0.819955 * x + 0.1734 * x * x + 0.0171201 * (x ** 3) + 0.000640711 * (x ** 4)

    def setup(self):
        if self.type not in BRDF_np._type_mapping:
            raise NotImplementedError(f"Unknown BRDF type: {self.type}")
        if self.type_id == BRDFTag.SPECULAR:
            self.is_delta = True
        elif self.type_id == BRDFTag.FRESNEL_BLEND:
            # Ashikhmin-Shirley normalisation sqrt((nu+1)(nv+1)) / 8pi kept in k_g[2] (brdf.py:127-128)
            self.k_g[2] = np.sqrt((self.k_g[0] + 1) * (self.k_g[1] + 1)) / (8.0 * np.pi)

    def export(self) -> np.ndarray:
        if self.type_id == -1:
            raise ValueError("It seems that this BRDF is not properly initialized with type_id = -1")
        rec = np.zeros((), dtype=BXDF_DTYPE)
        rec["kind"] = 0
        rec["type"] = self.type_id
        rec["is_delta"] = int(self.is_delta)
        rec["k_d"] = self.k_d
        rec["k_s"] = self.k_s
        rec["k_g"] = self.k_g
        rec["mean"] = np.float32([self.k_d.mean(), self.k_s.mean(), self.k_g.mean()])
        rec["ior"] = 1.0
        return rec

    def __repr__(self) -> str:
        return (f"<{self.type.capitalize()} BRDF, default:"
                f"[{int(self.kd_default), int(self.ks_default), int(self.kg_default)}]>")
