"""Volumetric path tracer behind the reference's renderer-class boundary (`rdr_mapping["vpt"]`, render.py:33; renderer/vpt.py:29-262).

Same constructor and driver surface as `Renderer`; the device side runs k_logic_vpt / k_trace_vpt (homogeneous media: the world's
free-space medium and the media attached to BSDF objects, HG / multi-lobe HG / Rayleigh phase functions, null surfaces, track_ray
transmittance).  Grid volumes (`<volume>`, bxdf/volume.py) are refused at scene-pack time.
"""
from typing import List

from adapt_b200.renderer.vanilla_renderer import Renderer


class VolumeRenderer(Renderer):
    def __init__(self, emitters: List, array_info: dict, objects: List, prop: dict, **kwargs):
        kwargs["integrator"] = "vpt"
        super().__init__(emitters, array_info, objects, prop, **kwargs)
