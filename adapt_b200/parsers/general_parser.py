"""Low-level XML element parsers (host side, numpy only).

Mirrors the behaviour of the reference's ``parsers/general_parser.py`` (get :12, parse_str :16,
rgb_parse :28, vec3d_parse :48, transform_parse :56, parse_sphere_element :100) so that the same
XML v1.1 scenes produce the same numbers.  Error behaviour (``ValueError`` on malformed nodes) is kept.
"""
from __future__ import annotations

import xml.etree.ElementTree as xet
from typing import Optional, Tuple

import numpy as np
from scipy.spatial.transform import Rotation as Rot

__all__ = ["get", "parse_str", "rgb_parse", "vec3d_parse", "transform_parse", "parse_sphere_element"]


def get(node: xet.Element, name: str, _type=float):
    """Attribute lookup with the reference's "0" default (general_parser.py:12-14)."""
    return _type(node.get(name, "0"))


def parse_str(val_str: str, no_else_branch: bool = False) -> np.ndarray:
    """"a,b,c" / "a b c" -> float32[3]; a lone scalar is broadcast to 3 channels (general_parser.py:16-26)."""
    for sep in (",", " "):
        if sep in val_str:
            return np.float32([float(p.strip()) for p in val_str.split(sep)])
    if no_else_branch:
        raise ValueError("Value can not be a single digit, should be a vector splitted by ',' or [space]")
    return np.float32([float(val_str.strip())] * 3)


def rgb_parse(elem: Optional[xet.Element]) -> np.ndarray:
    """<rgb value="#RRGGBB" | "a,b,c" | "s"> or <rgb r= g= b=> (missing channels default to 0;
    general_parser.py:28-46)."""
    if elem is None:
        raise ValueError("EmptyElementError: Element <RGB> is None.")
    val_str = elem.get("value")
    if val_str is None:
        if elem.get("r"):
            return np.float32([get(elem, "r"), get(elem, "g"), get(elem, "b")])
        raise ValueError("RGBError: RGB element does not contain valid field.")
    if val_str.startswith("#"):
        rgb = np.zeros(3, dtype=np.float32)
        for i in range(3):
            base = 1 + (i << 1)
            rgb[i] = int(val_str[base:base + 2], 16) / 255.0
        return rgb
    return parse_str(val_str)


def vec3d_parse(elem: xet.Element):
    """<point x= y= z=> (general_parser.py:48-54)."""
    if elem.tag == "point":
        if elem.find("value") is None:
            return np.float32([get(elem, "x"), get(elem, "y"), get(elem, "z")])
        return parse_str(elem.get("value"), no_else_branch=True)
    return None


def transform_parse(transform_elem: xet.Element) -> Tuple[Optional[np.ndarray], Optional[np.ndarray], Optional[np.ndarray]]:
    """translate / rotate (euler zxy degrees | quaternion | angle-axis) / scale / lookat.

    For <lookat> the "rotation" slot carries the unit view direction and the translation slot the
    origin; ``up`` is ignored (general_parser.py:56-98).  The odd angle-axis scaling of the reference
    (axis divided by ``|axis| * angle / 180 * pi``) is reproduced as is.
    """
    trans_r, trans_t, trans_s = None, None, None
    for child in transform_elem:
        tag = child.tag
        if tag == "translate":
            trans_t = np.float32([get(child, "x"), get(child, "y"), get(child, "z")])
        elif tag == "rotate":
            rot_type = child.get("type", "euler")
            if rot_type == "euler":
                angles = (get(child, "r"), get(child, "p"), get(child, "y"))
                trans_r = Rot.from_euler("zxy", angles, degrees=True).as_matrix()
            elif rot_type == "quaternion":
                trans_r = Rot.from_quat([get(child, "x"), get(child, "y"), get(child, "z"), get(child, "w")]).as_matrix()
            elif rot_type == "angle-axis":
                axis = np.float32([get(child, "x"), get(child, "y"), get(child, "z")])
                axis /= np.linalg.norm(axis) * get(child, "angle") / 180.0 * np.pi
                trans_r = Rot.from_rotvec(axis).as_matrix()
            else:
                raise ValueError(f"Unsupported rotation representation '{rot_type}'")
        elif tag == "scale":
            trans_s = np.float32([get(child, "x"), get(child, "y"), get(child, "z")])
        elif tag.lower() == "lookat":
            target_point = parse_str(child.get("target"))
            origin_point = parse_str(child.get("origin"))
            direction = target_point - origin_point
            dir_norm = np.linalg.norm(direction)
            if dir_norm < 1e-5:
                raise ValueError("Normal length too small: Target and origin seems to be the same point")
            trans_r = direction / dir_norm
            trans_t = origin_point
        else:
            raise ValueError(f"Unsupported transformation representation '{tag}'")
    return trans_r, trans_t, trans_s


def parse_sphere_element(elem: xet.Element):
    """<shape type="sphere"> -> ((1,2,3) [center; r,r,r], dummy normal) (general_parser.py:100-105)."""
    sphere_info = np.zeros((1, 2, 3), np.float32)
    sphere_info[0, 0] = vec3d_parse(elem.find("point"))
    radius = get(elem.find("float"), "value")
    sphere_info[0, 1] = np.full((3,), radius)
    return sphere_info, np.float32([[0, 1, 0]])
