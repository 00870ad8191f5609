"""Enumerant tags and math constants shared by host descriptors (reference renderer/constants.py:8-53)."""
import math

TRANSPORT_UNI = -1
TRANSPORT_RAD = 0
TRANSPORT_IMP = 1

INV_PI = 1.0 / math.pi
INV_2PI = INV_PI * 0.5
PI2 = 2.0 * math.pi
DEG2RAD = math.pi / 180.0
RAD2DEG = 180.0 * INV_PI


class BRDFTag:
    BLING_PHONG = 0
    LAMBERTIAN = 1
    SPECULAR = 2
    MICROFACET = 3
    MOD_PHONG = 4
    FRESNEL_BLEND = 5
    OREN_NAYAR = 6
    THIN_COAT = 7


# emitter type tags (emitters/abtract_source.py:29-33)
POINT_SOURCE = 0
AREA_SOURCE = 1
SPOT_SOURCE = 2
COLLIMATED_SOURCE = 4
