"""`taichi.math` subset used by the reference hot path (see taichi/__init__.py in this directory)."""
import math as _pm

import numpy as np

import taichi as _ti
from taichi import Matrix, Vector, _raw, _un, _VecType, _MatType

pi = _pm.pi
e = _pm.e
inf = float("inf")
nan = float("nan")

vec2 = _VecType(2)
vec3 = _VecType(3)
vec4 = _VecType(4)
ivec2 = _VecType(2, int)
ivec3 = _VecType(3, int)
mat2 = _MatType(2, 2)
mat3 = _MatType(3, 3)
mat4 = _MatType(4, 4)

sqrt, sin, cos, tan, asin, acos, exp, log, floor, ceil, pow, atan2 = (
    _ti.sqrt, _ti.sin, _ti.cos, _ti.tan, _ti.asin, _ti.acos, _ti.exp, _ti.log, _ti.floor, _ti.ceil, _ti.pow, _ti.atan2)
max, min = _ti.max, _ti.min      # noqa: A001
atan = _un(np.arctan)
log2 = _un(np.log2)
exp2 = _un(np.exp2)


def dot(a, b):
    # a.x*b.x + a.y*b.y + a.z*b.z in f32
    p = (_raw(a) * _raw(b)).astype(np.float32)
    acc = p[0]
    for k in range(1, len(p)):
        acc = np.float32(acc + p[k])
    return acc


def cross(a, b):
    a, b = _raw(a), _raw(b)
    return Vector._wrap(np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], np.float32))


def length(a):
    return a.norm()


def normalize(a):
    return a.normalized()


def mix(x, y, a):
    # taichi.math.mix: x * (1 - a) + y * a
    one = np.float32(1.0)
    if isinstance(x, Vector) or isinstance(y, Vector) or isinstance(a, Vector):
        rx, ry, ra = _raw(x), _raw(y), _raw(a)
        return Vector._wrap(np.asarray(rx * (one - ra) + ry * ra, np.float32))
    return np.float32(np.float32(x) * (one - np.float32(a)) + np.float32(y) * np.float32(a))


def clamp(x, lo, hi):
    return _ti.min(_ti.max(x, lo), hi)


def sign(x):
    if isinstance(x, Vector):
        return Vector._wrap(np.sign(x.a))
    return np.float32(np.sign(np.float32(x)))


def isnan(x):
    if isinstance(x, Vector):
        return Vector._wrap(np.isnan(x.a))
    return bool(np.isnan(x))


def isinf(x):
    if isinstance(x, Vector):
        return Vector._wrap(np.isinf(x.a))
    return bool(np.isinf(x))


def fract(x):
    return x - _ti.floor(x)


def radians(x):
    return np.float32(x) * np.float32(pi / 180.0)


def degrees(x):
    return np.float32(x) * np.float32(180.0 / pi)


def step(edge, x):
    return np.float32(0.0) if x < edge else np.float32(1.0)


def smoothstep(e0, e1, x):
    t = clamp((x - e0) / (e1 - e0), 0.0, 1.0)
    return t * t * (3.0 - 2.0 * t)


def reflect(i, n):
    return i - 2.0 * dot(n, i) * n


def eye(n):
    return Matrix(np.eye(n, dtype=np.float32))
