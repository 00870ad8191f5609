"""Minimal pure-Python stand-in for the `taichi` package: JUST enough of its API to execute the reference's
own `@ti.func` / `@ti.kernel` bodies (renderer/vanilla_renderer.py, tracer/*.py, bxdf/*.py, emitters/*.py,
sampler/*.py, la/*.py under /root/reference) as ordinary Python, one pixel-sample at a time, in float32.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_reference_golden.py, in the development container,
to produce golden vectors FROM THE REFERENCE'S SOURCE (Taichi 1.6.0 itself is not installable offline).
Nothing in the product, the -m gpu tests, smoke() or bench.py imports it.

Semantics reproduced (Taichi 1.6 language reference, as used by the path):
  * scalars are f32 / i32: every float that enters a vector or leaves `ti.random`, a field or a maths
    function is numpy.float32; Python float literals are "weak" (NEP 50) so f32 arithmetic stays f32;
  * vectors / matrices are value types: every operation returns a new object, field and struct reads
    return copies, so `a = b` followed by `a.fill(0)` cannot alias storage other than `b` itself;
  * `@ti.func` arguments are passed BY VALUE (structs are copied) unless annotated `ti.template()`,
    which passes by reference -- the reference relies on this (path_tracer.py:449-453 mutates `it`);
  * `%` on ints is floor-mod and `//` floor division (same as Python);
  * `ti.random(float)` = (u32 >> 8) / 2^24, `ti.random(int)` = the u32 reinterpreted as i32, both drawn from a
    pluggable generator (`set_rng`) so the golden script can key it by (pixel, sample) exactly like the oracle.
Not reproduced: LLVM fast-math contraction / reassociation (rounding-level differences only).
"""
from __future__ import annotations

import copy as _copy
import functools
import inspect
import math as _pm

import numpy as np

np.seterr(all="ignore")

f32 = np.float32
i32 = np.int32
u32 = np.uint32
f64 = np.float64
float32 = np.float32
int32 = np.int32
float64 = np.float64
float = float          # noqa: A001  (ti.float)
int = int              # noqa: A001
cpu = "cpu"
gpu = "gpu"
cuda = "cuda"
vulkan = "vulkan"
metal = "metal"


def init(*a, **k):
    return None


def _f(x):
    return np.float32(x)


# ------------------------------------------------------------------------------------------------ RNG
class _DefaultRng:
    def __init__(self):
        self.g = np.random.default_rng(0)

    def next_u32(self):
        return builtins_int(self.g.integers(0, 1 << 32))


import builtins as _b  # noqa: E402

builtins_int = _b.int
builtins_float = _b.float
_rng = _DefaultRng()


def set_rng(r):
    """r must provide next_u32() -> Python int in [0, 2^32)."""
    global _rng
    _rng = r


def random(dtype=builtins_float):
    u = _rng.next_u32()
    if dtype in (builtins_int, i32):
        return u - (1 << 32) if u >= (1 << 31) else u
    return np.float32(u >> 8) * np.float32(1.0 / 16777216.0)


# ------------------------------------------------------------------------------------------------ vectors / matrices
def _raw(x):
    if isinstance(x, (Vector, Matrix)):
        return x.a
    return x


class Vector:
    __slots__ = ("a",)
    __array_priority__ = 1000
    __array_ufunc__ = None           # numpy scalars defer to our reflected operators

    def __init__(self, *args, dt=None):
        if len(args) == 1:
            v = args[0]
            if isinstance(v, Vector):
                arr = v.a.copy()
            elif isinstance(v, (list, tuple)):
                arr = np.array([_raw(e) for e in v])
            else:
                arr = np.array(v)
        else:
            arr = np.array([_raw(e) for e in args])
        if arr.dtype.kind == "f" or dt in (builtins_float, f32):
            arr = arr.astype(np.float32)
        elif arr.dtype.kind in "iu":
            arr = arr.astype(np.int32)
        elif arr.dtype.kind == "b":
            pass
        self.a = arr

    # ---- construction helpers
    @staticmethod
    def _wrap(arr):
        v = Vector.__new__(Vector)
        if arr.dtype == np.float64:
            arr = arr.astype(np.float32)
        elif arr.dtype == np.int64:
            arr = arr.astype(np.int32)
        v.a = arr
        return v

    @staticmethod
    def field(n, dtype, shape=None):
        return Field((n,), dtype, shape)

    @staticmethod
    def zero(dt, n):
        return Vector._wrap(np.zeros(n, np.float32 if dt in (builtins_float, f32) else np.int32))

    # ---- value protocol
    def __len__(self):
        return len(self.a)

    def __iter__(self):
        return iter(self.a)          # numpy scalars (float32 / int32)

    def __getitem__(self, i):
        return self.a[i]

    def __setitem__(self, i, v):
        self.a[i] = v

    def __repr__(self):
        return f"Vector({self.a.tolist()})"

    def __copy__(self):
        return Vector._wrap(self.a.copy())

    def __deepcopy__(self, memo):
        return Vector._wrap(self.a.copy())

    x = property(lambda s: s.a[0])
    y = property(lambda s: s.a[1])
    z = property(lambda s: s.a[2])
    w = property(lambda s: s.a[3])
    n = property(lambda s: len(s.a))

    def _coerce(self, o):
        o = _raw(o)
        if isinstance(o, builtins_float) and self.a.dtype.kind in "iu":
            return np.float32(o)
        return o

    def _bin(self, o, op):
        return Vector._wrap(np.asarray(op(self.a, self._coerce(o))))

    def _rbin(self, o, op):
        return Vector._wrap(np.asarray(op(self._coerce(o), self.a)))

    __add__ = lambda s, o: s._bin(o, np.add)
    __radd__ = lambda s, o: s._rbin(o, np.add)
    __sub__ = lambda s, o: s._bin(o, np.subtract)
    __rsub__ = lambda s, o: s._rbin(o, np.subtract)
    __mul__ = lambda s, o: s._bin(o, np.multiply)
    __rmul__ = lambda s, o: s._rbin(o, np.multiply)
    __truediv__ = lambda s, o: s._bin(o, np.true_divide)
    __rtruediv__ = lambda s, o: s._rbin(o, np.true_divide)
    __pow__ = lambda s, o: s._bin(o, np.power)
    __rpow__ = lambda s, o: s._rbin(o, np.power)
    __neg__ = lambda s: Vector._wrap(-s.a)
    __lt__ = lambda s, o: s._bin(o, np.less)
    __le__ = lambda s, o: s._bin(o, np.less_equal)
    __gt__ = lambda s, o: s._bin(o, np.greater)
    __ge__ = lambda s, o: s._bin(o, np.greater_equal)

    def __eq__(self, o):
        return self._bin(o, np.equal)

    def __ne__(self, o):
        return self._bin(o, np.not_equal)

    __hash__ = None

    def __bool__(self):
        raise TypeError("truth value of a taichi vector is ambiguous (use .any()/.all())")

    # ---- methods used by the reference
    def norm(self, eps=0.0):
        return np.sqrt(np.dot(self.a, self.a) + np.float32(eps)).astype(np.float32)

    def norm_sqr(self):
        return np.float32(np.dot(self.a, self.a))

    def norm_inv(self, eps=0.0):
        return np.float32(1.0) / self.norm(eps)

    def normalized(self, eps=0.0):
        # taichi: v / (v.norm() + eps) -- one reciprocal, three multiplies under fast-math
        inv = np.float32(1.0) / (self.norm() + np.float32(eps))
        return Vector._wrap(self.a * inv)

    def dot(self, o):
        return np.float32(np.dot(self.a, _raw(o)))

    def cross(self, o):
        return Vector._wrap(np.cross(self.a, _raw(o)).astype(np.float32))

    def outer_product(self, o):
        return Matrix(np.outer(self.a, _raw(o)))

    def max(self):
        return self.a.max()

    def min(self):
        return self.a.min()

    def sum(self):
        return self.a.sum(dtype=self.a.dtype)

    def any(self):
        return bool(self.a.any())

    def all(self):
        return bool(self.a.all())

    def fill(self, v):
        self.a[...] = v

    def to_numpy(self):
        return self.a.copy()

    def to_list(self):
        return self.a.tolist()

    def cast(self, dt):
        return Vector._wrap(self.a.astype(np.float32 if dt in (builtins_float, f32) else np.int32))

    def __matmul__(self, o):
        return np.float32(np.dot(self.a, _raw(o)))


class Matrix:
    __slots__ = ("a",)
    __array_priority__ = 1000
    __array_ufunc__ = None           # numpy scalars defer to our reflected operators

    def __init__(self, arr):
        if isinstance(arr, Matrix):
            arr = arr.a
        arr = np.array([[_raw(e) for e in row] for row in arr]) if isinstance(arr, (list, tuple)) else np.array(arr)
        self.a = arr.astype(np.float32) if arr.dtype.kind == "f" else arr.astype(np.int32)

    @staticmethod
    def cols(vs):
        return Matrix(np.stack([_raw(v) for v in vs], axis=1))

    @staticmethod
    def rows(vs):
        return Matrix(np.stack([_raw(v) for v in vs], axis=0))

    @staticmethod
    def diag(dim, val):
        return Matrix(np.eye(dim, dtype=np.float32) * np.float32(val))

    @staticmethod
    def zero(dt, n, m=None):
        if m is None:
            return Vector.zero(dt, n)
        return Matrix(np.zeros((n, m), np.float32))

    @staticmethod
    def identity(dt, n):
        return Matrix(np.eye(n, dtype=np.float32))

    @staticmethod
    def field(n, m, dtype, shape=None):
        return Field((n, m), dtype, shape)

    def __copy__(self):
        return Matrix(self.a.copy())

    def __deepcopy__(self, memo):
        return Matrix(self.a.copy())

    def __repr__(self):
        return f"Matrix({self.a.tolist()})"

    def __getitem__(self, ij):
        return self.a[ij]

    def __setitem__(self, ij, v):
        self.a[ij] = v

    def __matmul__(self, o):
        if isinstance(o, Vector):
            # row-by-row dot products in f32 (what the generated code does)
            return Vector._wrap((self.a @ o.a).astype(np.float32))
        return Matrix((self.a @ _raw(o)).astype(np.float32))

    def _bin(self, o, op):
        return Matrix(op(self.a, _raw(o)))

    __add__ = lambda s, o: s._bin(o, np.add)
    __radd__ = lambda s, o: s._bin(o, np.add)
    __sub__ = lambda s, o: s._bin(o, np.subtract)
    __rsub__ = lambda s, o: Matrix(np.subtract(_raw(o), s.a))
    __mul__ = lambda s, o: s._bin(o, np.multiply)
    __rmul__ = lambda s, o: s._bin(o, np.multiply)
    __truediv__ = lambda s, o: s._bin(o, np.true_divide)
    __neg__ = lambda s: Matrix(-s.a)

    def transpose(self):
        return Matrix(self.a.T.copy())

    def determinant(self):
        m = self.a
        if m.shape == (3, 3):
            return np.float32(m[0, 0] * (m[1, 1] * m[2, 2] - m[2, 1] * m[1, 2]) - m[1, 0] * (m[0, 1] * m[2, 2] - m[2, 1] * m[0, 2])
                              + m[2, 0] * (m[0, 1] * m[1, 2] - m[1, 1] * m[0, 2]))
        return np.float32(np.linalg.det(m.astype(np.float64)))

    def inverse(self):
        # Taichi's closed-form 3x3 inverse: adjugate / determinant, all in f32
        m = self.a
        if m.shape != (3, 3):
            return Matrix(np.linalg.inv(m.astype(np.float64)).astype(np.float32))
        inv_det = np.float32(1.0) / self.determinant()
        E = lambda x, y: m[x % 3, y % 3]       # noqa: E731
        out = np.empty((3, 3), np.float32)
        for i in range(3):
            for j in range(3):
                out[i, j] = inv_det * (E(j + 1, i + 1) * E(j + 2, i + 2) - E(j + 2, i + 1) * E(j + 1, i + 2))
        return Matrix(out)

    def fill(self, v):
        self.a[...] = v

    def to_numpy(self):
        return self.a.copy()

    def trace(self):
        return np.float32(np.trace(self.a))


# ------------------------------------------------------------------------------------------------ fields / SNodes
class _Axis:
    def __init__(self, k):
        self.k = k


i, j, k, l = _Axis(0), _Axis(1), _Axis(2), _Axis(3)      # noqa: E741
ij = (i, j)
ijk = (i, j, k)

_struct_for_hook = None


def set_struct_for_hook(fn):
    """fn(field) -> iterable of index tuples, or None for the default (all indices). Lets the golden script
    restrict `for i, j in self.pixels` to a pixel subset and re-key the RNG per pixel."""
    global _struct_for_hook
    _struct_for_hook = fn


class Field:
    """Dense scalar / vector / matrix / struct field with lazy allocation (placed by an SNode or shaped at creation)."""

    def __init__(self, elem_shape, dtype, shape=None, struct_cls=None):
        self.elem_shape = tuple(elem_shape)
        self.struct_cls = struct_cls
        self.np_dtype = np.float32 if dtype in (builtins_float, f32, f64) else np.int32
        self.data = None
        self.shape = None
        self.written = set()
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        if isinstance(shape, builtins_int):
            shape = (shape,)
        self.shape = tuple(shape)
        if self.struct_cls is not None:
            self.data = np.empty(self.shape, dtype=object)
            for idx in np.ndindex(*self.shape):
                self.data[idx] = self.struct_cls()
        else:
            self.data = np.zeros(self.shape + self.elem_shape, self.np_dtype)

    @staticmethod
    def _idx(key):
        if key is None:
            return ()
        if isinstance(key, tuple):
            return tuple(builtins_int(x) for x in key)
        if isinstance(key, Vector):
            return tuple(builtins_int(x) for x in key.a)
        return (builtins_int(key),)

    def __getitem__(self, key):
        idx = self._idx(key)
        if self.struct_cls is not None:
            return self.data[idx]                 # structs in fields are references (`self.src_field[i].obj_ref_id = -1` works)
        v = self.data[idx]
        if self.elem_shape == ():
            return v.item() if self.np_dtype == np.int32 else np.float32(v)
        if len(self.elem_shape) == 1:
            return Vector._wrap(v.copy())
        return Matrix(v.copy())

    def __setitem__(self, key, val):
        idx = self._idx(key)
        self.written.add(idx)
        if self.struct_cls is not None:
            self.data[idx] = _copy.deepcopy(val)
        else:
            self.data[idx] = _raw(val)

    def __iter__(self):
        if _struct_for_hook is not None:
            it = _struct_for_hook(self)
            if it is not None:
                return iter(it)
        idxs = np.ndindex(*self.shape)
        if len(self.shape) == 1:
            return iter(range(self.shape[0]))
        return iter(idxs)

    def from_numpy(self, arr):
        arr = np.asarray(arr)
        if self.data is None:
            self._alloc(arr.shape[: arr.ndim - len(self.elem_shape)])
        self.data[...] = arr.astype(self.np_dtype).reshape(self.data.shape)

    def to_numpy(self):
        return self.data.copy()

    def fill(self, v):
        self.data[...] = v


class _SNode:
    def __init__(self, shape=(), kind="dense"):
        self.shape = tuple(shape)
        self.kind = kind
        self.fields = []

    def _child(self, axes, dims, kind):
        if isinstance(axes, _Axis):
            axes = (axes,)
        if isinstance(dims, builtins_int):
            dims = (dims,) * len(axes)
        shape = list(self.shape)
        for ax, d in zip(axes, dims):
            while len(shape) <= ax.k:
                shape.append(1)
            shape[ax.k] = shape[ax.k] * builtins_int(d) if ax.k < len(self.shape) else builtins_int(d)
        return _SNode(shape, kind)

    def dense(self, axes, dims):
        return self._child(axes, dims, "dense")

    def bitmasked(self, axes, dims):
        return self._child(axes, dims, "bitmasked")

    def pointer(self, axes, dims):
        return self._child(axes, dims, "pointer")

    def place(self, *fields):
        for f in fields:
            f._alloc(self.shape)
            self.fields.append(f)
        return self


class _Root(_SNode):
    pass


root = _Root()


def is_active(snode, idx):
    key = Field._idx(idx)
    return any(key in f.written for f in snode.fields)


def field(dtype, shape=None):
    return Field((), dtype, shape)


# ------------------------------------------------------------------------------------------------ structs
class _StructMeta(type):
    pass


def _default_for(anno):
    if anno in (builtins_int, i32):
        return 0
    if anno in (builtins_float, f32):
        return np.float32(0.0)
    if isinstance(anno, _VecType):
        return anno.zero()
    if isinstance(anno, _MatType):
        return Matrix(np.zeros((anno.n, anno.m), np.float32))
    if isinstance(anno, type) and issubclass(anno, _StructBase):
        return anno()
    raise TypeError(f"unsupported struct member type {anno!r}")


def _convert_for(anno, v):
    if anno in (builtins_int, i32):
        return builtins_int(v)
    if anno in (builtins_float, f32):
        return np.float32(v)
    if isinstance(anno, _VecType):
        return anno(v)
    if isinstance(anno, _MatType):
        return Matrix(v)
    if isinstance(anno, type) and issubclass(anno, _StructBase):
        return _copy.deepcopy(v)
    return v


class _StructBase:
    _members: dict = {}

    def __init__(self, **kw):
        for name, anno in self._members.items():
            object.__setattr__(self, name, _convert_for(anno, kw[name]) if name in kw else _default_for(anno))
        extra = set(kw) - set(self._members)
        if extra:
            raise TypeError(f"unknown struct members {extra}")

    def __getattribute__(self, name):
        v = object.__getattribute__(self, name)
        if isinstance(v, (Vector, Matrix)):
            return _copy.copy(v)              # value semantics: reads copy
        return v

    def __setattr__(self, name, v):
        anno = self._members.get(name)
        object.__setattr__(self, name, _convert_for(anno, v) if anno is not None else v)

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        for name in self._members:
            object.__setattr__(new, name, _copy.deepcopy(object.__getattribute__(self, name)))
        return new

    def __repr__(self):
        return f"{self.__class__.__name__}(" + ", ".join(f"{n}={object.__getattribute__(self, n)!r}" for n in self._members) + ")"

    @classmethod
    def field(cls, shape=None):
        return Field((), None, shape, struct_cls=cls)


def dataclass(cls):
    members = dict(getattr(cls, "__annotations__", {}))
    ns = {k: v for k, v in cls.__dict__.items() if k not in ("__dict__", "__weakref__", "__annotations__")}
    ns["_members"] = members
    new = type(cls.__name__, (_StructBase,), ns)
    new.__module__ = cls.__module__
    return new


class _Struct:
    @staticmethod
    def field(members, shape=None):
        cls = type("AnonStruct", (_StructBase,), {"_members": dict(members)})
        return Field((), None, shape, struct_cls=cls)


Struct = _Struct


# ------------------------------------------------------------------------------------------------ types
class _Template:
    def __call__(self):
        return self


class _TemplateMarker:
    pass


def template():
    return _TemplateMarker()


class _VecType:
    def __init__(self, n, dt=builtins_float):
        self.n, self.dt = n, dt

    def __call__(self, *args):
        if len(args) == 1 and not isinstance(args[0], (list, tuple, Vector, np.ndarray)):
            return Vector([args[0]] * self.n, dt=self.dt)           # broadcast a scalar
        flat = []
        if len(args) > 1:
            for a in args:
                if isinstance(a, Vector):
                    flat.extend(a.a.tolist() if False else list(a.a))
                else:
                    flat.append(a)
            return Vector(flat, dt=self.dt)
        return Vector(args[0], dt=self.dt)

    def zero(self):
        return Vector(np.zeros(self.n, np.float32 if self.dt in (builtins_float, f32) else np.int32))

    def field(self, shape=None):
        return Field((self.n,), self.dt, shape)


class _MatType:
    def __init__(self, n, m, dt=builtins_float):
        self.n, self.m = n, m

    def __call__(self, *args):
        if len(args) == 1:
            return Matrix(args[0])
        return Matrix(np.array([_raw(a) for a in args], np.float32).reshape(self.n, self.m))


class _NdArrayType:
    pass


class types:                                                       # noqa: N801
    @staticmethod
    def vector(n, dtype=builtins_float):
        return _VecType(n, dtype)

    @staticmethod
    def matrix(n, m, dtype=builtins_float):
        return _MatType(n, m, dtype)

    @staticmethod
    def ndarray(*a, **k):
        return _NdArrayType()

    @staticmethod
    def struct(**members):
        return type("AnonStruct", (_StructBase,), {"_members": dict(members)})


# ------------------------------------------------------------------------------------------------ decorators
def _is_struct(x):
    return isinstance(x, _StructBase)


def func(fn):
    """Arguments by value (struct copies) unless annotated ti.template()."""
    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        return fn
    params = list(sig.parameters.values())
    by_ref = [isinstance(p.annotation, _TemplateMarker) or p.name == "self" for p in params]
    names = [p.name for p in params]

    def by_value(a):
        if _is_struct(a) or isinstance(a, (Vector, Matrix)):
            return _copy.deepcopy(a)
        if type(a) is builtins_float:          # every real scalar inside a Taichi function is f32
            return np.float32(a)
        return a

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        args = list(args)
        for n, a in enumerate(args):
            if n < len(by_ref) and not by_ref[n]:
                args[n] = by_value(a)
        for key, a in list(kwargs.items()):
            if key in names and not by_ref[names.index(key)]:
                kwargs[key] = by_value(a)
        return fn(*args, **kwargs)

    return wrapper


def kernel(fn):
    return fn


def pyfunc(fn):
    return fn


def data_oriented(cls):
    return cls


class experimental:                                                # noqa: N801
    real_func = staticmethod(func)


real_func = func


def static(x, *rest):
    return x if not rest else (x,) + rest


def loop_config(**k):
    return None


def ndrange(*dims):
    import itertools
    rs = [range(*d) if isinstance(d, tuple) else range(builtins_int(d)) for d in dims]
    return itertools.product(*rs)


def grouped(x):
    return iter(x)


def cast(x, dt):
    if isinstance(x, Vector):
        return x.cast(dt)
    return builtins_int(x) if dt in (builtins_int, i32) else np.float32(x)


def sync():
    return None


class profiler:                                                    # noqa: N801
    @staticmethod
    def print_kernel_profiler_info(*a, **k):
        return None

    @staticmethod
    def clear_kernel_profiler_info(*a, **k):
        return None


class tools:                                                       # noqa: N801
    @staticmethod
    def imwrite(*a, **k):
        raise NotImplementedError


# ------------------------------------------------------------------------------------------------ maths
def _un(npf):
    def f(x):
        if isinstance(x, Vector):
            return Vector._wrap(npf(x.a).astype(x.a.dtype if x.a.dtype.kind == "f" else np.float32))
        return np.float32(npf(np.float32(x)))
    return f


sqrt = _un(np.sqrt)
sin = _un(np.sin)
cos = _un(np.cos)
tan = _un(np.tan)
asin = _un(np.arcsin)
acos = _un(np.arccos)
exp = _un(np.exp)
log = _un(np.log)
tanh = _un(np.tanh)
rsqrt = _un(lambda x: np.float32(1.0) / np.sqrt(x))


def floor(x, dtype=None):
    if isinstance(x, Vector):
        return Vector._wrap(np.floor(x.a))
    r = np.floor(np.float32(x))
    return builtins_int(r) if dtype in (builtins_int, i32) else np.float32(r)


def ceil(x, dtype=None):
    if isinstance(x, Vector):
        return Vector._wrap(np.ceil(x.a))
    r = np.ceil(np.float32(x))
    return builtins_int(r) if dtype in (builtins_int, i32) else np.float32(r)


def abs(x):                                                         # noqa: A001
    if isinstance(x, Vector):
        return Vector._wrap(np.abs(x.a))
    if isinstance(x, builtins_int):
        return _b.abs(x)
    return np.float32(np.abs(np.float32(x)))


def _is_int(x):
    return isinstance(x, (builtins_int, np.integer)) and not isinstance(x, (bool, np.bool_))


def _minmax(npf, pyf):
    def f(*xs):
        if len(xs) == 1:
            return xs[0]
        acc = xs[0]
        for o in xs[1:]:
            if isinstance(acc, Vector) or isinstance(o, Vector):
                acc = Vector._wrap(np.asarray(npf(_raw(acc), _raw(o))))
            elif _is_int(acc) and _is_int(o):
                acc = pyf(builtins_int(acc), builtins_int(o))
            else:
                acc = np.float32(npf(np.float32(acc), np.float32(o)))      # fminf / fmaxf: a NaN operand is dropped
        return acc
    return f


max = _minmax(np.fmax, _b.max)                                     # noqa: A001
min = _minmax(np.fmin, _b.min)                                     # noqa: A001


def pow(a, b):                                                     # noqa: A001
    if isinstance(a, Vector) or isinstance(b, Vector):
        ra, rb = _raw(a), _raw(b)
        if not isinstance(ra, np.ndarray):
            ra = np.float32(ra)
        if not isinstance(rb, np.ndarray):
            rb = np.float32(rb)
        return Vector._wrap(np.power(ra, rb).astype(np.float32))
    if _is_int(a) and _is_int(b):
        return builtins_int(a) ** builtins_int(b)
    return np.float32(np.power(np.float32(a), np.float32(b)))


def atan2(y, x):
    return np.float32(np.arctan2(np.float32(y), np.float32(x)))


def select(c, a, b):
    if isinstance(c, Vector) or isinstance(a, Vector) or isinstance(b, Vector):
        ra, rb = _raw(a), _raw(b)
        if isinstance(ra, builtins_float):
            ra = np.float32(ra)
        if isinstance(rb, builtins_float):
            rb = np.float32(rb)
        return Vector._wrap(np.where(_raw(c), ra, rb))
    r = a if c else b
    # both branches share one type in Taichi: a float branch makes the result f32
    if isinstance(a, (builtins_float, np.floating)) or isinstance(b, (builtins_float, np.floating)):
        if not isinstance(r, (bool, np.bool_)):
            return np.float32(r)
    return r


from . import math  # noqa: E402,F401
