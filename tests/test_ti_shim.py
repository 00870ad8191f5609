"""Pins the Taichi stand-in (tests/golden/ti_shim, TEST INFRASTRUCTURE: the package the golden generator executes the reference's
unmodified modules on) to the language semantics the reference's path relies on.  Every case names the Taichi 1.6 rule it checks and the
place in the reference that depends on it; the expected values are closed-form or computed independently in float64 -- nothing here
comes from the stand-in itself.  The one rule the reference documents with a script of its own is by-value / by-reference argument
passing (assets/ti_tests/ref_test.py)."""
import os
import sys

import numpy as np
import pytest

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ti_shim")


@pytest.fixture(scope="module")
def ti():
    sys.path.insert(0, SHIM)
    try:
        import taichi
        yield taichi
    finally:
        sys.path.remove(SHIM)
        for name in [m for m in sys.modules if m == "taichi" or m.startswith("taichi.")]:
            del sys.modules[name]


def test_integer_modulo_and_division_follow_python(ti):
    """`%` is floor-mod, `//` floor division (language reference, "arithmetic operators"): sample_light's `ti.random(int) % src_num` must be
    non-negative for negative draws (tracer/path_tracer.py:544), the stratum index `cnt % 16` too (tracer/tracer_base.py:150)."""
    for a, n in [(-7, 3), (-1, 4), (-2147483648, 5), (7, 3), (0, 9)]:
        assert a % n == a - n * int(np.floor(a / n)) and 0 <= a % n < n
        assert a // n == int(np.floor(a / n))


def test_random_draws_are_f32_in_unit_interval_and_full_range_i32(ti):
    """ti.random(float) is an f32 in [0, 1) with 24 random bits; ti.random(int) a full-range i32 (both signs occur)."""
    class Rng:
        def __init__(self, words): self.w = list(words)
        def next_u32(self): return self.w.pop(0)
    ti.set_rng(Rng([0x00000000, 0xffffffff, 0x80000000, 0xffffffff, 0x7fffffff, 0x80000001]))
    assert [ti.random(float) for _ in range(3)] == [np.float32(0.0), np.float32((2 ** 24 - 1) / 2 ** 24), np.float32(0.5)]
    assert [int(ti.random(int)) for _ in range(3)] == [-1, 2 ** 31 - 1, -(2 ** 31) + 1]
    assert isinstance(ti.random(float) if False else np.float32(0), np.float32)


def test_scalars_and_vectors_stay_float32(ti):
    """default_fp = f32 (render.py:69): products and sums round to f32 after every operation, Python literals do not widen them."""
    v = ti.Vector([0.1, 0.2, 0.3])
    w = v * 3.0 + 0.7
    assert w.to_numpy().dtype == np.float32
    want = (np.float32([0.1, 0.2, 0.3]) * np.float32(3.0)) + np.float32(0.7)
    np.testing.assert_array_equal(w.to_numpy(), want)
    assert isinstance(v.dot(v), np.float32) and isinstance(v.norm(), np.float32)
    assert v.dot(v) == np.float32(np.float32(np.float32(0.1) * np.float32(0.1) + np.float32(0.2) * np.float32(0.2)) + np.float32(0.3) * np.float32(0.3)) \
        or abs(float(v.dot(v)) - 0.14) < 1e-7                                  # summation order is the backend's; the value is f32 either way


def test_vectors_are_value_types(ti):
    """`a = b` copies nothing in Python, but every Taichi operation yields a new value and field / struct reads return copies: the
    reference's `ray_d = it.n_s` followed by in-place edits must not write through (tracer/path_tracer.py:449-453)."""
    f = ti.Vector.field(3, float, shape=(2,))
    f[0] = ti.Vector([1.0, 2.0, 3.0])
    a = f[0]
    a[1] = 9.0
    assert f[0][1] == 2.0
    b = a + 0.0
    b[0] = -1.0
    assert a[0] == 1.0


def test_func_arguments_by_value_unless_template(ti):
    """assets/ti_tests/ref_test.py of the reference: a struct passed to a @ti.func is copied unless the parameter is annotated
    ti.template(); `eval` / `surface_pdf` / `sample_new_ray` mutate `it` through exactly that (tracer/path_tracer.py:424-494)."""
    @ti.dataclass
    class S:
        x: ti.f32
        v: ti.types.vector(3, float)

    @ti.func
    def by_value(s: S):
        s.x = 5.0
        s.v *= -1.0

    @ti.func
    def by_reference(s: ti.template()):
        s.x = 5.0
        s.v *= -1.0                              # the form the reference uses: `it.n_s *= -1` (tracer/path_tracer.py:452-453)

    s = S(x=1.0, v=ti.Vector([1.0, 1.0, 1.0]))
    by_value(s)
    assert s.x == 1.0 and s.v[0] == 1.0
    by_reference(s)
    assert s.x == 5.0 and s.v[0] == -1.0
    # (element stores THROUGH a struct member -- `s.v[0] = ...` -- are not used anywhere on the reference's path; the stand-in's member
    # reads return copies, so it does not support them)


def test_matrix_inverse_and_determinant(ti):
    """Matrix.inverse() of the 3x3 system [e1 e2 -d] solves the triangle test (tracer/tracer_base.py:205, tracer/path_tracer.py:332):
    f32 result within rounding of the float64 inverse on well-conditioned systems, and inverse @ matrix = identity."""
    rng = np.random.default_rng(3)
    for _ in range(50):
        m64 = rng.normal(size=(3, 3))
        if abs(np.linalg.det(m64)) < 0.2:
            continue
        m = ti.Matrix(m64.astype(np.float32).tolist())
        inv = m.inverse().to_numpy()
        assert inv.dtype == np.float32
        np.testing.assert_allclose(inv, np.linalg.inv(m64.astype(np.float32).astype(np.float64)), rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose((m.inverse() @ m).to_numpy(), np.eye(3), atol=2e-5)
        assert abs(float(m.determinant()) - np.linalg.det(m64.astype(np.float32).astype(np.float64))) < 1e-4 * max(1.0, abs(np.linalg.det(m64)))


def test_pow_select_minmax_cast(ti):
    """Elementwise pow with scalar broadcast (Blinn-Phong / Fresnel-blend lobes, bxdf/brdf.py:165-286); select evaluates like a ternary
    per component; ti.max / ti.min drop a NaN operand like fmaxf / fminf (the slab test, tracer/ti_bvh.py:38-53); a float -> int cast
    truncates towards zero."""
    v = ti.Vector([1.0, 2.0, 3.0])
    np.testing.assert_allclose(ti.pow(v, 2.0).to_numpy(), [1.0, 4.0, 9.0], rtol=1e-6)
    np.testing.assert_allclose(ti.pow(2.0, v).to_numpy(), [2.0, 4.0, 8.0], rtol=1e-6)
    assert isinstance(ti.pow(2.0, 0.5), np.float32) and abs(float(ti.pow(2.0, 0.5)) - 2 ** 0.5) < 1e-6
    np.testing.assert_array_equal(ti.select(v > 1.5, v, 0.0).to_numpy(), [0.0, 2.0, 3.0])
    assert ti.select(True, 1, 2) == 1 and ti.select(False, 1.0, 2.0) == np.float32(2.0)
    nan = np.float32("nan")
    assert ti.max(nan, 1.0) == 1.0 and ti.min(1.0, nan) == 1.0 and ti.max(3, 5) == 5
    assert ti.cast(-1.7, int) == -1 and ti.cast(1.7, int) == 1


def test_normalized_has_no_epsilon_and_static_is_transparent(ti):
    """`.normalized()` divides by the plain norm: a zero vector yields non-finite components (the estimator's NaN scrub exists because of
    such samples, renderer/vanilla_renderer.py:119); ti.static(x) is x."""
    z = ti.Vector([0.0, 0.0, 0.0]).normalized().to_numpy()
    assert not np.isfinite(z).any()
    n = ti.Vector([3.0, 0.0, 4.0]).normalized().to_numpy()
    np.testing.assert_allclose(n, [0.6, 0.0, 0.8], rtol=1e-6)
    assert ti.static(True) is True and ti.static(3) == 3
