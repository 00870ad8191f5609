"""Device code of the `pt` path, run WITHOUT a GPU: csrc/pt_shade.cuh (every BRDF / BSDF model) and csrc/pt_trace.cuh (the traversal
over the device BVH layout) are compiled as host C++ by tests/dev_host and held to the same references as their GPU twins in
tests/test_reference_golden.py / tests/test_gpu_parity.py: the reference-generated BxDF tables and the oracle's intersections.
This does not replace the `-m gpu` tests (it is g++ code generation, not ptxas'); it lets the CPU suite catch a wrong formula or a
wrong RNG draw order in the shipped device headers."""
import numpy as np
import pytest

from conftest import load_scene

import os
HERE = os.path.dirname(os.path.abspath(__file__))
BX = os.path.join(HERE, "golden", "reference_bxdf.npz")


def _relerr(x, y):
    x = np.nan_to_num(x, nan=-777.0, posinf=1e30, neginf=-1e30)
    y = np.nan_to_num(y, nan=-777.0, posinf=1e30, neginf=-1e30)
    return float(np.max(np.abs(x - y) / np.maximum(1e-3, np.abs(y))))


def _dev(scene_root, scene, name, **kw):
    from adapt_b200._lib import pack_scene
    from dev_host import DevHostScene
    e, a, o, c = load_scene(scene_root, scene, name, 8, 8, **kw)
    ps = pack_scene(e, a, o, c)
    return DevHostScene(ps, for_vpt=False), ps, (e, a, o, c)


@pytest.mark.parametrize("vset", ["A", "B"])
def test_device_bxdf_models_match_reference_tables(scene_root, vset):
    """eval / pdf / sample of every surface model of allbxdf.xml, device code on the host, against the tables produced by the
    reference's own PathTracer.eval / surface_pdf / sample_new_ray (same tolerances as the GPU twin)."""
    g = np.load(BX)
    dev, _, (e, a, o, c) = _dev(scene_root, "test", "allbxdf.xml")
    n_obj = g[vset + "/pdf"].shape[0]
    assert n_obj == len(o)
    seed, ts = int(g["seed"]), int(g[vset + "/two_sides"])
    for ob in range(n_obj):
        got = dev.bxdf_batch(ob, g[vset + "/n_s"][ob], g[vset + "/n_g"][ob], g[vset + "/incid"][ob], g[vset + "/out"][ob], bool(ts), seed)
        name = str(g["names"][ob])
        assert _relerr(got["eval"], g[vset + "/eval"][ob]) < 5e-4, name
        assert _relerr(got["pdf"], g[vset + "/pdf"][ob]) < 5e-4, name
        assert np.abs(got["s_dir"] - g[vset + "/s_dir"][ob]).max() < 2e-5, name
        assert _relerr(got["s_spec"], g[vset + "/s_spec"][ob]) < 5e-4, name
        assert _relerr(got["s_pdf"], g[vset + "/s_pdf"][ob]) < 5e-4, name
        np.testing.assert_array_equal(got["s_flag"], g[vset + "/s_flag"][ob])


def _random_rays(n, seed, lo=-0.5, hi=6.0):
    rng = np.random.default_rng(seed)
    ro = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    rd = rng.normal(size=(n, 3)); rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    tm = rng.uniform(0.5, 6.0, n).astype(np.float32)
    return ro, rd.astype(np.float32), tm


@pytest.mark.parametrize("scene,name", [("test", "allbxdf.xml"), ("csphere", "balls-mono.xml"), ("cbox", "cbox.xml")])
def test_device_traversal_matches_oracle_intersections(scene_root, oracle_lib, scene, name):
    """trace<> of pt_trace.cuh (near-first stack traversal of the 64-byte nodes, Cramer triangle test, sphere test) against the oracle's
    restatement of ray_intersect / does_intersect, including axis-aligned rays (0 * inf in the slab test)."""
    from oracle.pt_oracle import OracleScene
    dev, ps, _ = _dev(scene_root, scene, name)
    osc = OracleScene(ps)
    ro, rd, tm = _random_rays(40000, 1)
    rd[:300] = np.eye(3, dtype=np.float32)[np.arange(300) % 3] * np.where(np.arange(300) % 2, 1, -1)[:, None]
    g, ref = dev.intersect_batch(ro, rd), osc.intersect_batch(ro, rd)
    same = g["prim"] == ref["prim"]
    # a different primitive is only acceptable as an exact tie (coplanar faces: the boxes of cbox.xml stand ON the floor)
    differ = ~same
    assert ((g["prim"] >= 0) == (ref["prim"] >= 0))[differ].all() or differ.mean() < 5e-4
    tie = differ & (g["prim"] >= 0) & (ref["prim"] >= 0) & (np.abs(g["t"] - ref["t"]) <= 1e-4 * np.maximum(1.0, ref["t"]))
    assert (differ & ~tie).mean() < 5e-4
    np.testing.assert_allclose(g["t"][same], ref["t"][same], rtol=1e-4, atol=2e-5)
    hit = same & (ref["prim"] >= 0)
    np.testing.assert_array_equal(g["obj"][hit], ref["obj"][hit])
    ga, ra = dev.intersect_batch(ro, rd, tm, any_hit=True), osc.intersect_batch(ro, rd, tm, any_hit=True)
    assert (ga["prim"] == ra["prim"]).mean() > 0.9995


def test_division_free_index_arithmetic_is_exact():
    """k_logic turns a 64-bit work id into (sample, pixel slot) and a pixel index into (column, row) with a floating-point estimate and one
    correction step instead of integer divisions, and takes floor-modulo of powers of two with an AND (pt_common.cuh): equal to the plain
    operators on edge values (multiples of the divisor +- 2, ids up to 2^50, the BASELINE film sizes) and on 2 million random cases."""
    import dev_host
    lib = dev_host.load()
    assert lib.dev_host_index_arith_check(7, 2_000_000) == 0
