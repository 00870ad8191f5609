"""Pins the oracle (and, on the GPU, the CUDA path) against golden vectors produced BY THE REFERENCE'S OWN SOURCE.

tests/golden/reference_pt.npz and reference_bxdf.npz were generated in the development container by
tests/golden/make_reference_golden.py: it imports the unmodified reference modules from /root/reference on top of a
pure-Python stand-in for Taichi (tests/golden/ti_shim/) and executes `Renderer.render`
(renderer/vanilla_renderer.py:32-120) and `PathTracer.eval / surface_pdf / sample_new_ray`
(tracer/path_tracer.py:424-494) as Python, drawing random numbers from the PCG32 stream keyed (seed, pixel, sample)
that the oracle and the kernels use.  Nothing here reads /root/reference at run time.

Tolerances: the shim computes in IEEE float32 without FMA contraction, the oracle with it (like Taichi's LLVM
fast-math and nvcc): values agree to ~1e-6 relative, except single pixel-samples that sit on one of the estimator's
thresholds (a shadow ray leaving a curved mesh that re-hits its neighbour triangle at t ~ 1e-4, tracer_base.py:195-197;
Russian roulette) and "flip".  Pixels holding a flipped sample are counted and bounded; all others must agree.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_scene, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
PT = os.path.join(HERE, "golden", "reference_pt.npz")
BX = os.path.join(HERE, "golden", "reference_bxdf.npz")

# tag -> (scene dir, xml) in this repo's scenes/ tree (same files the reference rendered)
SCENES = {
    "cbox": ("cbox", "cbox.xml"), "cbox_b8_uniform": ("cbox", "cbox.xml"), "cbox_point": ("cbox", "cbox-point.xml"),
    "complex": ("cbox", "complex.xml"), "balls_mono": ("csphere", "balls-mono.xml"), "mix_balls": ("csphere", "mix-balls.xml"),
    "balls_multi": ("csphere", "balls-multi.xml"), "allbxdf": ("test", "allbxdf.xml"),
    "allbxdf_nomis_norr": ("test", "allbxdf.xml"), "allbxdf_bvh": ("test", "allbxdf.xml"),
    "textured": ("test", "textured.xml"),          # albedo / normal / bump maps on meshes and spheres
}
# scenes made of Lambertian planes and boxes only have no chaotic samples: exact agreement is demanded there
SMOOTH = {"cbox", "cbox_b8_uniform", "cbox_point"}


def _case(g, scene_root, tag):
    w, h, spp, seed, mb, mis, rr, strat, bvh = (int(x) for x in g[tag + "/meta"])
    scene, name = SCENES[tag]
    e, a, o, c = load_scene(scene_root, scene, name, w, h, max_bounce=mb, use_mis=bool(mis), use_rr=bool(rr),
                            stratified_sampling=bool(strat))
    c["accelerator"] = "bvh" if bvh else "none"
    return (e, a, o, c), spp, seed


def _flip_stats(img, ref):
    d = np.abs(img - ref).sum(-1)
    match = d <= 1e-3 * np.maximum(1.0, np.abs(ref).sum(-1))
    return match, 1.0 - float(match.mean())


@pytest.fixture(scope="module")
def golden():
    return np.load(PT)


@pytest.mark.parametrize("tag", sorted(SCENES))
def test_scene_ingestion_matches_reference_parsers(golden, scene_root, tag):
    """parsers/xml_parser.py::scene_parsing + obj_loader + *_np descriptors of the reference vs the re-hosted ones."""
    (e, a, o, c), _, _ = _case(golden, scene_root, tag)
    g = golden
    np.testing.assert_array_equal(np.asarray(a["primitives"], np.float32), g[tag + "/primitives"])
    np.testing.assert_allclose(a["n_g"], g[tag + "/n_g"], atol=1e-7)
    np.testing.assert_allclose(a["n_s"], g[tag + "/n_s"], atol=1e-7)
    idx = a["indices"] if a["indices"] is not None else []
    np.testing.assert_array_equal(np.asarray(idx, np.int64), g[tag + "/sphere_indices"])
    np.testing.assert_allclose(np.asarray([ob.aabb for ob in o], np.float32), g[tag + "/obj_aabb"], atol=1e-6)
    np.testing.assert_array_equal([ob.tri_num for ob in o], g[tag + "/obj_tri_num"])
    np.testing.assert_array_equal([ob.emitter_ref_id for ob in o], g[tag + "/emitter_ref"])
    np.testing.assert_allclose(np.asarray([em.intensity for em in e], np.float32), g[tag + "/emitter_intensity"], rtol=1e-6)
    np.testing.assert_allclose(np.asarray([em.inv_area for em in e], np.float32), g[tag + "/emitter_inv_area"], rtol=1e-5)
    for key, attr in (("bxdf_kd", "k_d"), ("bxdf_ks", "k_s"), ("bxdf_kg", "k_g")):
        np.testing.assert_allclose(np.asarray([getattr(ob.bsdf, attr) for ob in o], np.float32), g[tag + "/" + key], rtol=1e-6, atol=1e-7)
    # camera (tracer_base.py:36-75)
    from adapt_b200._lib import pack_scene
    ps = pack_scene(e, a, o, c)
    np.testing.assert_allclose(np.asarray(ps.desc.cam_r[:9], np.float32).reshape(3, 3), g[tag + "/cam_r"], atol=1e-6)
    np.testing.assert_allclose(np.asarray(ps.desc.cam_t[:3], np.float32), g[tag + "/cam_t"], atol=1e-6)
    assert abs(1.0 / ps.desc.inv_focal - float(g[tag + "/focal"][0])) < 1e-3
    if tag + "/uvs" in g.files:
        np.testing.assert_allclose(a["uvs"], g[tag + "/uvs"], atol=1e-7)
    for key in ("albedo", "normal", "bump"):
        if tag + "/tex_" + key in g.files:       # (has texture, w, h, scale_u, scale_v) per object; atlas offsets are the packer's business
            rows = []
            for ob in o:
                t = ob.texture_group.get(key)
                rows.append([0, 0, 0, 1, 1] if t is None else [1, t.w, t.h, t.scale_u, t.scale_v])
            np.testing.assert_allclose(np.asarray(rows, np.float32), g[tag + "/tex_" + key])
            assert c["packed_textures"][key] is not None and ps.desc.tex_size[("albedo", "normal", "bump").index(key)] > 0


@pytest.mark.parametrize("tag", sorted(SCENES))
def test_oracle_matches_reference_render(golden, scene_root, tag):
    from adapt_b200._lib import pack_scene
    from oracle.pt_oracle import OracleScene
    (e, a, o, c), spp, seed = _case(golden, scene_root, tag)
    acc, cn = OracleScene(pack_scene(e, a, o, c, seed=seed)).render(spp)
    ref = golden[tag + "/color"]
    assert acc.shape == ref.shape and cn["paths"] == ref.shape[0] * ref.shape[1] * spp
    match, flipped = _flip_stats(acc, ref)
    if tag in SMOOTH:
        assert flipped == 0.0 and rel_l2(acc, ref) < 2e-5
    else:
        assert flipped < 0.04, f"{flipped:.2%} of pixels hold a sample that differs"
        assert rel_l2(acc[match], ref[match]) < 5e-4
        np.testing.assert_allclose(acc.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-2)


def _bxdf_records(scene_root):
    from adapt_b200._lib import adapt_bxdf, pack_scene
    e, a, o, c = load_scene(scene_root, "test", "allbxdf.xml", 4, 4)
    bx = pack_scene(e, a, o, c).keep["bxdfs"]
    return [adapt_bxdf.from_buffer_copy(bx[k].tobytes()) for k in range(len(o))]


def _relerr(x, y):
    x = np.nan_to_num(x, nan=-777.0, posinf=1e30, neginf=-1e30)
    y = np.nan_to_num(y, nan=-777.0, posinf=1e30, neginf=-1e30)
    return float(np.max(np.abs(x - y) / np.maximum(1e-3, np.abs(y))))


@pytest.mark.parametrize("vset", ["A", "B"])
def test_oracle_bxdf_models_match_reference(oracle_lib, scene_root, vset):
    """eval / pdf / sample of every BRDF and BSDF type against the reference's own functions (bxdf/brdf.py, bxdf/bsdf.py)."""
    g = np.load(BX)
    lib = oracle_lib
    fp = C.POINTER(C.c_float)
    lib.oracle_bxdf_eval2.argtypes = [C.c_void_p, fp, fp, fp, fp, C.c_float, C.c_int, fp, fp]
    lib.oracle_bxdf_sample2.argtypes = [C.c_void_p, fp, fp, fp, C.c_float, C.c_int, C.c_uint64, C.c_uint32, fp, fp, fp, C.POINTER(C.c_int32)]
    _p = lambda x: x.ctypes.data_as(fp)      # noqa: E731
    recs = _bxdf_records(scene_root)
    n_obj, N = g[vset + "/pdf"].shape
    assert n_obj == len(recs)
    wi, seed, ts = float(g["world_ior"]), int(g["seed"]), int(g[vset + "/two_sides"])
    types_seen = set()
    for ob, b in enumerate(recs):
        types_seen.add((b.kind, b.type))
        ev = np.zeros((N, 3), np.float32); pdf = np.zeros(N, np.float32)
        sd = np.zeros((N, 3), np.float32); ss = np.zeros((N, 3), np.float32); sp = np.zeros(N, np.float32); sf = np.zeros(N, np.int32)
        for k in range(N):
            ns, ng, ii, oo = (g[f"{vset}/{x}"][ob, k].copy() for x in ("n_s", "n_g", "incid", "out"))
            spec = np.zeros(3, np.float32); p = C.c_float(0); fl = C.c_int32(0)
            lib.oracle_bxdf_eval2(C.byref(b), _p(ns), _p(ng), _p(ii), _p(oo), wi, ts, _p(spec), C.byref(p))
            ev[k], pdf[k] = spec, p.value
            d = np.zeros(3, np.float32); s = np.zeros(3, np.float32)
            lib.oracle_bxdf_sample2(C.byref(b), _p(ns), _p(ng), _p(ii), wi, ts, seed, k, _p(d), _p(s), C.byref(p), C.byref(fl))
            sd[k], ss[k], sp[k], sf[k] = d, s, p.value, fl.value
        name = str(g["names"][ob])
        assert _relerr(ev, g[vset + "/eval"][ob]) < 5e-4, name
        assert _relerr(pdf, g[vset + "/pdf"][ob]) < 5e-4, name
        assert np.abs(sd - g[vset + "/s_dir"][ob]).max() < 2e-5, name
        assert _relerr(ss, g[vset + "/s_spec"][ob]) < 5e-4, name
        assert _relerr(sp, g[vset + "/s_pdf"][ob]) < 5e-4, name
        np.testing.assert_array_equal(sf, g[vset + "/s_flag"][ob])
    # every surface model of the path is covered: 8 BRDF types, det-refraction and Lambertian-transmission BSDFs
    assert {(0, t) for t in range(8)} <= types_seen and {(1, 0), (1, 1)} <= types_seen


def test_bvh_cpp_dropin_signature(scene_root):
    """Boundary #2: adapt_b200.bvh_cpp.bvh_build has the reference's signature and array shapes (tracer/bvh/bvh.cpp:274-296);
    the reference's own traversal code consumed its output when the allbxdf_bvh fixture was generated."""
    from adapt_b200 import bvh_cpp
    e, a, o, c = load_scene(scene_root, "test", "allbxdf.xml", 4, 4)
    obj_info = np.zeros((2, len(o)), np.int32)
    for k, ob in enumerate(o):
        obj_info[0, k] = ob.meshes.shape[0]; obj_info[1, k] = ob.type
    lo = np.min([ob.aabb[0] for ob in o], axis=0).astype(np.float32) - 0.1
    hi = np.max([ob.aabb[1] for ob in o], axis=0).astype(np.float32) + 0.1
    bvh_minmax, node_minmax, bvh_info, node_info = bvh_cpp.bvh_build(a["primitives"], obj_info, lo, hi)
    n_ref, n_node = bvh_info.size // 2, node_info.size // 3
    assert bvh_minmax.shape == (n_ref * 6,) and node_minmax.shape == (n_node * 6,) and bvh_minmax.dtype == np.float32
    assert n_ref == a["primitives"].shape[0] and node_info.dtype == np.int32 and bvh_info.dtype == np.int32
    with pytest.raises(RuntimeError):
        bvh_cpp.bvh_build(a["primitives"], obj_info[:1], lo, hi)


# ------------------------------------------------------------------------------------------------ GPU: the CUDA path against the same vectors
@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(SCENES))
def test_cuda_matches_reference_render(golden, scene_root, tag):
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from adapt_b200.build import build
    build()
    from adapt_b200.renderer.vanilla_renderer import Renderer
    (e, a, o, c), spp, seed = _case(golden, scene_root, tag)
    r = Renderer(e, a, o, c, seed=seed)
    r.render_batch(spp)
    acc = r.color.to_numpy()
    ref = golden[tag + "/color"]
    assert np.isfinite(acc).all() and r.stats()["paths"] == ref.shape[0] * ref.shape[1] * spp
    match, flipped = _flip_stats(acc, ref)
    if tag in SMOOTH:
        assert flipped == 0.0 and rel_l2(acc, ref) < 2e-5
    else:
        assert flipped < 0.07, f"{flipped:.2%} of pixels hold a sample that differs"
        assert rel_l2(acc[match], ref[match]) < 5e-4
        np.testing.assert_allclose(acc.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-2)


def _ref_bvh_cpp():
    """The reference's own pybind11 module, compiled from tracer/bvh/bvh.cpp by `make -C oracle ref` (oracle/_ref/)."""
    import glob
    import importlib.util
    hits = glob.glob(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "bvh_cpp*.so"))
    if not hits:
        pytest.skip("oracle/_ref/bvh_cpp*.so not built (needs the reference tree: make -C oracle ref)")
    spec = importlib.util.spec_from_file_location("bvh_cpp", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("scene,name,big", [("cbox", "cbox.xml", None), ("test", "allbxdf.xml", None), ("cbox", "bunny90k.xml", "bunny90k")])
def test_oracle_bvh_builder_equals_compiled_reference(scene_root, scene, name, big):
    """The restated SAH builder (oracle/pt_oracle.cpp) against the reference's own bvh.cpp compiled from its sources:
    all four arrays bit-identical (bvh_minmax, node_minmax, bvh_info, node_info), 34 ... 89 900 primitives."""
    ref = _ref_bvh_cpp()
    from oracle.pt_oracle import bvh_build as oracle_build
    if big:
        from adapt_b200.scenes import ensure_big_meshes
        ensure_big_meshes(scene_root, (big,))
    e, a, o, c = load_scene(scene_root, scene, name, 8, 8)
    obj_info = np.zeros((2, len(o)), np.int32)
    for k, ob in enumerate(o):
        obj_info[0, k] = ob.meshes.shape[0]; obj_info[1, k] = ob.type
    cam_t = np.float32(c["transform"][1])
    lo = np.minimum(np.min([ob.aabb[0] for ob in o], axis=0), cam_t).astype(np.float32) - np.float32(0.1)     # path_tracer.py:130-138
    hi = np.maximum(np.max([ob.aabb[1] for ob in o], axis=0), cam_t).astype(np.float32) + np.float32(0.1)
    prims = np.ascontiguousarray(a["primitives"], np.float32)
    want = ref.bvh_build(prims, obj_info, lo, hi)
    got = oracle_build(prims, obj_info, lo, hi)
    for w, g, what in zip(want, got, ("bvh_minmax", "node_minmax", "bvh_info", "node_info")):
        np.testing.assert_array_equal(np.asarray(g).ravel(), np.asarray(w).ravel(), err_msg=what)
    # the drop-in of boundary #2 returns the same shapes and a tree over the same primitives (its own SAH builder)
    from adapt_b200 import bvh_cpp as dropin
    mine = dropin.bvh_build(prims, obj_info, lo, hi)
    assert mine[0].shape == np.asarray(want[0]).shape and mine[2].shape == np.asarray(want[2]).shape
    assert sorted(mine[2].reshape(-1, 2)[:, 1].tolist()) == sorted(np.asarray(want[2]).reshape(-1, 2)[:, 1].tolist())


def test_watermark_matches_reference():
    """utils/watermark.py of the reference (stamp bitmap, placement, quantile normalisation, crop) on a synthetic film."""
    from adapt_b200.utils.watermark import apply_watermark, water_mark
    g = np.load(os.path.join(HERE, "golden", "reference_watermark.npz"))
    np.testing.assert_array_equal(water_mark, g["stamp"])

    class Film:
        def to_numpy(self):
            return g["film"].copy()

    class Rdr:
        pixels = Film()
    for tag, crop, norm, wm in (("plain", False, 0.0, True), ("norm", False, 0.99, True), ("nostamp", False, 0.0, False), ("crop", True, 0.0, True)):
        r = Rdr()
        r.do_crop, r.start_x, r.end_x, r.start_y, r.end_y = crop, 5, 30, 10, 60
        np.testing.assert_allclose(apply_watermark(r, norm, False, wm), g[tag], rtol=1e-6, err_msg=tag)


@pytest.mark.gpu
@pytest.mark.parametrize("vset", ["A", "B"])
def test_cuda_bxdf_models_match_reference(scene_root, vset):
    """The device code of every BRDF / BSDF (csrc/pt_shade.cuh, through adapt_bxdf_batch) against the reference's own
    eval / surface_pdf / sample_new_ray tables: same tolerances as the oracle's half of this test."""
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from adapt_b200.build import build
    build()
    from adapt_b200.renderer.vanilla_renderer import Renderer
    g = np.load(BX)
    e, a, o, c = load_scene(scene_root, "test", "allbxdf.xml", 4, 4)
    r = Renderer(e, a, o, c)
    n_obj, N = g[vset + "/pdf"].shape
    assert n_obj == len(o)
    seed, ts = int(g["seed"]), int(g[vset + "/two_sides"])
    for ob in range(n_obj):
        got = r.bxdf_batch(ob, g[vset + "/n_s"][ob], g[vset + "/n_g"][ob], g[vset + "/incid"][ob], g[vset + "/out"][ob], bool(ts), seed)
        name = str(g["names"][ob])
        assert _relerr(got["eval"], g[vset + "/eval"][ob]) < 5e-4, name
        assert _relerr(got["pdf"], g[vset + "/pdf"][ob]) < 5e-4, name
        assert np.abs(got["s_dir"] - g[vset + "/s_dir"][ob]).max() < 2e-5, name
        assert _relerr(got["s_spec"], g[vset + "/s_spec"][ob]) < 5e-4, name
        assert _relerr(got["s_pdf"], g[vset + "/s_pdf"][ob]) < 5e-4, name
        np.testing.assert_array_equal(got["s_flag"], g[vset + "/s_flag"][ob])


@pytest.mark.gpu
def test_cuda_tiny_pool_with_several_material_groups(scene_root):
    """Regression for a bug the SIMT emulator found (tests/test_wavefront_emulated.py): scenes with several material groups run one
    k_logic launch per group, and with a pool of a few hundred slots the shadow queue's segments used to overflow into each other.
    Kept at the end of the last GPU test file on purpose (it was added after the round's GPU time was spent)."""
    from adapt_b200._lib import pack_scene
    from adapt_b200.renderer.vanilla_renderer import Renderer
    from oracle.pt_oracle import OracleScene
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 16, 16)
    r = Renderer(e, a, o, c, seed=5, pool_size=256)
    r.render_batch(2)
    img = r.color.to_numpy()
    ref, _ = OracleScene(pack_scene(e, a, o, c, seed=5)).render(2)
    match, flipped = _flip_stats(img, ref)
    assert flipped <= 0.01 and rel_l2(img[match], ref[match]) < 1e-4
