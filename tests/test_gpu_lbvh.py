"""Device BVH builders -- GPU half: the tree built by bvh_device.cu (through adapt_create with bvh_builder = 1, the linear BVH, or 2, the
level-synchronous binned SAH and the default, and read back with adapt_bvh_export) is held to the CPU emulation of the same per-element steps bit for bit, traced against the host-SAH handle,
and rendered: tree shape must not change a result (closest hit is unique)."""
import os

import numpy as np
import pytest

from conftest import load_scene, rel_l2
from lbvh_host import build_tree, validate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Renderer():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from adapt_b200.build import build
    build()
    from adapt_b200.renderer.vanilla_renderer import Renderer as R
    return R


def _tables(a, objs):
    prims = a["primitives"].reshape(-1, 9)
    n = prims.shape[0]
    sph = np.zeros(n, np.uint8)
    if a["indices"] is not None:
        sph[np.asarray(a["indices"], np.int64)] = 1
    return prims, sph


def _rays(prims, n, seed):
    rng = np.random.default_rng(seed)
    v = prims.reshape(-1, 3, 3)
    lo, hi = v.min((0, 1)) - 0.5, v.max((0, 1)) + 0.5
    ro = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    w = rng.dirichlet([1.0, 1.0, 1.0], n)                            # a point inside a random primitive (away from shared edges)
    tgt = (v[rng.integers(0, v.shape[0], n)] * w[:, :, None]).sum(1)
    rd = tgt - ro
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    return ro, rd.astype(np.float32)


SCENES = [("cbox", "cbox.xml"), ("csphere", "balls-mono.xml"), ("test", "allbxdf.xml"), ("cbox", "bunny90k.xml")]


def _load(scene_root, scene, name, size):
    if name == "bunny90k.xml":
        from adapt_b200.scenes import ensure_big_meshes
        ensure_big_meshes(scene_root, ("bunny90k",))
    return load_scene(scene_root, scene, name, size, size)


@pytest.mark.parametrize("builder", ["lbvh", "sah_device"])
@pytest.mark.parametrize("scene,name", SCENES)
def test_device_tree_equals_emulated_tree(Renderer, scene_root, scene, name, builder):
    e, a, o, c = _load(scene_root, scene, name, 32)
    r = Renderer(e, a, o, c, bvh_builder=builder)
    ex = r.bvh_export()
    assert ex["builder"] == {"lbvh": 1, "sah_device": 2}[builder] and ex["n_prims"] == a["primitives"].shape[0]
    prims, sph = _tables(a, o)
    rc, depth = validate(ex["nodes"], ex["prims"], prims, sph)
    assert rc == 0 and depth == ex["depth"]
    # scenes of <= 64 primitives are traced through the 8-wide tree: the device SAH builder then keeps leaves of <= 3 and collapses it too
    wide = ex["n_nodes8"] > 0
    assert not (wide and builder == "lbvh")
    ref = build_tree(prims, sph, max_leaf=3 if wide else 4, builder=builder, order_seed=3, eight=wide)   # emulated threads in a shuffled order
    if wide:
        assert np.array_equal(ex["nodes8"], ref["nodes8"]) and ex["depth8"] == ref["depth8"]
    assert ex["n_nodes"] == ref["nodes"].shape[0] and ex["depth"] == ref["depth"]
    # geometry words of the records and the whole node array, bit for bit (object / class words depend on the scene tables)
    assert np.array_equal(ex["prims"][:, :10].view(np.uint32), ref["prims"][:, :10].view(np.uint32))
    assert np.array_equal(ex["nodes"].view(np.uint32), ref["nodes"].view(np.uint32))
    r.close()


@pytest.mark.parametrize("builder", ["lbvh", "sah_device"])
@pytest.mark.parametrize("scene,name", SCENES)
def test_same_hits_and_same_image_as_host_sah_tree(Renderer, scene_root, scene, name, builder):
    size, spp = 64, 4
    e, a, o, c = _load(scene_root, scene, name, size)
    r_l = Renderer(e, a, o, c, seed=2, bvh_builder=builder)
    r_s = Renderer(e, a, o, c, seed=2, bvh_builder="sah")
    assert r_s.bvh_export(arrays=False)["builder"] == 3             # "sah" = the host builder, whatever the default is
    prims, _ = _tables(a, o)
    ro, rd = _rays(prims, 20000, 3)
    h_l, h_s = r_l.intersect_batch(ro, rd), r_s.intersect_batch(ro, rd)
    assert np.array_equal(h_l["t"], h_s["t"])                       # the same primitive test decides, whatever the tree
    hit = h_s["prim"] >= 0
    assert np.array_equal(h_l["prim"] >= 0, hit)
    assert (h_l["prim"][hit] != h_s["prim"][hit]).mean() < 5e-3     # exact ties (same t) on shared edges may resolve either way
    s_l, s_s = r_l.intersect_batch(ro, rd, any_hit=True), r_s.intersect_batch(ro, rd, any_hit=True)
    assert np.array_equal(s_l["prim"], s_s["prim"])
    r_l.render_batch(spp); r_s.render_batch(spp)
    img_l, img_s = r_l.pixels.to_numpy(), r_s.pixels.to_numpy()
    assert np.isfinite(img_l).all()
    assert rel_l2(img_l, img_s) < 1e-5                              # summation order of the atomics only
    st_l, st_s = r_l.stats(), r_s.stats()
    assert st_l["rays_closest"] == st_s["rays_closest"] and st_l["paths"] == st_s["paths"] == size * size * spp
    r_l.close(); r_s.close()


@pytest.mark.parametrize("builder", ["lbvh", "sah", "sah_device"])
def test_update_geometry_rebuilds_and_renders_like_a_fresh_scene(Renderer, scene_root, builder):
    """adapt_update_geometry: move one mesh, rebuild (on the device for "lbvh" and "sah_device"), and get the image a renderer
    created on the moved geometry gives."""
    size, spp = 64, 4
    e, a, o, c = _load(scene_root, "cbox", "bunny90k.xml", size)
    r = Renderer(e, a, o, c, seed=5, bvh_builder=builder)
    r.render_batch(1)                                               # work in flight before the update
    first = r.bvh_export(arrays=False)
    moved = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    big = max(range(len(o)), key=lambda k: o[k].tri_num)            # the 90k-triangle mesh
    p0 = sum(x.tri_num for x in o[:big]); p1 = p0 + o[big].tri_num
    moved["primitives"][p0:p1] += np.float32([0.35, 0.2, -0.3])      # a translation keeps n_g / n_s valid
    r.update_geometry(moved["primitives"], moved["n_g"], moved["n_s"])
    r.reset_accumulation()
    r.render_batch(spp)
    img = r.pixels.to_numpy()
    again = r.bvh_export()
    print(f"[{builder}] first build {first['build_ms']:.2f} ms, rebuild {again['build_ms']:.2f} ms, {again['n_nodes']} nodes")
    prims, sph = _tables(moved, o)
    assert validate(again["nodes"], again["prims"], prims, sph)[0] == 0
    if builder != "sah":
        ref = build_tree(prims, sph, max_leaf=4, builder=builder)
        assert np.array_equal(again["nodes"].view(np.uint32), ref["nodes"].view(np.uint32))
    fresh = Renderer(e, moved, o, c, seed=5, bvh_builder=builder)
    fresh.render_batch(spp)
    ref_img = fresh.pixels.to_numpy()
    assert np.isfinite(img).all() and rel_l2(img, ref_img) < 1e-5
    e0, a0, o0, c0 = _load(scene_root, "cbox", "bunny90k.xml", size)
    still = Renderer(e0, a0, o0, c0, seed=5, bvh_builder=builder)
    still.render_batch(spp)
    assert rel_l2(still.pixels.to_numpy(), ref_img) > 1e-2          # the move is visible: the update really changed the scene
    r.close(); fresh.close(); still.close()


@pytest.mark.parametrize("builder", ["sah", "lbvh"])
def test_refit_geometry_keeps_the_tree_and_renders_like_a_fresh_scene(Renderer, scene_root, builder, monkeypatch):
    """adapt_refit_geometry: deform and move the 90k-triangle mesh, recompute only the boxes on the device (same node count, same child
    codes), and get the image of a renderer created on the moved geometry; the refitted boxes enclose every primitive (validate)."""
    monkeypatch.setenv("ADAPT_TRACE_MODE", "1")                     # the refit path is the binary tree's
    size, spp = 64, 4
    e, a, o, c = _load(scene_root, "cbox", "bunny90k.xml", size)
    r = Renderer(e, a, o, c, seed=5, bvh_builder=builder)
    r.render_batch(1)
    first = r.bvh_export()
    moved = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    big = max(range(len(o)), key=lambda k: o[k].tri_num)
    p0 = sum(x.tri_num for x in o[:big]); p1 = p0 + o[big].tri_num
    pr = moved["primitives"][p0:p1]
    centre = pr.reshape(-1, 3).mean(axis=0)
    pr[:] = (pr - centre) * np.float32([1.08, 0.94, 1.03]) + centre + np.float32([0.15, 0.1, -0.2])     # squash + move: not a rigid motion
    e1, e2 = pr[:, 1] - pr[:, 0], pr[:, 2] - pr[:, 0]
    ng = np.cross(e1, e2); ng /= np.maximum(np.linalg.norm(ng, axis=1, keepdims=True), 1e-20)
    moved["n_g"][p0:p1] = ng.astype(np.float32)
    r.update_geometry(moved["primitives"], moved["n_g"], moved["n_s"], refit=True)
    r.reset_accumulation()
    r.render_batch(spp)
    img = r.pixels.to_numpy()
    again = r.bvh_export()
    print(f"[{builder}] build {first['build_ms']:.2f} ms, refit {again['build_ms']:.3f} ms, {again['n_nodes']} nodes")
    assert again["n_nodes"] == first["n_nodes"]
    np.testing.assert_array_equal(again["nodes"].view(np.uint32)[:, 12:14], first["nodes"].view(np.uint32)[:, 12:14])      # same topology
    prims, sph = _tables(moved, o)
    assert validate(again["nodes"], again["prims"], prims, sph)[0] == 0
    fresh = Renderer(e, moved, o, c, seed=5, bvh_builder=builder)
    fresh.render_batch(spp)
    assert np.isfinite(img).all() and rel_l2(img, fresh.pixels.to_numpy()) < 1e-5
    assert again["build_ms"] < 5.0


@pytest.mark.parametrize("scene,name", [("test", "allbxdf.xml"), ("cbox", "bunny90k.xml")])
def test_device_cw8_tree_equals_emulated_tree_and_renders_like_the_host_tree(Renderer, scene_root, scene, name, monkeypatch):
    """The compressed 8-wide tree collapsed on the device from the device-SAH hierarchy (ADAPT_TRACE_MODE=3 + bvh_builder "sah_device"):
    80-byte nodes and re-ordered records equal the CPU harness's bit for bit, an independent decoder traces it like brute force, and the
    image equals the one rendered through the host builder's 8-wide tree."""
    from lbvh_host import cw8_trace_check
    monkeypatch.setenv("ADAPT_TRACE_MODE", "3")
    size, spp = 64, 4
    e, a, o, c = _load(scene_root, scene, name, size)
    r_d = Renderer(e, a, o, c, seed=4, bvh_builder="sah_device")
    ex = r_d.bvh_export()
    assert ex["builder"] == 2 and ex["n_nodes8"] > 0 and ex["depth8"] >= 1
    prims, sph = _tables(a, o)
    ref = build_tree(prims, sph, max_leaf=3, builder="sah_device", eight=True, order_seed=2)
    assert ex["n_nodes8"] == ref["nodes8"].shape[0] and ex["depth8"] == ref["depth8"]
    assert np.array_equal(ex["nodes8"], ref["nodes8"])
    assert np.array_equal(ex["nodes"].view(np.uint32), ref["nodes"].view(np.uint32))
    assert np.array_equal(ex["prims"][:, :10].view(np.uint32), ref["prims"][:, :10].view(np.uint32))
    ro, rd = _rays(prims, 3000, 5)
    rc, t8, _ = cw8_trace_check(ex["nodes8"], ex["prims"], prims, sph, ro, rd)
    assert rc == 0
    h_d = r_d.intersect_batch(ro, rd)                               # single-ray hook: the binary tree over the same records
    hit8 = t8 < 1e7                                                  # the CPU decoder rounds without FMA contraction: distances agree to a few ulp
    both = hit8 & (h_d["prim"] >= 0)
    assert (hit8 != (h_d["prim"] >= 0)).mean() < 2e-3              # grazing rays may fall either way between the two roundings
    assert np.allclose(h_d["t"][both], t8[both], rtol=1e-4, atol=0) and (np.abs(h_d["t"][both] - t8[both]) > 1e-5 * t8[both]).mean() < 2e-3
    r_h = Renderer(e, a, o, c, seed=4, bvh_builder="sah")
    assert r_h.bvh_export(arrays=False)["n_nodes8"] > 0
    r_d.render_batch(spp); r_h.render_batch(spp)
    img_d, img_h = r_d.pixels.to_numpy(), r_h.pixels.to_numpy()
    assert np.isfinite(img_d).all() and rel_l2(img_d, img_h) < 1e-5
    assert r_d.stats()["rays_closest"] == r_h.stats()["rays_closest"]
    r_d.close(); r_h.close()


def test_default_builder_is_the_device_sah_builder(Renderer, scene_root, monkeypatch):
    monkeypatch.delenv("ADAPT_BVH_BUILDER", raising=False)
    e, a, o, c = _load(scene_root, "cbox", "bunny90k.xml", 32)
    r = Renderer(e, a, o, c)
    ex = r.bvh_export(arrays=False)
    assert ex["builder"] == 2 and ex["build_ms"] < 15.0
    r.close()


def test_default_handle_falls_back_to_the_host_builder(Renderer, scene_root, monkeypatch):
    """A handle whose builder nobody chose builds on the host when the device tree cannot be used (here: a test hook declares it deeper than
    the traversal stack); a handle that asked for the device builder gets the error instead."""
    monkeypatch.delenv("ADAPT_BVH_BUILDER", raising=False)
    monkeypatch.setenv("ADAPT_TEST_DEVICE_BUILD_TOO_DEEP", "1")
    e, a, o, c = _load(scene_root, "test", "allbxdf.xml", 32)
    r = Renderer(e, a, o, c, seed=1)
    assert r.bvh_export(arrays=False)["builder"] == 3
    r.render_batch(2)
    img = r.pixels.to_numpy()
    r.close()
    with pytest.raises(Exception, match="deeper than the traversal stack"):
        Renderer(e, a, o, c, seed=1, bvh_builder="sah_device")
    monkeypatch.delenv("ADAPT_TEST_DEVICE_BUILD_TOO_DEEP")
    r2 = Renderer(e, a, o, c, seed=1, bvh_builder="sah")
    r2.render_batch(2)
    assert rel_l2(img, r2.pixels.to_numpy()) < 1e-6
    r2.close()
