"""GPU parity tests: the sm_100a wavefront path tracer (through the C ABI / Renderer class) against the
CPU oracle on the same seeded inputs.  Tolerance: north_star asks for <= 2e-3 relative L2 on the HDR
mean buffer; with the shared counter-based RNG the two sides agree far better than that, the residue
being fp-rounding "flips" of individual paths (threshold tests such as t > 1e-4, Russian roulette)."""
import os

import numpy as np
import pytest

from conftest import load_scene, rel_l2

pytestmark = pytest.mark.gpu

TOL = 2e-3          # north_star tolerance on relative L2 of the HDR buffer


@pytest.fixture(scope="module")
def Renderer():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from adapt_b200.build import build
    build()
    from adapt_b200.renderer.vanilla_renderer import Renderer as R
    return R


def _oracle(e, a, o, c, seed, **kw):
    from adapt_b200._lib import pack_scene
    from oracle.pt_oracle import OracleScene
    return OracleScene(pack_scene(e, a, o, c, seed=seed), **kw)


def test_native_library_is_the_one_loaded(Renderer, scene_root):
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 16, 16)
    r = Renderer(e, a, o, c)
    r.render_batch(1)
    assert r.pixels.to_numpy().shape == (16, 16, 3)
    assert any("libadapt_b200.so" in line for line in open("/proc/self/maps"))
    st = r.stats()
    assert st["kernel_launches"] >= 3 and st["paths"] == 256


SCENES = [("cbox", "cbox.xml", 0), ("csphere", "balls-mono.xml", 0), ("test", "allbxdf.xml", 3)]


def _flip_stats(img, ref):
    """The reference estimator is chaotic at a few thresholds (e.g. a scattered ray re-hitting its own
    sphere at t ~ 1e-4, tracer_base.py:195-197): one ulp decides whether such a sample survives, so no two
    builds -- of the reference itself, of the oracle, or of this library -- agree on those samples.
    Pixels holding such a "flipped" sample are counted; all other pixels must agree to fp rounding."""
    d = np.abs(img - ref).sum(-1)
    match = d <= 1e-3 * np.maximum(1.0, np.abs(ref).sum(-1))
    return match, 1.0 - float(match.mean())


@pytest.mark.parametrize("scene,name,seed", SCENES)
def test_shared_rng_parity(Renderer, scene_root, scene, name, seed):
    size, spp = 96, 16
    e, a, o, c = load_scene(scene_root, scene, name, size, size)
    r = Renderer(e, a, o, c, seed=seed)
    r.render_batch(spp)
    img = r.pixels.to_numpy()
    st = r.stats()
    acc, cn = _oracle(e, a, o, c, seed).render(spp)
    ref = acc / spp
    assert np.isfinite(img).all()
    match, flipped = _flip_stats(img, ref)
    assert flipped < 0.05                                       # a few % of pixels contain a flipped sample at most
    assert rel_l2(img[match], ref[match]) < 1e-4                # everything else agrees to fp rounding
    # whole-image relative L2 (north_star tolerance).  allbxdf has a directly visible sphere light whose NEE rays are exactly the chaotic
    # case above, so at 16 spp single flips carry emitter-sized radiance: its whole-image figure is asserted at convergence
    # (tests/test_gpu_baseline_configs.py::test_allbxdf_converged_whole_image)
    if scene != "test":
        assert rel_l2(img, ref) < TOL
    assert st["paths"] == cn["paths"] == size * size * spp
    # closest-hit rays: the GPU skips the reference's unused trace after the last bounce
    assert abs(st["rays_closest"] - cn["rays_closest_useful"]) <= 2e-3 * cn["rays_closest_useful"]
    assert st["rays_shadow"] <= cn["rays_shadow"]               # zero-payload shadow rays are not traced


def test_single_sample_flip_rate(Renderer, scene_root):
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 128, 128)
    r = Renderer(e, a, o, c, seed=0)
    r.render_batch(1)
    img = r.pixels.to_numpy()
    ref, _ = _oracle(e, a, o, c, 0).render(1)
    _, flips = _flip_stats(img, ref)
    assert flips < 1e-2, f"{flips:.4%} of pixel-samples differ"


def test_converged_image_within_tolerance(Renderer, scene_root):
    """High spp on a small film: differences must keep shrinking (no bias), well inside 2e-3."""
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 32, 32)
    spp = 1024
    r = Renderer(e, a, o, c, seed=5)
    r.render_batch(spp)
    img = r.pixels.to_numpy()
    acc, _ = _oracle(e, a, o, c, 5).render(spp)
    assert rel_l2(img, acc / spp) < 1e-3


def test_independent_seeds_agree_statistically(Renderer, scene_root):
    """Different seeds on the two sides: only the estimator's expectation is shared."""
    e, a, o, c = load_scene(scene_root, "test", "allbxdf.xml", 24, 24)
    spp = 2048
    r = Renderer(e, a, o, c, seed=101)
    r.render_batch(spp)
    img = r.pixels.to_numpy()
    acc, _ = _oracle(e, a, o, c, 202).render(spp)
    ref = acc / spp
    np.testing.assert_allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=1.5e-2)
    # 4x4 box-filtered images agree to a few percent
    pool = lambda x: x.reshape(6, 4, 6, 4, 3).mean(axis=(1, 3))     # noqa: E731
    assert rel_l2(pool(img), pool(ref)) < 5e-2


def test_golden_fixtures(Renderer, scene_root):
    from golden.make_golden import CASES
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_small.npz"))
    for tag, scene, name, size, spp, seed in CASES:
        e, a, o, c = load_scene(scene_root, scene, name, size, size)
        r = Renderer(e, a, o, c, seed=seed)
        r.render_batch(spp)
        assert rel_l2(r.pixels.to_numpy(), g[tag]) < TOL, tag


def _random_rays(n, seed, lo=(0.1, 0.1, 0.1), hi=(5.4, 5.4, 5.5)):
    rng = np.random.default_rng(seed)
    ro = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    rd = rng.normal(size=(n, 3)).astype(np.float32)
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    return ro, rd, rng.uniform(0.3, 6.0, n).astype(np.float32)


def test_intersect_stage_parity_small(Renderer, scene_root):
    e, a, o, c = load_scene(scene_root, "test", "allbxdf.xml", 16, 16)
    r = Renderer(e, a, o, c)
    osc = _oracle(e, a, o, c, 0)
    ro, rd, tm = _random_rays(100000, 0)
    # axis-aligned and grazing rays: the slab test sees 0 * inf there
    rd[:300] = np.eye(3, dtype=np.float32)[np.arange(300) % 3] * np.where(np.arange(300) % 2, 1, -1)[:, None]
    g, ref = r.intersect_batch(ro, rd), osc.intersect_batch(ro, rd)
    same = g["prim"] == ref["prim"]
    assert same.mean() > 0.9995
    np.testing.assert_allclose(g["t"][same], ref["t"][same], rtol=1e-4, atol=2e-5)
    hit = same & (ref["prim"] >= 0)
    np.testing.assert_array_equal(g["obj"][hit], ref["obj"][hit])
    tri = hit & (ref["obj"] < 6)                                   # triangles carry barycentrics
    np.testing.assert_allclose(g["u"][tri], ref["u"][tri], atol=1e-3)
    np.testing.assert_allclose(g["v"][tri], ref["v"][tri], atol=1e-3)
    ga, ra = r.intersect_batch(ro, rd, tm, any_hit=True), osc.intersect_batch(ro, rd, tm, any_hit=True)
    assert (ga["prim"] == ra["prim"]).mean() > 0.9995


def test_intersect_stage_parity_90k_triangles(Renderer, scene_root):
    from adapt_b200.scenes import ensure_big_meshes
    ensure_big_meshes(scene_root, ("bunny90k",))
    e, a, o, c = load_scene(scene_root, "cbox", "bunny90k.xml", 16, 16)
    assert a["primitives"].shape[0] == 89888 + 12
    r = Renderer(e, a, o, c)
    osc = _oracle(e, a, o, c, 0)                                   # XML asks for the BVH path in the oracle
    ro, rd, tm = _random_rays(60000, 3)
    g, ref = r.intersect_batch(ro, rd), osc.intersect_batch(ro, rd)
    same = g["prim"] == ref["prim"]
    assert same.mean() > 0.999
    np.testing.assert_allclose(g["t"][same], ref["t"][same], rtol=5e-5, atol=5e-6)
    ga, ra = r.intersect_batch(ro, rd, tm, any_hit=True), osc.intersect_batch(ro, rd, tm, any_hit=True)
    assert (ga["prim"] == ra["prim"]).mean() > 0.999


def test_parity_90k_triangles_render(Renderer, scene_root):
    from adapt_b200.scenes import ensure_big_meshes
    ensure_big_meshes(scene_root, ("bunny90k",))
    e, a, o, c = load_scene(scene_root, "cbox", "bunny90k.xml", 96, 54)
    r = Renderer(e, a, o, c, seed=2)
    r.render_batch(8)
    img = r.pixels.to_numpy()
    acc, cn = _oracle(e, a, o, c, 2).render(8)
    match, flipped = _flip_stats(img, acc / 8)
    assert flipped < 0.02 and rel_l2(img[match], (acc / 8)[match]) < 1e-4 and rel_l2(img, acc / 8) < TOL
    st = r.stats()
    assert abs(st["rays_closest"] - cn["rays_closest_useful"]) <= 2e-3 * cn["rays_closest_useful"]


def test_checkpoint_roundtrip_and_incremental_render(Renderer, scene_root):
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 48, 48)
    full = Renderer(e, a, o, c, seed=9)
    full.render_batch(8)
    ref = full.pixels.to_numpy()
    first = Renderer(e, a, o, c, seed=9)
    for _ in range(4):
        first.render(0, 0, 0, 0, 0, 0)                             # the driver's one-spp-per-call loop
    ck = first.get_check_point()
    assert ck["counter"] == 4 and ck["accumulation"].shape == (48, 48, 3) and first.cnt[None] == 4
    for key in ("w", "h", "crop_x", "crop_y", "crop_rx", "crop_ry", "focal", "num_objects", "num_prims", "cam_orient", "src_num", "cam_t"):
        assert key in ck
    second = Renderer(e, a, o, c, seed=9)
    second.load_check_point(ck)
    second.render_batch(4)
    assert second.cnt[None] == 8
    np.testing.assert_allclose(second.pixels.to_numpy(), ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(second.color.to_numpy(), ref * 8, rtol=1e-5, atol=1e-5)
    bad = dict(ck); bad["w"] = 47
    with pytest.raises(ValueError):
        Renderer(e, a, o, c, seed=9).load_check_point(bad)


def test_crop_window(Renderer, scene_root):
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 64, 64)
    full = Renderer(e, a, o, c, seed=1); full.render_batch(4)
    ref = full.pixels.to_numpy()
    c["film"].update(crop_x=32, crop_y=20, crop_rx=10, crop_ry=6)
    r = Renderer(e, a, o, c, seed=1); r.render_batch(4)
    assert r.do_crop and (r.start_x, r.end_x, r.start_y, r.end_y) == (22, 42, 14, 26)
    img = r.pixels.to_numpy()
    np.testing.assert_allclose(img[22:42, 14:26], ref[22:42, 14:26], rtol=1e-5, atol=1e-6)
    mask = np.ones((64, 64), bool); mask[22:42, 14:26] = False
    assert not img[mask].any()


def test_partition_invariance_on_one_gpu(Renderer, scene_root):
    """world_size 2 emulated sequentially: the two partial framebuffers sum to the single-GPU one."""
    from adapt_b200.dist import device_tensor_view, tile_partition
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 64, 48)
    one = Renderer(e, a, o, c, seed=4); one.render_batch(4)
    ref = one.color.to_numpy()
    parts = []
    for rank in range(2):
        r = Renderer(e, a, o, c, seed=4, pixel_list=tile_partition(64, 48, rank, 2, tile=16))
        r.render_batch(4); r.synchronize()
        ptr, n = r.accum_device_ptr()
        t = device_tensor_view(ptr, n, 0)                           # what the NCCL reduce operates on
        host = r.color.to_numpy()
        np.testing.assert_array_equal(t.cpu().numpy().reshape(64, 48, 3), host)
        parts.append(host)
    assert not (parts[0].any(axis=-1) & parts[1].any(axis=-1)).any()      # disjoint ownership
    np.testing.assert_allclose(parts[0] + parts[1], ref, rtol=1e-5, atol=1e-6)


def test_full_size_properties(Renderer, scene_root):
    """BASELINE config 3 at full 1920x1080: size-independent properties."""
    from adapt_b200.scenes import ensure_big_meshes
    ensure_big_meshes(scene_root, ("bunny90k",))
    e, a, o, c = load_scene(scene_root, "cbox", "bunny90k.xml")
    assert (c["film"]["width"], c["film"]["height"], c["max_bounce"]) == (1920, 1080, 16)
    r = Renderer(e, a, o, c, seed=0)
    r.render_batch(2)
    img = r.pixels.to_numpy()
    st = r.stats()
    assert img.shape == (1920, 1080, 3) and np.isfinite(img).all() and (img >= 0).all()
    assert st["paths"] == 1920 * 1080 * 2
    assert st["paths"] <= st["rays_closest"] <= st["paths"] * 16
    # a 96 x 64 window of the full-size film through the oracle, same seed: those pixels must agree
    from adapt_b200.dist import tile_partition
    win = tile_partition(1920, 1080, 0, 1, window=(900, 996, 380, 444))
    acc, _ = _oracle(e, a, o, c, 0).render(2, pixel_list=win)
    ii, jj = win // 1080, win % 1080
    got, want = img[ii, jj], acc[ii, jj] / 2
    match, flipped = _flip_stats(got[None], want[None])
    assert flipped < 0.02 and rel_l2(got[match[0]], want[match[0]]) < 1e-4 and rel_l2(got, want) < 2e-2
    # rendering more spp only refines: 2 + 2 spp equals 4 spp of a fresh renderer
    r.render_batch(2)
    r2 = Renderer(e, a, o, c, seed=0); r2.render_batch(4)
    assert rel_l2(r.pixels.to_numpy(), r2.pixels.to_numpy()) < 1e-5


def test_max_bounce_override_and_zero_bounce(Renderer, scene_root):
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 32, 32)
    r0 = Renderer(e, a, o, c, max_bounce=0); r0.render_batch(2)
    assert not r0.pixels.to_numpy().any() and r0.stats()["paths"] == 2048
    r8 = Renderer(e, a, o, c, seed=0, max_bounce=8); r8.render_batch(8)
    c8 = dict(c); c8["max_bounce"] = 8
    acc, _ = _oracle(e, a, o, c8, 0).render(8)
    assert r8.max_bounce == 8 and rel_l2(r8.pixels.to_numpy(), acc / 8) < TOL


def test_full_size_orb500k_window(Renderer, scene_root):
    """BASELINE config 4 at full 1920x1080 (501 126 triangles: glass shell, GGX shell, Fresnel-blend core, 24 bounces): a
    window of the full-size film against the oracle (its skip-pointer BVH path), plus size-independent properties."""
    from adapt_b200.dist import tile_partition
    from adapt_b200.scenes import ensure_big_meshes
    ensure_big_meshes(scene_root, ("orb500k",))
    e, a, o, c = load_scene(scene_root, "cbox", "orb500k.xml")
    assert (c["film"]["width"], c["film"]["height"], c["max_bounce"]) == (1920, 1080, 24)
    assert a["primitives"].shape[0] > 500000
    r = Renderer(e, a, o, c, seed=0)
    r.render_batch(2)
    img = r.pixels.to_numpy()
    st = r.stats()
    assert img.shape == (1920, 1080, 3) and np.isfinite(img).all() and (img >= 0).all()
    assert st["paths"] == 1920 * 1080 * 2 and st["paths"] <= st["rays_closest"] <= st["paths"] * 24
    win = tile_partition(1920, 1080, 0, 1, window=(928, 992, 460, 508))          # 64 x 48 pixels through the orb
    acc, cn = _oracle(e, a, o, c, 0).render(2, pixel_list=win)
    ii, jj = win // 1080, win % 1080
    got, want = img[ii, jj], acc[ii, jj] / 2
    match, flipped = _flip_stats(got[None], want[None])
    assert flipped < 0.05 and rel_l2(got[match[0]], want[match[0]]) < 2e-4
    np.testing.assert_allclose(got.mean(axis=0), want.mean(axis=0), rtol=0.1)
    # the same film from a renderer that only owns the window's pixels: identical samples (partition invariance)
    rw = Renderer(e, a, o, c, seed=0, pixel_list=win)
    rw.render_batch(2)
    np.testing.assert_allclose(rw.pixels.to_numpy()[ii, jj], got, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("w,h,pool", [(37, 23, 256), (1, 1, 0), (130, 3, 512)])
def test_ragged_film_and_tiny_pool(Renderer, scene_root, w, h, pool):
    """Film sizes that are not multiples of the 4 x 8 pixel patches / 32-item work groups, and a pool far smaller than the
    film (256 slots: every slot is regenerated hundreds of times, work stripes run dry and are re-probed constantly)."""
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", w, h)
    r = Renderer(e, a, o, c, seed=11, pool_size=pool)
    r.render_batch(3); r.render_batch(2)                       # two batches: the work-id space continues across calls
    img = r.pixels.to_numpy()
    st = r.stats()
    acc, cn = _oracle(e, a, o, c, 11).render(5)
    assert img.shape == (w, h, 3) and st["paths"] == cn["paths"] == w * h * 5
    match, flipped = _flip_stats(img, acc / 5)
    assert flipped <= max(0.05, 1.5 / (w * h)) and rel_l2(img[match], (acc / 5)[match]) < 1e-4
    assert abs(st["rays_closest"] - cn["rays_closest_useful"]) <= max(3, 5e-3 * cn["rays_closest_useful"])


def test_no_shadow_rays_and_no_mis(Renderer, scene_root):
    """num_shadow_ray = 0: light only arrives through emission hits (the NEE loop is empty, inv_num_shadow_ray = 1)."""
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 40, 40, num_shadow_ray=0, use_mis=False)
    r = Renderer(e, a, o, c, seed=2)
    r.render_batch(16)
    img = r.pixels.to_numpy()
    acc, cn = _oracle(e, a, o, c, 2).render(16)
    assert r.stats()["rays_shadow"] == 0 and cn["rays_shadow"] == 0
    match, flipped = _flip_stats(img, acc / 16)
    assert flipped < 0.05 and rel_l2(img[match], (acc / 16)[match]) < 1e-4 and img.max() > 0
