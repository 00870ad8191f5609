"""Device BVH builder (adapt_b200/csrc/bvh_lbvh.h + bvh_device.cu) -- CPU half: the per-element steps the
CUDA kernels run are executed as serial loops by tests/lbvh_host and the resulting trees are checked structurally and by
tracing rays through them against brute force.  The GPU half (tests/test_gpu_lbvh.py) holds the device-built tree to the
emulated one bit for bit."""
import numpy as np
import pytest

from conftest import load_scene
from lbvh_host import build_tree, trace_check, validate


def _mesh(nu=48, nv=40):
    from adapt_b200.scenes import _sph, param_surface
    c = np.float64([2.78, 1.4, 2.8])
    V, _, F = param_surface(nu, nv, lambda T, Pm: c + _sph(T, Pm, 1.2 + 0.08 * np.sin(7 * T) * np.sin(5 * Pm)))
    return V[F].astype(np.float32).reshape(-1, 9)


def _rays(recs, n, seed):
    rng = np.random.default_rng(seed)
    cen = recs[:, :3]
    lo, hi = cen.min(0) - 1.0, cen.max(0) + 1.0
    ro = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    tgt = cen[rng.integers(0, recs.shape[0], n)] + rng.normal(0, 0.05, (n, 3))
    rd = tgt - ro
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    rd[: n // 8] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 8)] * rng.choice([-1.0, 1.0], (n // 8, 1))   # axis-aligned rays
    return ro, rd.astype(np.float32)


def _check(prims, sph=None, max_leaf=4, n_rays=1500, seed=0, **kw):
    t = build_tree(prims, sph, max_leaf=max_leaf, **kw)
    rc, depth = validate(t["nodes"], t["prims"], prims, sph)
    assert rc == 0, f"validator code {rc}"
    assert depth == t["depth"] <= 64
    ro, rd = _rays(t["prims"], n_rays, seed)
    p, tt, bp, bt, npr = trace_check(t["nodes"], t["prims"], ro, rd)
    assert np.array_equal(tt, bt)                                   # same closest distance as brute force, bit for bit
    assert ((p >= 0) == (bp >= 0)).all()
    return t, npr


@pytest.mark.parametrize("max_leaf", [1, 2, 4, 8])
def test_mesh_tree_is_sound_and_traces_like_brute_force(max_leaf):
    prims = _mesh()
    t, _ = _check(prims, max_leaf=max_leaf)
    n = prims.shape[0]
    assert t["nodes"].shape[0] <= n - 1
    if max_leaf == 1:
        assert t["nodes"].shape[0] == n - 1                         # full binary tree
    # the root box is the union of the primitive boxes
    v = prims.reshape(-1, 3, 3)
    assert np.allclose(t["root_box"][:3], v.min((0, 1)), atol=2e-4) and np.allclose(t["root_box"][3:], v.max((0, 1)), atol=2e-4)


def test_tree_quality_close_to_sah():
    """Linear BVH against the library's binned-SAH tree on the same mesh and rays: the traversal-step overhead stays moderate
    (it is what the opt-in costs at render time)."""
    prims = _mesh(96, 80)
    t_l = build_tree(prims, max_leaf=4)
    t_s = build_tree(prims, max_leaf=4, builder="sah")
    assert validate(t_s["nodes"], t_s["prims"], prims)[0] == 0
    ro, rd = _rays(t_l["prims"], 1500, 1)
    *_, n_l = trace_check(t_l["nodes"], t_l["prims"], ro, rd)
    p_s, t_sah, bp, bt, n_s = trace_check(t_s["nodes"], t_s["prims"], ro, rd)
    assert np.array_equal(t_sah, bt)
    assert n_l < 1.6 * n_s


# ---- the level-synchronous binned-SAH builder (builder 2) through the same harness ------------------------------------------------
@pytest.mark.parametrize("max_leaf", [1, 2, 4, 8])
def test_sah_device_tree_is_sound_and_traces_like_brute_force(max_leaf):
    prims = _mesh()
    t, _ = _check(prims, max_leaf=max_leaf, builder="sah_device")
    n = prims.shape[0]
    if max_leaf == 1:
        assert t["nodes"].shape[0] == n - 1
    v = prims.reshape(-1, 3, 3)
    assert np.allclose(t["root_box"][:3], v.min((0, 1)), atol=2e-4) and np.allclose(t["root_box"][3:], v.max((0, 1)), atol=2e-4)


def test_sah_device_tree_does_not_depend_on_thread_order():
    """Stable partition through the scan, atomics only for counts and min / max: whatever order the per-position steps run in, the
    same tree comes out (what makes the B200-built tree comparable bit for bit, tests/test_gpu_lbvh.py)."""
    prims = _mesh(64, 50)
    t0 = build_tree(prims, max_leaf=4, builder="sah_device")
    for seed in (1, 9):
        t1 = build_tree(prims, max_leaf=4, builder="sah_device", order_seed=seed)
        assert np.array_equal(t0["nodes"].view(np.uint32), t1["nodes"].view(np.uint32))
        assert np.array_equal(t0["prims"].view(np.uint32), t1["prims"].view(np.uint32))


def test_sah_device_tree_quality_matches_the_host_sah_tree():
    """Node visits + primitive tests per ray of the device-SAH tree against the linear BVH and the host builder's tree on the same mesh
    and rays: clearly below the linear BVH and level with the host tree (same heuristic, same bins, same centres)."""
    from lbvh_host import last_prims_tested
    prims = _mesh(96, 80)
    res = {}
    for b in ("lbvh", "sah_device", "sah"):
        t = build_tree(prims, max_leaf=4, builder=b)
        ro, rd = _rays(t["prims"], 1500, 1)
        *_, nodes_per_ray = trace_check(t["nodes"], t["prims"], ro, rd)
        res[b] = nodes_per_ray + last_prims_tested() / 1500
    assert res["sah_device"] < 0.9 * res["lbvh"]
    assert res["sah_device"] < 1.03 * res["sah"]


@pytest.mark.parametrize("n", [5, 6, 9, 17])
def test_sah_device_small_scenes(n):
    rng = np.random.default_rng(n)
    prims = (rng.random((n, 3, 3)) * 0.3 + rng.random((n, 1, 3)) * 2).astype(np.float32).reshape(-1, 9)
    for ml in (1, 2, 4):
        _check(prims, max_leaf=ml, n_rays=300, seed=n, builder="sah_device")


def test_sah_device_coincident_centres_split_on_position():
    """Thousands of primitives with the same box centre: no plane separates them, the ranges are halved by position and the tree
    stays balanced (depth ~ log2 n) and sound."""
    rng = np.random.default_rng(3)
    n = 3000
    d = rng.normal(0, 1, (n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    e = np.cross(d, rng.normal(0, 1, (n, 3))); e /= np.linalg.norm(e, axis=1, keepdims=True)
    c = np.float64([1.0, 2.0, 3.0])
    # triangles whose bounding boxes are all centred on c: vertex pairs symmetric about c would be degenerate, so use spheres
    prims = np.zeros((n, 9), np.float32)
    prims[:, :3] = c
    prims[:, 3:6] = (0.1 + rng.random((n, 1))) * np.ones((1, 3))
    sph = np.ones(n, np.uint8)
    t = build_tree(prims, sph, max_leaf=4, builder="sah_device")
    rc, depth = validate(t["nodes"], t["prims"], prims, sph)
    assert rc == 0 and depth <= 14


def test_sah_device_spheres_and_triangles_mixed(scene_root):
    e, a, o, c = load_scene(scene_root, "test", "allbxdf.xml", 16, 16)
    prims = a["primitives"].reshape(-1, 9)
    sph = np.zeros(prims.shape[0], np.uint8)
    if a["indices"] is not None:
        sph[np.asarray(a["indices"], np.int64)] = 1
    _check(prims, sph, max_leaf=4, builder="sah_device")


# ---- the compressed 8-wide tree collapsed from the device-SAH hierarchy (bvh_lbvh.h: cw8_*) -------------------------------------------
@pytest.mark.parametrize("max_leaf", [1, 2, 3])
def test_cw8_collapse_is_sound_and_traces_like_brute_force(max_leaf):
    """An independent decoder of the 80-byte nodes (dequantised child boxes, slot metadata) finds every record exactly once, every child box
    encloses its primitives, and its closest hits equal brute force bit for bit; the binary tree over the re-ordered records stays sound."""
    from lbvh_host import cw8_trace_check
    prims = _mesh()
    t = build_tree(prims, max_leaf=max_leaf, builder="sah_device", eight=True)
    assert validate(t["nodes"], t["prims"], prims)[0] == 0
    assert 1 <= t["depth8"] <= 12 and t["nodes8"].shape[0] < t["nodes"].shape[0]
    ro, rd = _rays(t["prims"], 1500, 2)
    _, tt, _, bt, _ = trace_check(t["nodes"], t["prims"], ro, rd)
    rc, t8, p8 = cw8_trace_check(t["nodes8"], t["prims"], prims, None, ro, rd)
    assert rc == 0
    assert np.array_equal(tt, bt) and np.array_equal(t8, bt)


def test_cw8_collapse_spheres_triangles_and_tiny_ranges(scene_root):
    from lbvh_host import cw8_trace_check
    e, a, o, c = load_scene(scene_root, "test", "allbxdf.xml", 16, 16)
    prims = a["primitives"].reshape(-1, 9)
    sph = np.zeros(prims.shape[0], np.uint8)
    if a["indices"] is not None:
        sph[np.asarray(a["indices"], np.int64)] = 1
    for sub in (prims.shape[0], 9, 5, 4):                       # the whole scene, and scenes of a few primitives (root with leaf children only)
        t = build_tree(prims[:sub], sph[:sub], max_leaf=3, builder="sah_device", eight=True)
        ro, rd = _rays(t["prims"], 400, sub)
        _, tt, _, bt, _ = trace_check(t["nodes"], t["prims"], ro, rd)
        rc, t8, _ = cw8_trace_check(t["nodes8"], t["prims"], prims[:sub], sph[:sub], ro, rd)
        assert rc == 0 and np.array_equal(tt, bt) and np.array_equal(t8, bt)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9])
def test_tiny_scenes(n):
    rng = np.random.default_rng(n)
    prims = rng.random((n, 9)).astype(np.float32)
    _check(prims, max_leaf=4, n_rays=200)


def test_spheres_and_triangles_mixed(scene_root):
    _, a, objs, _ = load_scene(scene_root, "csphere", "balls-mono.xml")
    prims = a["primitives"].reshape(-1, 9)
    sph = np.zeros(prims.shape[0], np.uint8)
    if a["indices"] is not None:
        sph[np.asarray(a["indices"], np.int64)] = 1
    assert sph.any()
    _check(prims, sph, max_leaf=4)
    _check(prims, sph, max_leaf=1)


def test_axis_aligned_flat_triangles(scene_root):
    _, a, _, _ = load_scene(scene_root, "cbox", "cbox.xml")        # walls: boxes of zero thickness get the 1e-4 pad
    _check(a["primitives"].reshape(-1, 9), max_leaf=4)
    _check(a["primitives"].reshape(-1, 9), max_leaf=1)


def test_coincident_centres_split_on_position():
    """Identical Morton keys (all primitives share one centre, or sit on a plane): the position tie-break keeps the radix
    tree balanced instead of degenerating into a chain deeper than the traversal stack."""
    rng = np.random.default_rng(5)
    one = rng.random((1, 9)).astype(np.float32)
    prims = np.repeat(one, 3000, 0)
    t, _ = _check(prims, max_leaf=4, n_rays=100)
    assert t["depth"] <= 14
    # two clusters of duplicates + distinct primitives
    prims = np.concatenate([np.repeat(one, 500, 0), np.repeat(one + 3.0, 700, 0), rng.random((300, 9)).astype(np.float32) * 4])
    _check(prims, max_leaf=4, n_rays=300)


def test_records_carry_object_and_class():
    prims = _mesh(12, 10)
    n = prims.shape[0]
    prim_obj = (np.arange(n) % 3).astype(np.int32)
    cls = np.uint8([1, 5, 8])
    t = build_tree(prims, None, prim_obj, cls, max_leaf=4)
    pid = t["prims"][:, 9].view(np.uint32)
    ob = t["prims"][:, 10].view(np.uint32)
    cl = t["prims"][:, 11].view(np.uint32)
    assert sorted(pid.tolist()) == list(range(n))
    assert np.array_equal(ob, prim_obj[pid].astype(np.uint32))
    assert np.array_equal(cl, cls[prim_obj[pid]].astype(np.uint32))


@pytest.mark.parametrize("builder,seed", [("lbvh", 0), ("lbvh", 7), ("sah", 0), ("sah", 3)])
def test_refit_keeps_the_topology_and_encloses_the_deformed_mesh(builder, seed):
    """adapt_refit_geometry's per-element steps (bvh_lbvh.h: refit_*) as serial loops: after a non-rigid deformation the tree has the same
    child codes, every box encloses what is below it (validate), and closest hits equal brute force bit for bit -- whatever order the
    nodes' "threads" run in (the arrival counters decide who carries a node's box upwards)."""
    from lbvh_host import refit_tree
    prims = _mesh(40, 32)
    t = build_tree(prims, max_leaf=4, builder=builder)
    codes = t["nodes"].view(np.uint32)[:, 12:14].copy()
    v = prims.reshape(-1, 3, 3).copy()
    c = v.reshape(-1, 3).mean(0)
    v = ((v - c) * np.float32([1.15, 0.85, 1.05]) + c + np.float32([0.3, -0.2, 0.1]) + 0.03 * np.sin(9 * v[..., [1, 2, 0]])).astype(np.float32)
    moved = v.reshape(-1, 9)
    refit_tree(t, moved, order_seed=seed)
    np.testing.assert_array_equal(t["nodes"].view(np.uint32)[:, 12:14], codes)
    rc, _ = validate(t["nodes"], t["prims"], moved)
    assert rc == 0, f"validator code {rc}"
    ro, rd = _rays(t["prims"], 1200, 5)
    p, tt, bp, bt, _ = trace_check(t["nodes"], t["prims"], ro, rd)
    assert np.array_equal(tt, bt) and ((p >= 0) == (bp >= 0)).all() and (p >= 0).sum() > 300
    # refitting back to the rest pose gives the records of a fresh build bit for bit, and its boxes up to the ulps the refit adds per
    # level (a node's box is the union of child boxes that were already widened by one ulp when they were stored): never smaller
    refit_tree(t, prims, order_seed=seed + 1)
    fresh = build_tree(prims, max_leaf=4, builder=builder)
    np.testing.assert_array_equal(t["prims"].view(np.uint32), fresh["prims"].view(np.uint32))
    if builder == "lbvh":
        lo_cols, hi_cols = [0, 2, 4, 6, 8, 10], [1, 3, 5, 7, 9, 11]
        assert (t["nodes"][:, lo_cols] <= fresh["nodes"][:, lo_cols]).all() and (t["nodes"][:, hi_cols] >= fresh["nodes"][:, hi_cols]).all()
        np.testing.assert_allclose(t["nodes"][:, :12], fresh["nodes"][:, :12], rtol=1e-5, atol=1e-6)


def test_refit_of_a_single_leaf_scene():
    from lbvh_host import refit_tree
    prims = _mesh(40, 32)[:3]
    t = build_tree(prims, max_leaf=4)
    moved = (prims.reshape(-1, 3, 3) + np.float32([1.0, 2.0, -1.0])).reshape(-1, 9)
    refit_tree(t, moved)
    assert validate(t["nodes"], t["prims"], moved)[0] == 0
