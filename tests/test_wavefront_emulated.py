"""The KERNELS of libadapt_b200, run WITHOUT a GPU.  adapt_b200/csrc/pt_kernels.cuh (k_classify, k_logic, k_trace, k_closest,
k_logic_vpt) is compiled as host C++ and executed under the SIMT emulator of tests/dev_host/simt_emu.h -- every CUDA thread a fiber,
warp votes / shuffles / block barriers as real lock-step collectives -- in the launch order of adapt_abi.cu::launch_iteration, over
pool / queue / counter arrays laid out as adapt_create lays them out.  The film is compared with the oracle's on the same seeded
inputs, so slot state, regeneration, striped work claims, class lists, queue appends and the vote-scheduled traversal with lane
refill are covered by the CPU suite.  It found a real bug on its first day: the shadow queue's segments overflowed when a scene with
several material groups ran in a pool of a few hundred slots (adapt_create now sizes them for the per-group launches).
Not covered: ptxas' code generation, races, performance -- the `-m gpu` tests stay the parity tests proper."""
import numpy as np
import pytest

from conftest import load_scene, rel_l2


def _flip(img, ref):
    d = np.abs(img - ref).sum(-1)
    match = d <= 1e-3 * np.maximum(1.0, np.abs(ref).sum(-1))
    return match, 1.0 - float(match.mean())


def _run(scene_root, scene, name, w, h, spp, pool, integrator="pt", seed=5, trace_grid=2, **kw):
    from adapt_b200._lib import pack_scene
    from dev_host import wavefront_render
    from oracle.pt_oracle import OracleScene
    e, a, o, c = load_scene(scene_root, scene, name, w, h, **kw)
    ps = pack_scene(e, a, o, c, seed=seed, integrator=integrator)
    img, st = wavefront_render(ps, spp, pool, trace_grid)
    ref, cn = OracleScene(ps).render(spp)
    return img, st, ref, cn


# (scene, xml, film, spp, pool slots, allowed fraction of pixels with a flipped sample)
PT_CASES = [
    ("cbox", "cbox.xml", (16, 16), 2, 256, 0.0),                 # one material group: k_logic<M_SIMPLE>, thread t owns slot t
    ("cbox", "cbox.xml", (13, 7), 3, 256, 0.0),                  # ragged film, work ranges that end inside a 32-item group
    ("csphere", "balls-mono.xml", (16, 16), 2, 256, 0.01),       # several groups: k_classify lists + one launch per group, 4 shadow rays
    ("csphere", "balls-mono.xml", (16, 16), 2, 1024, 0.01),      # pool larger than the job: no regeneration into used slots
    ("test", "allbxdf.xml", (16, 16), 2, 512, 0.05),             # every BxDF model, brdf_two_sides, five emitter kinds
    ("test", "textured.xml", (16, 16), 2, 256, 0.02),            # albedo / normal / bump maps
]


@pytest.mark.parametrize("scene,name,film,spp,pool,max_flipped", PT_CASES)
def test_emulated_pt_wavefront_matches_oracle(scene_root, oracle_lib, scene, name, film, spp, pool, max_flipped):
    img, st, ref, cn = _run(scene_root, scene, name, film[0], film[1], spp, pool)
    assert st["paths"] == cn["paths"] == film[0] * film[1] * spp
    assert st["rays_closest"] == cn["rays_closest_useful"]          # same rays, ray for ray (the unused trace after the last bounce is skipped)
    assert st["rays_shadow"] <= cn["rays_shadow"]
    match, flipped = _flip(img, ref)
    assert flipped <= max_flipped
    assert rel_l2(img[match], ref[match]) < 2e-5


def test_pool_size_and_trace_grid_do_not_change_the_image(scene_root, oracle_lib):
    """Sample k of pixel p draws the same RNG stream whatever slot, iteration or warp it runs in."""
    a, _, _, _ = _run(scene_root, "csphere", "balls-mono.xml", 12, 12, 2, 256, trace_grid=1)
    b, _, _, _ = _run(scene_root, "csphere", "balls-mono.xml", 12, 12, 2, 768, trace_grid=3)
    assert rel_l2(a, b) < 1e-6


def test_shadow_queue_segments_hold_the_per_group_launches(scene_root, oracle_lib, monkeypatch):
    """Regression for the overflow found with this emulator: with the segment capacity of one launch (WF_OLD_SEGCAP) a 256-slot
    pool loses shadow payloads on a scene with several material groups; with adapt_create's sizing it does not."""
    img, _, ref, _ = _run(scene_root, "csphere", "balls-mono.xml", 16, 16, 2, 256)
    assert _flip(img, ref)[1] == 0.0
    monkeypatch.setenv("WF_OLD_SEGCAP", "1")
    bad, _, _, _ = _run(scene_root, "csphere", "balls-mono.xml", 16, 16, 2, 256)
    assert _flip(bad, ref)[1] > 0.02


def test_samples_split_over_calls_like_a_resumed_checkpoint(scene_root, oracle_lib):
    """cnt_origin: 2 spp, then 3 more starting at sample 3, add up to 5 spp rendered at once (stratum and RNG stream follow the sample number)."""
    from adapt_b200._lib import pack_scene
    from dev_host import wavefront_render
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 12, 12)
    ps = pack_scene(e, a, o, c, seed=9)
    first, _ = wavefront_render(ps, 2, 256)
    rest, _ = wavefront_render(ps, 3, 256, cnt_start=2)
    whole, _ = wavefront_render(ps, 5, 512)
    assert rel_l2(first + rest, whole) < 1e-6


def test_second_render_call_while_paths_are_in_flight(scene_root, oracle_lib, monkeypatch):
    """adapt_render returns when its samples are handed out; a second call raises the work limit while stragglers are still in the pool.
    Work ids are absolute and gap-free, so the film equals the one-call film."""
    from adapt_b200._lib import pack_scene
    from dev_host import wavefront_render
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 14, 11)
    ps = pack_scene(e, a, o, c, seed=4)
    one, st1 = wavefront_render(ps, 4, 256)
    monkeypatch.setenv("WF_SPLIT_SPP", "1")
    two, st2 = wavefront_render(ps, 4, 256)
    assert st1["paths"] == st2["paths"] == 14 * 11 * 4
    assert rel_l2(two, one) < 1e-6


def test_tile_partition_and_crop_window(scene_root, oracle_lib):
    """Multi-GPU tile split on one CPU: two handles that own interleaved 32x32 tiles render disjoint pixels whose sum is the whole film
    (what the NCCL reduce adds up); a crop window renders exactly its pixels."""
    from adapt_b200._lib import pack_scene
    from adapt_b200.dist import tile_partition
    from dev_host import wavefront_render
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 40, 36)
    whole, _ = wavefront_render(pack_scene(e, a, o, c, seed=3), 1, 512)
    parts = []
    for rank in range(2):
        ps = pack_scene(e, a, o, c, seed=3, pixel_list=tile_partition(40, 36, rank, 2))
        img, st = wavefront_render(ps, 1, 256)
        assert st["paths"] == ps.desc.n_pixels
        parts.append(img)
    assert not np.logical_and(np.abs(parts[0]).sum(-1) > 0, np.abs(parts[1]).sum(-1) > 0).any()
    assert rel_l2(parts[0] + parts[1], whole) < 1e-6
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 24, 24)
    c["film"].update(crop_x=10, crop_y=12, crop_rx=4, crop_ry=5)
    ps = pack_scene(e, a, o, c, seed=3)
    img, st = wavefront_render(ps, 2, 256)
    full, _ = wavefront_render(pack_scene(*load_scene(scene_root, "cbox", "cbox.xml", 24, 24), seed=3), 2, 256)
    inside = np.zeros((24, 24), bool); inside[6:14, 7:17] = True
    assert st["paths"] == inside.sum() * 2 and (img[~inside] == 0).all()
    assert rel_l2(img[inside], full[inside]) < 1e-6


VPT_CASES = [("cbox", "cbox.xml", 16, 2, 256, {}), ("test", "media.xml", 16, 2, 256, {}), ("test", "media-clear.xml", 16, 2, 512, {}),
             ("csphere", "balls-mono.xml", 12, 2, 256, {}),                                  # no media: four shadow rays per vertex
             ("test", "media.xml", 12, 2, 256, dict(num_shadow_ray=3, use_rr=False, max_bounce=5)),
             ("test", "media.xml", 12, 2, 1024, dict(use_mis=False)),
             ("test", "allbxdf.xml", 12, 2, 256, {}),                                        # two-sided BRDFs: the M_ALL | M_TEXTURED instantiation
             ("test", "textured.xml", 12, 2, 256, {})]                                       # albedo textures


@pytest.mark.parametrize("scene,name,size,spp,pool,kw", VPT_CASES)
def test_emulated_vpt_kernels_match_oracle(scene_root, oracle_lib, scene, name, size, spp, pool, kw):
    """k_logic_vpt + k_trace_vpt (shadow queue -> re-arming transmittance stream, then the closest-hit stream): the kernels that have
    not run on a GPU yet run here."""
    img, st, ref, cn = _run(scene_root, scene, name, size, size, spp, pool, integrator="vpt", **kw)
    assert st["paths"] == cn["paths"] == size * size * spp
    match, flipped = _flip(img, ref)
    assert flipped <= (0.06 if name == "allbxdf.xml" else 0.02) and rel_l2(img[match], ref[match]) < 3e-5


SWEEP_SCENES = [("cbox", "cbox-point.xml"), ("csphere", "mix-balls.xml"), ("test", "allbxdf.xml"), ("test", "media.xml")]
SWEEP_FLAGS = [dict(num_shadow_ray=0), dict(max_bounce=0), dict(max_bounce=1), dict(use_rr=False, max_bounce=5), dict(use_mis=False),
               dict(stratified_sampling=False), dict(anti_alias=False), dict(num_shadow_ray=3)]


@pytest.mark.parametrize("integrator", ["pt", "vpt"])
@pytest.mark.parametrize("scene,name", SWEEP_SCENES)
def test_emulated_kernels_over_integrator_flags(scene_root, oracle_lib, scene, name, integrator):
    """Every integrator switch the XML / command line can flip (no shadow rays, zero and one bounce, no Russian roulette, no MIS,
    uniform and no anti-aliasing, three shadow rays) on a 10x9 film in a 256-slot pool, kernels against the oracle.  (A sweep of nine
    scenes x nine flag sets x both integrators was run once when the emulator was written: 162 combinations, all within these bounds
    except the Fresnel-blend `pow` noise of balls-multi.xml, 1e-4 on the matching pixels.)"""
    for kw in SWEEP_FLAGS:
        img, st, ref, cn = _run(scene_root, scene, name, 10, 9, 2, 256, integrator=integrator, seed=7, **kw)
        assert st["paths"] == cn["paths"] == 180, kw
        if integrator == "pt" and kw.get("max_bounce", 1) > 0:          # with zero bounces the primary ray's result is unused too: not traced
            assert st["rays_closest"] == cn["rays_closest_useful"], kw
        match, flipped = _flip(img, ref)
        assert flipped <= 0.08, kw
        if match.any() and np.abs(ref[match]).sum() > 0:
            assert rel_l2(img[match], ref[match]) < 5e-5, kw


# ---------------------------------------------------------------------------------------------- compressed 8-wide BVH (ADAPT_TRACE_MODE=3)
@pytest.mark.parametrize("scene,name,film,spp,pool,max_flipped", [
    ("cbox", "cbox.xml", (16, 16), 2, 256, 0.0),
    ("csphere", "balls-mono.xml", (16, 16), 2, 256, 0.01),       # spheres and triangles in the same leaves
    ("test", "allbxdf.xml", (16, 16), 2, 512, 0.05),
    ("cbox", "bunny90k.xml", (10, 10), 1, 256, 0.02),            # 89 900 primitives: seven levels of 8-wide nodes
])
def test_emulated_wavefront_over_the_8_wide_tree(scene_root, oracle_lib, monkeypatch, scene, name, film, spp, pool, max_flipped):
    """k_trace<3>: trace_stream_cw8 over the 80-byte quantised nodes collapsed from the binary tree (bvh_build.cpp: to_gpu_layout with
    `eight`) -- same rays, ray for ray, and the same film as the oracle's brute-force / skip-pointer intersectors."""
    if name == "bunny90k.xml":
        from adapt_b200.scenes import ensure_big_meshes
        ensure_big_meshes(scene_root, ("bunny90k",))
    monkeypatch.setenv("ADAPT_TRACE_MODE", "3")
    img, st, ref, cn = _run(scene_root, scene, name, film[0], film[1], spp, pool)
    assert st["paths"] == cn["paths"] == film[0] * film[1] * spp
    assert st["rays_closest"] == cn["rays_closest_useful"]
    match, flipped = _flip(img, ref)
    assert flipped <= max_flipped
    assert rel_l2(img[match], ref[match]) < 2e-5


def test_8_wide_tree_volumetric_streams(scene_root, oracle_lib, monkeypatch):
    """k_trace_vpt<3>: the transmittance stream re-arms its lanes segment by segment over the 8-wide tree too."""
    monkeypatch.setenv("ADAPT_TRACE_MODE", "3")
    img, st, ref, cn = _run(scene_root, "test", "media.xml", 12, 12, 2, 256, integrator="vpt")
    match, flipped = _flip(img, ref)
    assert st["paths"] == cn["paths"] and flipped <= 0.02 and rel_l2(img[match], ref[match]) < 2e-5


def test_camera_rays_culled_against_the_scene_box(scene_root, oracle_lib, monkeypatch):
    """A film much wider than the scene: most camera rays miss the scene's bounding box and end in k_logic where they are generated
    (adapt_create: cull_primary).  Same film, same path and ray counts with the cull on and off."""
    kw = dict(fov=100.0)
    on, st_on, ref, cn = _run(scene_root, "cbox", "cbox.xml", 24, 12, 2, 256, **kw)
    monkeypatch.setenv("ADAPT_CULL_PRIMARY", "0")
    off, st_off, _, _ = _run(scene_root, "cbox", "cbox.xml", 24, 12, 2, 256, **kw)
    assert rel_l2(on, off) < 1e-6 and rel_l2(on, ref) < 2e-5
    assert st_on["paths"] == st_off["paths"] == cn["paths"]
    assert st_on["rays_closest"] == st_off["rays_closest"] == cn["rays_closest_useful"]     # culled rays are ray_intersect calls too
    assert st_off["rays_culled"] == 0 and st_on["rays_culled"] > 0.3 * 24 * 12 * 2          # ... that never reach the trace kernel
    assert st_on["iterations"] <= st_off["iterations"]
