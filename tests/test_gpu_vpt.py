"""Volumetric integrator (`--type vpt`, renderer/vpt.py over homogeneous media) -- GPU half: k_logic_vpt / k_trace_vpt through the C ABI
against the oracle and against renders produced by the reference's own vpt code (tests/golden/reference_vpt.npz).  First run on a B200 in
session r02a (memcheck clean, 12/12 green); the same kernels also run under the SIMT emulator (tests/test_wavefront_emulated.py)."""
import os

import numpy as np
import pytest

from conftest import load_scene, rel_l2

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def Renderer():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from adapt_b200.build import build
    build()
    from adapt_b200.renderer.vanilla_renderer import Renderer as R
    return R


def _flip(img, ref):
    d = np.abs(img - ref).sum(-1)
    match = d <= 1e-3 * np.maximum(1.0, np.abs(ref).sum(-1))
    return match, 1.0 - float(match.mean())


@pytest.mark.parametrize("scene,name,size,spp,kw", [("cbox", "cbox.xml", 64, 8, {}), ("test", "media.xml", 64, 8, {}),
                                                    ("test", "media-clear.xml", 64, 8, {}),
                                                    ("test", "media.xml", 48, 4, dict(use_mis=False, use_rr=False, max_bounce=6)),
                                                    ("csphere", "balls-mono.xml", 48, 4, {}), ("test", "allbxdf.xml", 48, 4, {}),
                                                    ("test", "textured.xml", 48, 4, {})])
def test_vpt_shared_rng_parity_with_oracle(Renderer, scene_root, scene, name, size, spp, kw):
    from adapt_b200._lib import pack_scene
    from oracle.pt_oracle import OracleScene
    e, a, o, c = load_scene(scene_root, scene, name, size, size, **kw)
    r = Renderer(e, a, o, c, seed=5, integrator="vpt")
    r.render_batch(spp)
    img = r.pixels.to_numpy()
    ref, cn = OracleScene(pack_scene(e, a, o, c, seed=5, integrator="vpt")).render(spp)
    ref = ref / spp
    assert np.isfinite(img).all() and r.stats()["paths"] == cn["paths"] == size * size * spp
    match, flipped = _flip(img, ref)
    assert flipped < 0.03 and rel_l2(img[match], ref[match]) < 1e-4 and rel_l2(img, ref) < 2e-3


@pytest.mark.parametrize("tag", ["vpt_cbox", "vpt_media", "vpt_media_clear", "vpt_media_nomis_norr"])
def test_vpt_matches_reference_render(Renderer, scene_root, tag):
    from test_vpt_oracle import VPT, _scene
    g = np.load(VPT)
    (e, a, o, c), spp, seed = _scene(g, scene_root, tag)
    r = Renderer(e, a, o, c, seed=seed, integrator="vpt")
    r.render_batch(spp)
    img = r.color.to_numpy()
    ref = g[tag + "/color"]
    match, flipped = _flip(img, ref)
    assert flipped < 0.03 and rel_l2(img[match], ref[match]) < 1e-4 and rel_l2(img, ref) < 2e-3


def test_vpt_is_partition_and_pool_invariant(Renderer, scene_root):
    e, a, o, c = load_scene(scene_root, "test", "media.xml", 40, 40)
    imgs = []
    for pool in (0, 512):
        r = Renderer(e, a, o, c, seed=2, integrator="vpt", pool_size=pool)
        r.render_batch(2); r.render_batch(3)
        imgs.append(r.pixels.to_numpy())
    assert rel_l2(imgs[0], imgs[1]) < 1e-5
