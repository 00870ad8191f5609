"""Several GPUs behind ONE handle (adapt_scene_desc.n_devices > 1; Renderer(device_ids=[...])): the library replicates the scene, splits
the film into interleaved tiles and gathers it with peer-to-peer loads when it is read.  Needs at least two visible devices (skipped
otherwise; run with `gpurun --gpus 2`).  The one-process-per-GPU path (torchrun + NCCL reduce) is covered by tests/test_dist_gloo.py on
the CPU and by bench.py --gpus N."""
import numpy as np
import pytest

from conftest import load_scene, rel_l2

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.fixture(scope="module")
def Renderer():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    if torch.cuda.device_count() < 2:
        pytest.skip("one visible GPU: the multi-device handle needs two")
    from adapt_b200.build import build
    build()
    from adapt_b200.renderer.vanilla_renderer import Renderer as R
    return R


@pytest.mark.parametrize("scene,name,film,kw", [("csphere", "balls-mono.xml", (160, 96), {}), ("test", "allbxdf.xml", (96, 96), {}),
                                                ("cbox", "cbox.xml", (40, 40), dict(integrator="vpt"))])
def test_group_handle_renders_the_single_device_film(Renderer, scene_root, scene, name, film, kw):
    """Sample k of pixel p draws the same RNG stream on whatever device it runs: the gathered film equals the one-device film up to the
    summation order of the atomics."""
    e, a, o, c = load_scene(scene_root, scene, name, film[0], film[1])
    n = min(_n_devices(), 4)
    one = Renderer(e, a, o, c, seed=7, **kw)
    one.render_batch(3); one.render_batch(2)
    ref = one.pixels.to_numpy()
    grp = Renderer(e, a, o, c, seed=7, device_ids=list(range(n)), **kw)
    grp.render_batch(3); grp.render_batch(2)
    img = grp.pixels.to_numpy()
    assert rel_l2(img, ref) < 1e-6
    st, s1 = grp.stats(), one.stats()
    assert st["paths"] == s1["paths"] == film[0] * film[1] * 5 and st["rays_closest"] == s1["rays_closest"]
    np.testing.assert_allclose(grp.color.to_numpy(), one.color.to_numpy(), rtol=1e-5, atol=1e-6)
    assert grp.cnt[None] == 5


def test_group_handle_checkpoint_round_trip(Renderer, scene_root):
    """get_check_point / load_check_point through a multi-device handle: 2 spp, checkpoint, a fresh handle resumes with 3 more = 5 spp at once."""
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml", 96, 64)
    ids = [0, 1]
    a1 = Renderer(e, a, o, c, seed=3, device_ids=ids); a1.render_batch(2)
    ck = a1.get_check_point()
    assert ck["counter"] == 2 and ck["accumulation"].shape == (96, 64, 3)
    a2 = Renderer(e, a, o, c, seed=3, device_ids=ids); a2.load_check_point(ck); a2.render_batch(3)
    whole = Renderer(e, a, o, c, seed=3); whole.render_batch(5)
    assert rel_l2(a2.pixels.to_numpy(), whole.pixels.to_numpy()) < 1e-6


def test_group_handle_crop_window_and_small_films(Renderer, scene_root):
    """A crop window is split between the devices (not the whole film), and a film with fewer 32 x 32 tiles than devices gets smaller tiles."""
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 128, 96)
    c["film"].update(crop_x=60, crop_y=40, crop_rx=20, crop_ry=12)
    one = Renderer(e, a, o, c, seed=1); one.render_batch(2)
    grp = Renderer(e, a, o, c, seed=1, device_ids=[0, 1]); grp.render_batch(2)
    ref, img = one.pixels.to_numpy(), grp.pixels.to_numpy()
    assert rel_l2(img, ref) < 1e-6 and not img[:40].any() and img[40:80, 28:52].any()
    assert grp.stats()["paths"] == 40 * 24 * 2
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 24, 24)
    g2 = Renderer(e, a, o, c, seed=1, device_ids=[0, 1]); g2.render_batch(2)
    o2 = Renderer(e, a, o, c, seed=1); o2.render_batch(2)
    assert rel_l2(g2.pixels.to_numpy(), o2.pixels.to_numpy()) < 1e-6
