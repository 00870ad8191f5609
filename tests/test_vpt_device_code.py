"""Volumetric integrator -- device-code half that runs WITHOUT a GPU.  adapt_b200/csrc/pt_volume.cuh (with the shading, emitter and
traversal functions of pt_shade.cuh / pt_path.cuh / pt_trace.cuh it calls -- the very functions the `pt` kernels run) is compiled as host
C++ by tests/dev_host and driven path by path like the wavefront will drive it (trace -> vol_shade_step -> transmittance segments).
Checked against (a) the CPU oracle on the same seeded inputs, (b) the renders of the reference's own vpt code, (c) the oracle's medium
functions one by one; and the header is compiled by nvcc for sm_100a with every template instantiated.  What this cannot cover is the
launch glue (slots, queues), which does not exist yet: adapt_create rejects integrator = 1."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import load_scene, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _flip(img, ref):
    d = np.abs(img - ref).sum(-1)
    match = d <= 1e-3 * np.maximum(1.0, np.abs(ref).sum(-1))
    return match, 1.0 - float(match.mean())


CASES = [("cbox", "cbox.xml", 24, 4, {}), ("test", "media.xml", 32, 4, {}), ("test", "media-clear.xml", 32, 4, {}),
         ("test", "media.xml", 24, 3, dict(use_mis=False, use_rr=False, max_bounce=6)),
         ("test", "media-clear.xml", 24, 2, dict(num_shadow_ray=3)),
         ("csphere", "balls-mono.xml", 24, 3, {}),                     # no media at all: vpt degenerates to surface transport
         ("test", "allbxdf.xml", 24, 3, {}),                           # every BRDF / BSDF model, brdf_two_sides, five emitter kinds
         ("test", "textured.xml", 24, 3, {})]                          # albedo textures on meshes and spheres (normal / bump maps: not in vpt)


@pytest.mark.parametrize("scene,name,size,spp,kw", CASES)
def test_device_code_matches_oracle(scene_root, oracle_lib, scene, name, size, spp, kw):
    from adapt_b200._lib import pack_scene
    from dev_host import DevHostScene
    from oracle.pt_oracle import OracleScene
    e, a, o, c = load_scene(scene_root, scene, name, size, size, **kw)
    ps = pack_scene(e, a, o, c, seed=5, integrator="vpt")
    ref, cn = OracleScene(ps).render(spp)
    img, st = DevHostScene(ps).render(spp)
    assert st["paths"] == cn["paths"] == size * size * spp
    assert np.isfinite(img).all()
    match, flipped = _flip(img, ref)
    assert flipped < 0.02                                  # measured <= 0.2 % (threshold flips, as between oracle and reference)
    assert rel_l2(img[match], ref[match]) < 1e-4           # measured 2e-7 ... 7e-6
    assert rel_l2(img, ref) < 2e-3                         # north-star tolerance


@pytest.mark.parametrize("tag", ["vpt_cbox", "vpt_media", "vpt_media_clear", "vpt_media_nomis_norr"])
def test_device_code_matches_reference_render(scene_root, tag):
    """The same device functions against the renders of the reference's own renderer/vpt.py (tests/golden/reference_vpt.npz)."""
    from adapt_b200._lib import pack_scene
    from dev_host import DevHostScene
    from test_vpt_oracle import SCENES, VPT, _scene
    g = np.load(VPT)
    (e, a, o, c), spp, seed = _scene(g, scene_root, tag)
    img, _ = DevHostScene(pack_scene(e, a, o, c, seed=seed, integrator="vpt")).render(spp)
    ref = g[tag + "/color"]
    match, flipped = _flip(img, ref)
    assert flipped < 0.03 and rel_l2(img[match], ref[match]) < 1e-4 and rel_l2(img, ref) < 2e-3
    assert tag in SCENES


def _medium(kind, par=(0.0, 0.0, 0.0), pdf=(1.0, 0.0, 0.0), u_a=(0.0, 0.0, 0.0), u_s=(1.0, 1.0, 1.0)):
    from adapt_b200._lib import MEDIUM_DTYPE
    rec = np.zeros(1, dtype=MEDIUM_DTYPE)
    rec["type"], rec["ior"], rec["par"], rec["pdf"], rec["u_a"], rec["u_s"] = kind, 1.0, par, pdf, u_a, u_s
    rec["u_e"] = np.float32(u_a) + np.float32(u_s)
    return rec


@pytest.mark.parametrize("kind,par,pdf", [(0, (0.6, 0, 0), (1, 0, 0)), (0, (0.0, 0, 0), (1, 0, 0)), (1, (0.8, -0.3, 0.1), (0.5, 0.3, 0.2)),
                                          (2, (0, 0, 0), (1, 0, 0)), (-1, (0, 0, 0), (1, 0, 0))])
def test_medium_functions_equal_the_oracles(oracle_lib, kind, par, pdf):
    from dev_host import load
    from test_vpt_oracle import _phase_lib
    dev, orc = load(), _phase_lib(oracle_lib)
    m = _medium(kind, par, pdf, u_a=(0.1, 0.3, 0.0), u_s=(0.9, 0.5, 0.4))
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    n = 4000
    rng = np.random.default_rng(3)
    v = rng.normal(size=(2, n, 3)); v /= np.linalg.norm(v, axis=-1, keepdims=True)
    incid, out = v[0].astype(np.float32), v[1].astype(np.float32)
    a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
    dev.dev_host_phase_eval(m.ctypes.data, incid.ctypes.data_as(fp), out.ctypes.data_as(fp), n, a.ctypes.data_as(fp))
    orc.oracle_phase_eval(m.ctypes.data, incid.ctypes.data_as(fp), out.ctypes.data_as(fp), n, b.ctypes.data_as(fp))
    np.testing.assert_allclose(a, b, rtol=2e-6)
    one = np.float32([0.36, -0.48, 0.8])
    d1, d2 = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    dev.dev_host_phase_sample(m.ctypes.data, one.ctypes.data_as(fp), 9, n, d1.ctypes.data_as(fp), a.ctypes.data_as(fp))
    orc.oracle_phase_sample(m.ctypes.data, one.ctypes.data_as(fp), 9, n, d2.ctypes.data_as(fp), b.ctypes.data_as(fp))
    np.testing.assert_allclose(d1, d2, atol=2e-6)
    np.testing.assert_allclose(a, b, rtol=1e-5)
    if kind >= 0:
        mi1, mi2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        b1, b2 = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        dev.dev_host_medium_sample_mfp(m.ctypes.data, 1.3, 4, n, mi1.ctypes.data_as(ip), a.ctypes.data_as(fp), b1.ctypes.data_as(fp))
        orc.oracle_medium_sample_mfp(m.ctypes.data, 1.3, 4, n, mi2.ctypes.data_as(ip), b.ctypes.data_as(fp), b2.ctypes.data_as(fp))
        assert np.array_equal(mi1, mi2)
        np.testing.assert_allclose(a, b, rtol=2e-6)
        np.testing.assert_allclose(b1, b2, rtol=1e-5)


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc (cross-compiles without a GPU)")
def test_volume_header_compiles_for_sm_100a(tmp_path):
    """nvcc -gencode arch=compute_100a,code=sm_100a on a kernel that instantiates vol_shade_step / vol_transmit_step: the header is
    valid DEVICE code (no spills into an unreasonable frame, no host-only constructs)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = tmp_path / "vol_check.cu"
    src.write_text('''
#include "pt_volume.cuh"
using namespace adapt;
__global__ void k_vol_check(SceneView sv, VolumeView vv, VolPath* paths, const HitRec* hits, VolRequest* reqs, int* outcome, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    VolPath p = paths[i];
    VolRequest r[VOL_MAX_REQUESTS]; int nr = 0;
    outcome[i] = (int)vol_shade_step<M_ALL | M_TEXTURED>(sv, vv, p, hits[i], r, nr);
    for (int k = 0; k < nr; k++) {
        VolTransmit t; vol_transmit_begin(t, r[k]);
        HitRec h; unsigned a = 0, b = 0;
        do { trace<false, false>(sv, t.point, t.dir, vol_transmit_tmax(t), h, a, b); } while (vol_transmit_step(sv, vv, t, h));
        p.color += r[k].payload * t.tr;
        reqs[i * VOL_MAX_REQUESTS + k] = r[k];
    }
    paths[i] = p;
}
''')
    out = tmp_path / "vol_check.cubin"
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-cubin", "-Xptxas", "-v",
                          "-I" + os.path.join(ROOT, "adapt_b200", "csrc"), "-o", str(out), str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    assert out.exists() and "k_vol_check" in res.stderr
