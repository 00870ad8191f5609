"""Volumetric integrator (`--type vpt`, renderer/vpt.py) -- oracle half.  The CPU oracle's restatement of vpt over
homogeneous media (world fog + media attached to BSDF objects; grid volumes are not restated) is held to renders produced by the
reference's OWN code (tests/golden/make_reference_golden.py vpt -> reference_vpt.npz: the unmodified renderer/vpt.py, bxdf/medium.py,
bxdf/phase.py, sampler/phase_sampling.py executed on the Taichi stand-in with the shared counter-keyed RNG), plus closed-form checks of
the phase functions and of the free-flight sampling.  The device kernels are covered by tests/test_gpu_vpt.py (-m gpu) and, without a GPU, by
tests/test_wavefront_emulated.py; there is no CPU fallback."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_scene, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
VPT = os.path.join(HERE, "golden", "reference_vpt.npz")
SCENES = {"vpt_cbox": ("cbox", "cbox.xml"), "vpt_media": ("test", "media.xml"), "vpt_media_clear": ("test", "media-clear.xml"),
          "vpt_media_nomis_norr": ("test", "media.xml"), "vpt_media_bvh_nsr2": ("test", "media-clear.xml")}


@pytest.fixture(scope="module")
def golden():
    return np.load(VPT)


def _scene(g, scene_root, tag):
    w, h, spp, seed, mb, mis, rr, strat, bvh = (int(x) for x in g[tag + "/meta"])
    scene, name = SCENES[tag]
    e, a, o, c = load_scene(scene_root, scene, name, w, h, max_bounce=mb, use_mis=bool(mis), use_rr=bool(rr), stratified_sampling=bool(strat))
    if tag.endswith("nsr2"):
        c["num_shadow_ray"] = 2
    c["accelerator"] = "bvh" if bvh else "none"
    return (e, a, o, c), spp, seed


@pytest.mark.parametrize("tag", sorted(SCENES))
def test_oracle_vpt_matches_reference_render(golden, scene_root, oracle_lib, tag):
    from adapt_b200._lib import pack_scene
    from oracle.pt_oracle import OracleScene
    (e, a, o, c), spp, seed = _scene(golden, scene_root, tag)
    acc, cn = OracleScene(pack_scene(e, a, o, c, seed=seed, integrator="vpt")).render(spp)
    ref = golden[tag + "/color"]
    assert acc.shape == ref.shape and cn["paths"] == ref.shape[0] * ref.shape[1] * spp
    d = np.abs(acc - ref).sum(-1)
    match = d <= 1e-3 * np.maximum(1.0, np.abs(ref).sum(-1))
    assert 1.0 - match.mean() < 0.02                    # measured: 0 ... 0.7 % of pixels hold a sample that flipped at a threshold
    assert rel_l2(acc[match], ref[match]) < 1e-4        # measured: 4e-7 ... 2e-5
    assert rel_l2(acc, ref) < 2e-3                      # north-star tolerance on the whole buffer (measured <= 2.5e-5)


def test_media_reach_the_c_abi(scene_root):
    """adapt_medium records: one per object (transparent for BRDF objects) + the world medium last."""
    from adapt_b200._lib import MEDIUM_DTYPE, pack_scene
    e, a, o, c = load_scene(scene_root, "test", "media.xml", 8, 8)
    ps = pack_scene(e, a, o, c, integrator="vpt")
    md = ps.keep["media"]
    assert md.dtype == MEDIUM_DTYPE and md.shape == (len(o) + 1,) and ps.desc.integrator == 1
    kinds = {type(x.bsdf).__name__ + ":" + x.bsdf.type: k for k, x in enumerate(o)}
    glass, wax, smoke = kinds["BSDF_np:det-refraction"], kinds["BSDF_np:lambertian"], kinds["BSDF_np:null"]
    assert md[glass]["type"] == 0 and np.isclose(md[glass]["ior"], 1.45) and np.allclose(md[glass]["u_s"], [0.9, 0.6, 0.4])
    assert np.allclose(md[glass]["u_e"], md[glass]["u_a"] + md[glass]["u_s"])
    assert md[wax]["type"] == 1 and np.allclose(md[wax]["pdf"], [0.5, 0.3, 0.2]) and np.allclose(md[wax]["par"], [0.8, -0.3, 0.1])
    assert md[smoke]["type"] == 2
    assert all(md[k]["type"] == -1 for k, x in enumerate(o) if type(x.bsdf).__name__ == "BRDF_np")
    assert md[-1]["type"] == 1 and np.allclose(md[-1]["u_s"], [0.06, 0.07, 0.09])
    assert pack_scene(e, a, o, c).desc.integrator == 0
    with pytest.raises(NotImplementedError):
        pack_scene(e, a, o, c, integrator="bdpt")


def test_cuda_library_rejects_unknown_integrators_instead_of_falling_back(scene_root):
    """0 = pt and 1 = vpt have kernels; any other integrator must be refused (before looking for a GPU), never rendered as something else."""
    from adapt_b200._lib import load_library, pack_scene
    lib = load_library()
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 8, 8)
    ps = pack_scene(e, a, o, c, integrator="vpt")
    ps.desc.integrator = 2
    h = C.c_void_p()
    assert lib.adapt_create(C.byref(h), C.byref(ps.desc)) == -1 and not h
    assert b"integrator" in lib.adapt_last_error()


# ------------------------------------------------------------------------------------------------ closed-form checks of the medium code
def _phase_lib(oracle_lib):
    fp = C.POINTER(C.c_float)
    oracle_lib.oracle_phase_eval.argtypes = [C.c_void_p, fp, fp, C.c_int, fp]
    oracle_lib.oracle_phase_sample.argtypes = [C.c_void_p, fp, C.c_uint64, C.c_int, fp, fp]
    oracle_lib.oracle_medium_sample_mfp.argtypes = [C.c_void_p, C.c_float, C.c_uint64, C.c_int, C.POINTER(C.c_int32), fp, fp]
    return oracle_lib


def _medium(kind, par=(0.0, 0.0, 0.0), pdf=(1.0, 0.0, 0.0), u_a=(0.0, 0.0, 0.0), u_s=(1.0, 1.0, 1.0)):
    from adapt_b200._lib import MEDIUM_DTYPE
    rec = np.zeros(1, dtype=MEDIUM_DTYPE)
    rec["type"], rec["ior"], rec["par"], rec["pdf"], rec["u_a"], rec["u_s"] = kind, 1.0, par, pdf, u_a, u_s
    rec["u_e"] = np.float32(u_a) + np.float32(u_s)
    return rec


def _sphere_dirs(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("kind,par,pdf", [(0, (0.6, 0, 0), (1, 0, 0)), (0, (-0.4, 0, 0), (1, 0, 0)), (0, (0.0, 0, 0), (1, 0, 0)),
                                          (1, (0.8, -0.3, 0.1), (0.5, 0.3, 0.2)), (2, (0, 0, 0), (1, 0, 0))])
def test_phase_functions_integrate_to_one(oracle_lib, kind, par, pdf):
    """eval_p is a density over the sphere (bxdf/phase.py:64-79): Monte-Carlo integral over uniform directions = 1."""
    lib = _phase_lib(oracle_lib)
    m = _medium(kind, par, pdf)
    n = 400000
    out = _sphere_dirs(n, 1)
    incid = np.repeat(np.float32([[0.3, -0.5, 0.81]]) / np.linalg.norm([0.3, -0.5, 0.81]), n, 0).astype(np.float32)
    val = np.zeros(n, np.float32)
    fp = C.POINTER(C.c_float)
    lib.oracle_phase_eval(m.ctypes.data, incid.ctypes.data_as(fp), out.ctypes.data_as(fp), n, val.ctypes.data_as(fp))
    assert abs(float(val.mean()) * 4.0 * np.pi - 1.0) < 0.02


@pytest.mark.parametrize("kind,par", [(0, (0.6, 0, 0)), (0, (-0.4, 0, 0)), (2, (0, 0, 0))])
def test_phase_sampling_follows_its_density(oracle_lib, kind, par):
    """sample_p (phase.py:36-62 + phase_sampling.py): the returned value is the density of the returned direction, and the mean cosine
    of the samples is the analytic one (g for Henyey-Greenstein, 0 for Rayleigh)."""
    lib = _phase_lib(oracle_lib)
    m = _medium(kind, par)
    n = 200000
    incid = np.float32([0.0, 0.6, 0.8])
    dirs = np.zeros((n, 3), np.float32); p = np.zeros(n, np.float32)
    fp = C.POINTER(C.c_float)
    lib.oracle_phase_sample(m.ctypes.data, incid.ctypes.data_as(fp), 7, n, dirs.ctypes.data_as(fp), p.ctypes.data_as(fp))
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-4)
    # the local frame is rotated onto the incident direction (delocalize_rotate): cos(theta) is measured against it
    cos_t = dirs @ incid
    g = par[0] if kind == 0 else 0.0
    assert abs(float(cos_t.mean()) - g) < 0.01
    # the reported density equals eval_p of the sampled direction (eval_p takes ray_in pointing at the vertex: cos = -dot(in, out),
    # and the sampler's cos(theta) is taken about +incid -- the reference's own sign convention, kept as is)
    val = np.zeros(n, np.float32)
    inc_rep = np.repeat(-incid[None], n, 0).astype(np.float32)
    lib.oracle_phase_eval(m.ctypes.data, inc_rep.ctypes.data_as(fp), dirs.ctypes.data_as(fp), n, val.ctypes.data_as(fp))
    assert np.allclose(val, p, rtol=2e-3, atol=1e-6)


def test_free_flight_sampling_is_unbiased(oracle_lib):
    """Medium.sample_mfp (medium.py:88-108): E[beta * 1(surface reached)] = transmittance exp(-u_e d) per channel, and
    E[beta * 1(medium event)] = albedo-weighted probability of scattering before d, u_s / u_e * (1 - exp(-u_e d))."""
    lib = _phase_lib(oracle_lib)
    u_a, u_s = (0.1, 0.3, 0.0), (0.9, 0.5, 0.4)
    m = _medium(0, (0.2, 0, 0), u_a=u_a, u_s=u_s)
    n, d = 400000, 1.7
    is_mi = np.zeros(n, np.int32); t = np.zeros(n, np.float32); beta = np.zeros((n, 3), np.float32)
    fp = C.POINTER(C.c_float)
    lib.oracle_medium_sample_mfp(m.ctypes.data, d, 3, n, is_mi.ctypes.data_as(C.POINTER(C.c_int32)), t.ctypes.data_as(fp), beta.ctypes.data_as(fp))
    u_e = np.float64(u_a) + np.float64(u_s)
    surf = (beta * (is_mi == 0)[:, None]).mean(0)
    med = (beta * (is_mi == 1)[:, None]).mean(0)
    assert np.allclose(surf, np.exp(-u_e * d), rtol=0.02)
    assert np.allclose(med, np.float64(u_s) / u_e * (1.0 - np.exp(-u_e * d)), rtol=0.02)
    assert (t[is_mi == 0] == np.float32(d)).all() and (t[is_mi == 1] < d).all()
