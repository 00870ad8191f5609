"""Scene fixtures: writes AdaPT XML v1.1 scenes + Wavefront OBJ meshes that the (re-hosted) parser loads.

Two groups:
  * the Cornell fixtures the reference ships and BASELINE.json's configs 1-2 are quoted on
    (``cbox/cbox.xml``, ``csphere/balls-mono.xml``) plus ``test/allbxdf.xml`` which exercises every
    BxDF / emitter branch of the hot path.  The mesh numbers are the classic Cornell-box measurements
    (in units of 100 mm) with the same vertex/face order as the reference's ``scenes/meshes/cornell``
    so that primitive ids -- and with them the emitter-triangle pick ``rand % mesh_num`` -- agree.
  * deterministic analytic stand-ins for the big meshes the reference does not ship
    (SURVEY 8(d)): ``bunny90k`` (89 888 tris), ``orb500k`` (501 126 tris), ``car290k`` (290 322 tris).
    No RNG is involved, so every machine generates byte-identical files.

    python -m adapt_b200.scenes [--root scenes] [--big]
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np

# ------------------------------------------------------------------------------------------------
# OBJ writer
# ------------------------------------------------------------------------------------------------


def write_obj(path: str, name: str, verts: np.ndarray, faces: np.ndarray, normals: Optional[np.ndarray] = None,
              face_normals: Optional[np.ndarray] = None, uvs: Optional[np.ndarray] = None):
    """faces: (F,3) 0-based vertex ids. normals: per-vertex (same indexing) or, with face_normals
    (F,3) 0-based ids into `normals`, an explicit vn table. uvs: per-vertex."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    lines = [f"o {name}"]
    lines += ["v %.6f %.6f %.6f" % tuple(v) for v in verts]
    if normals is not None:
        lines += ["vn %.4f %.4f %.4f" % tuple(n) for n in normals]
    if uvs is not None:
        lines += ["vt %.6f %.6f" % tuple(t) for t in uvs]
    f1 = faces + 1
    if normals is None:
        lines += ["f %d %d %d" % tuple(f) for f in f1]
    else:
        fn = (faces if face_normals is None else face_normals) + 1
        if uvs is None:
            lines += ["f %d//%d %d//%d %d//%d" % (a, na, b, nb, c, nc) for (a, b, c), (na, nb, nc) in zip(f1, fn)]
        else:
            lines += ["f %d/%d/%d %d/%d/%d %d/%d/%d" % (a, a, na, b, b, nb, c, c, nc) for (a, b, c), (na, nb, nc) in zip(f1, fn)]
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


# ------------------------------------------------------------------------------------------------
# Cornell box data (vertex order and triangulation as in the reference's meshes)
# ------------------------------------------------------------------------------------------------
_Q = lambda a, b: np.int64([a, b])      # noqa: E731  two triangles as 0-based index triples

CORNELL: Dict[str, dict] = {
    "cbox_floor": dict(
        v=[(5.528, 0, 0), (0, 0, 0), (0, 0, 5.592), (5.496, 0, 5.592)], vn=[(0, 1, 0)],
        vt=[(1.0, 0.829052), (0.0, 0.829052), (0.0, 0.154667), (0.994211, 0.154667)],
        f=[(1, 3, 0), (1, 2, 3)], fn=[(0, 0, 0), (0, 0, 0)]),
    "cbox_ceiling": dict(
        v=[(5.56, 5.488, 0), (5.56, 5.487999, 5.592), (0, 5.487999, 5.592), (0, 5.488, 0)], vn=[(0, -1, 0), (0, 1, 0)],
        f=[(1, 3, 0), (1, 2, 3)], fn=[(0, 0, 0), (0, 0, 0)]),
    "cbox_back": dict(
        v=[(5.496, 0, 5.592), (0, 0, 5.592), (0, 5.487999, 5.592), (5.56, 5.487999, 5.592)], vn=[(0, 0, -1)],
        f=[(0, 2, 3), (0, 1, 2)], fn=[(0, 0, 0), (0, 0, 0)]),
    "cbox_greenwall": dict(
        v=[(0, 0, 5.592), (0, 0, 0), (0, 5.488, 0), (0, 5.487999, 5.592)], vn=[(1, 0, 0)],
        f=[(1, 3, 0), (1, 2, 3)], fn=[(0, 0, 0), (0, 0, 0)]),
    "cbox_redwall": dict(
        v=[(5.528, 0, 0), (5.496, 0, 5.592), (5.56, 5.487999, 5.592), (5.56, 5.488, 0)],
        vn=[(-1.0, 0.0058, 0.0), (-0.9999, 0.0117, -0.0057)],
        f=[(0, 2, 3), (0, 1, 2)], fn=[(0, 0, 0), (1, 1, 1)]),
    "cbox_luminaire": dict(
        v=[(3.43, 5.488, 2.27), (3.43, 5.488, 3.32), (2.13, 5.488, 3.32), (2.13, 5.488, 2.27)], vn=[(0, -1, 0)],
        f=[(1, 3, 0), (1, 2, 3)], fn=[(0, 0, 0), (0, 0, 0)]),
}


def _box_mesh(top: Sequence[Sequence[float]], side_order: Sequence[Sequence[int]], side_normals, height: float):
    """Cornell blocks: a top quad, four side quads and a bottom quad, 24 vertices, (a,b,c),(a,c,d) fans."""
    t = [tuple(p) for p in top]
    b = [(p[0], 0.0, p[2]) for p in t]
    verts: List[tuple] = list(t)
    for (i0, i1) in side_order:
        verts += [b[i0], t[i0], t[i1], b[i1]]
    verts += [b[3], b[2], b[1], b[0]]
    faces, fn = [], []
    for q in range(6):
        a = 4 * q
        faces += [(a, a + 1, a + 2), (a, a + 2, a + 3)]
        fn += [(q, q, q), (q, q, q)]
    vn = [(0, 1, 0)] + list(side_normals) + [(0, -1, 0)]
    return dict(v=verts, vn=vn, f=faces, fn=fn)


CORNELL["cbox_smallbox"] = _box_mesh(
    [(1.3, 1.65, 0.65), (0.82, 1.65, 2.25), (2.4, 1.65, 2.72), (2.9, 1.65, 1.14)],
    [(3, 2), (0, 3), (1, 0), (2, 1)],
    [(0.9534, 0, 0.3017), (0.2928, 0, -0.9562), (-0.9578, 0, -0.2873), (-0.2851, 0, 0.9585)], 1.65)
CORNELL["cbox_largebox"] = _box_mesh(
    [(4.23, 3.3, 2.47), (2.65, 3.3, 2.96), (3.14, 3.3, 4.56), (4.72, 3.3, 4.06)],
    [(0, 3), (3, 2), (2, 1), (1, 0)],
    [(0.9556, 0, -0.2945), (0.3017, 0, 0.9534), (-0.9562, 0, 0.2928), (-0.2962, 0, -0.9551)], 3.3)


def write_cornell_meshes(mesh_dir: str):
    for name, m in CORNELL.items():
        write_obj(os.path.join(mesh_dir, name + ".obj"), name, np.float64(m["v"]), np.int64(m["f"]),
                  normals=np.float64(m["vn"]), face_normals=np.int64(m["fn"]),
                  uvs=np.float64(m["vt"]) if "vt" in m else None)


# ------------------------------------------------------------------------------------------------
# analytic big meshes
# ------------------------------------------------------------------------------------------------
def param_surface(nu: int, nv: int, fn, closed_u: bool = True):
    """Triangulate r(theta, phi) on an nu x nv quad grid -> verts, per-vertex normals, faces (2*nu*nv)."""
    th = np.linspace(0.02, np.pi - 0.02, nv + 1)                  # open at the poles: no degenerate triangles
    ph = np.linspace(0.0, 2.0 * np.pi, nu + 1)
    T, Pm = np.meshgrid(th, ph, indexing="ij")                    # (nv+1, nu+1)
    P = fn(T, Pm)                                                 # (nv+1, nu+1, 3)
    eps = 1e-4
    dT = (fn(T + eps, Pm) - fn(T - eps, Pm)) / (2 * eps)
    dP = (fn(T, Pm + eps) - fn(T, Pm - eps)) / (2 * eps)
    N = np.cross(dP, dT)
    N /= np.linalg.norm(N, axis=-1, keepdims=True)
    idx = np.arange((nv + 1) * (nu + 1)).reshape(nv + 1, nu + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)], 0)
    # interleave the two triangles of each quad so neighbouring primitives are spatially close
    faces = faces.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    return P.reshape(-1, 3), N.reshape(-1, 3), faces


def _sph(T, Pm, r):
    return np.stack([r * np.sin(T) * np.cos(Pm), r * np.cos(T), r * np.sin(T) * np.sin(Pm)], -1)


def bunny90k_mesh():
    c = np.float64([2.78, 1.4, 2.8])
    return param_surface(212, 212, lambda T, Pm: c + _sph(T, Pm, 1.2 + 0.08 * np.sin(7 * T) * np.sin(5 * Pm)))


def orb_shell_mesh(radius: float, bump: float):
    c = np.float64([2.78, 1.5, 2.8])
    return param_surface(289, 289, lambda T, Pm: c + _sph(T, Pm, radius + bump * np.sin(9 * T) * np.sin(6 * Pm)))


def car290k_mesh():
    c = np.float64([2.78, 0.95, 2.8])

    def fn(T, Pm):
        e1, e2 = 0.5, 0.6                                          # superellipsoid exponents
        sp = lambda x, e: np.sign(x) * np.abs(x) ** e             # noqa: E731
        x = 2.0 * sp(np.sin(T), e1) * sp(np.cos(Pm), e2)
        y = 0.8 * sp(np.cos(T), e1)
        z = 1.0 * sp(np.sin(T), e1) * sp(np.sin(Pm), e2)
        return c + np.stack([x, y, z], -1)
    return param_surface(381, 381, fn)


# ------------------------------------------------------------------------------------------------
# XML writer
# ------------------------------------------------------------------------------------------------
_CAM = dict(target="2.78, 2.73, -7.99", origin="2.78, 2.73, -8.00", up="0, 1, 0")


def _sensor(width, height, max_bounce, nsr, accelerator=None, extra=None):
    s = ['<sensor type="perspective">', '<float name="fov" value="39.3077"/>',
         f'<integer name="max_bounce" value="{max_bounce}"/>', f'<integer name="num_shadow_ray" value="{nsr}"/>',
         '<boolean name="use_rr" value="true"/>', '<boolean name="anti_alias" value="true"/>',
         '<boolean name="stratified_sampling" value="true"/>', '<boolean name="use_mis" value="true"/>']
    if accelerator:
        s.append(f'<string name="accelerator" value="{accelerator}"/>')
    for line in (extra or []):
        s.append(line)
    s += ['<transform name="toWorld">',
          f'<lookat target="{_CAM["target"]}" origin="{_CAM["origin"]}" up="{_CAM["up"]}"/>', '</transform>',
          '<film type="film">', f'<integer name="width" value="{width}"/>', f'<integer name="height" value="{height}"/>',
          '</film>', '</sensor>']
    return s


def _brdf(kind, _id, **rgb):
    out = [f'<brdf type="{kind}" id="{_id}">']
    for k, v in rgb.items():
        out.append(f'<rgb name="{k}" {v}/>' if "=" in v else f'<rgb name="{k}" value="{v}"/>')
    return out + ["</brdf>"]


def _bsdf(kind, _id, k_d, ior):
    return [f'<bsdf type="{kind}" id="{_id}">', f'<rgb name="k_d" value="{k_d}"/>', '<medium type="transparent">',
            f'<float name="ior" value="{ior}"/>', "</medium>", "</bsdf>"]


def _obj(path, material, emitter=None, translate=None, euler=None):
    out = ['<shape type="obj">', f'<string name="filename" value="{path}"/>']
    if translate is not None:
        out += ['<transform name="toWorld">', '<translate x="%g" y="%g" z="%g"/>' % tuple(translate), "</transform>"]
    if euler is not None:
        out += ['<transform name="toWorld">', '<rotate type="euler" r="%g" p="%g" y="%g"/>' % tuple(euler), "</transform>"]
    out.append(f'<ref type="material" id="{material}"/>')
    if emitter:
        out.append(f'<ref type="emitter" id="{emitter}"/>')
    return out + ["</shape>"]


def _sphere(center, radius, material, emitter=None):
    out = ['<shape type="sphere">', '<point name="center" x="%g" y="%g" z="%g"/>' % tuple(center),
           f'<float name="radius" value="{radius}"/>', f'<ref type="material" id="{material}"/>']
    if emitter:
        out.append(f'<ref type="emitter" id="{emitter}"/>')
    return out + ["</shape>"]


_WORLD = ['<world name="free-space">', '<rgb name="skybox" value="0.0"/>', '<rgb name="ambient" value="0.0"/>',
          '<medium type="transparent">', '<float name="ior" value="1.0"/>', "</medium>", "</world>"]
_AREA = ['<emitter type="area" id="area">', '<rgb name="emission" value="70.0, 63.2, 60.3"/>',
         '<rgb name="scaler" value="0.6"/>', "</emitter>"]
_M = "../meshes/cornell/"


def _write_xml(path, blocks):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    lines = ["<?xml version='1.0' encoding='utf-8'?>", '<scene version="1.1">']
    for b in blocks:
        lines += b
    lines.append("</scene>")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def _room(wall="phong", with_light=True):
    """Cornell room shapes: luminaire (area light) + floor/ceiling/back/green/red, reference object order."""
    shapes = []
    if with_light:
        shapes += _obj(_M + "cbox_luminaire.obj", "light", emitter="area", translate=(0, -0.001, 0))
    shapes += _obj(_M + "cbox_floor.obj", "diffuse", translate=(0, 0, 0))
    shapes += _obj(_M + "cbox_ceiling.obj", "diffuse") + _obj(_M + "cbox_back.obj", "diffuse")
    shapes += _obj(_M + "cbox_greenwall.obj", "right_wall") + _obj(_M + "cbox_redwall.obj", "left_wall")
    mats = _brdf(wall, "diffuse", k_d="#D2D2D2", k_g="1.0", k_s="0.0") + _brdf(wall, "left_wall", k_d="#DD2525", k_g="1.0", k_s="0.0") \
        + _brdf(wall, "right_wall", k_d="#25DD25", k_g="1.0", k_s="0.0") + _brdf(wall, "light", k_d="#CCCCCC", k_g="1.0", k_s="0.0")
    return mats, shapes


def write_cbox(root):
    """scenes/cbox/cbox.xml of the reference: 7 objects / 34 triangles, one point light, all Lambertian."""
    mats = _brdf("lambertian", "box", k_d="#BCBCBC", k_g="1.0", k_s="0.0") + _brdf("lambertian", "white", k_d="#BDBDBD", k_g="1.0", k_s="0.0") \
        + _brdf("lambertian", "left_wall", k_d="#DD2525", k_g="1.0", k_s="0.0") + _brdf("lambertian", "right_wall", k_d="#25DD25", k_g="1.0", k_s="0.0")
    light = ['<emitter type="point" id="point">', '<rgb name="emission" value="60.0, 60.0, 60.0"/>', '<rgb name="scaler" value="0.2"/>',
             '<point name="center" x="2.779" y="4.5" z="3"/>', "</emitter>"]
    shapes = _obj(_M + "cbox_floor.obj", "white", translate=(0, 0, 0)) + _obj(_M + "cbox_ceiling.obj", "white") \
        + _obj(_M + "cbox_back.obj", "white") + _obj(_M + "cbox_greenwall.obj", "right_wall") + _obj(_M + "cbox_redwall.obj", "left_wall") \
        + _obj(_M + "cbox_smallbox.obj", "box", euler=(0, 0, 0)) + _obj(_M + "cbox_largebox.obj", "box", euler=(0, 0, 0))
    world = ['<world name="free-space">', '<rgb name="skybox" value="0.0"/>', '<rgb name="ambient" value="0.0"/>', '<medium type="hg">',
             '<rgb name="u_a" value="0.0"/>', '<rgb name="u_s" value="0.2"/>', '<rgb name="par" value="0.9"/>',
             '<float name="ior" value="1.0"/>', "</medium>", "</world>"]
    _write_xml(os.path.join(root, "cbox", "cbox.xml"), [_sensor(512, 512, 12, 1), mats, light, shapes, world])


def write_balls_mono(root):
    """scenes/csphere/balls-mono.xml of the reference: 12 triangles + 6 spheres, area light, six BxDF kinds, nsr 4."""
    mats, room = _room("phong")
    mats += _brdf("fresnel-blend", "fresnel", k_d="#CACACA", k_s="#333333", k_g='r="10" g="1000"')
    mats += _brdf("mod-phong", "glossy", k_d="#BCBCBC", k_g="10.0", k_s="#424242")
    mats += _brdf("lambertian", "white", k_d="#FFFFFF", k_g="1.0", k_s="0.0")
    mats += _brdf("specular", "mirror", k_d="#DEDEDE", k_g="1.0", k_s="0.0")
    mats += _bsdf("det-refraction", "glass", "#FAFAFA", 1.5)
    balls = _sphere((4.5, 0.6, 1.1), 0.6, "mirror") + _sphere((4.2, 0.5, 4.1), 0.5, "fresnel") + _sphere((3.2, 0.4, 0.8), 0.4, "white") \
        + _sphere((2.7, 0.4, 3.8), 0.4, "white") + _sphere((0.9, 0.5, 0.6), 0.5, "glossy") + _sphere((1.7, 1.2, 1.9), 1.2, "glass")
    _write_xml(os.path.join(root, "csphere", "balls-mono.xml"), [_sensor(512, 512, 16, 4), mats, _AREA, room + balls, _WORLD])


def write_allbxdf(root):
    """Coverage scene: every BRDF/BSDF type and every emitter type of the hot path, two-sided BRDFs,
    a sphere area light next to a mesh area light (so sample_light takes its two-draw branch)."""
    mats, room = _room("lambertian")
    mats += _brdf("oren-nayar", "pbr-diffuse", k_d="#18455c", sigma='r="35.0"', k_s="0.0")
    mats += _brdf("thin-coat", "plastic", k_d="#18455c", sigma='r="35.0" b="1.9"', k_s="#FFFFFF")
    mats += _brdf("microfacet", "ggx", k_d="#E0C080", roughness="0.3", ref_ior="1.0, 1.5, 0.0")
    mats += _brdf("fresnel-blend", "fresnel", k_d="#CACACA", k_s="#333333", k_g='r="10" g="1000"')
    mats += _brdf("mod-phong", "glossy", k_d="#BCBCBC", k_g="10.0", k_s="#424242")
    mats += _brdf("phong", "phong", k_d="#909090", k_g="20.0", k_s="#404040")
    mats += _brdf("specular", "mirror", k_d="#DEDEDE", k_g="1.0", k_s="0.0")
    mats += _bsdf("det-refraction", "glass", "#FAFAFA", 1.5) + _bsdf("lambertian", "frosted", "#E0E0FF", 1.3)
    lights = _AREA + ['<emitter type="area" id="ball-light">', '<rgb name="emission" value="8.0, 9.0, 12.0"/>', "</emitter>",
                      '<emitter type="point" id="pt">', '<rgb name="emission" value="3.0, 3.0, 2.0"/>', '<point name="center" x="1.0" y="4.0" z="1.0"/>', "</emitter>",
                      '<emitter type="spot" id="spot">', '<rgb name="emission" value="20.0, 12.0, 12.0"/>',
                      '<point name="pos" x="4.6" y="5.0" z="1.0"/>', '<point name="dir" x="-0.3" y="-1.0" z="0.4"/>',
                      '<float name="half-angle" value="25.0"/>', "</emitter>",
                      '<emitter type="collimated" id="beam">', '<rgb name="emission" value="30.0, 30.0, 10.0"/>',
                      '<point name="pos" x="2.78" y="5.2" z="4.6"/>', '<point name="dir" x="0.0" y="-1.0" z="-0.1"/>',
                      '<float name="radius" value="0.4"/>', "</emitter>"]
    balls = _sphere((4.5, 0.6, 1.1), 0.6, "mirror") + _sphere((4.2, 0.5, 4.1), 0.5, "fresnel") + _sphere((3.2, 0.4, 0.8), 0.4, "pbr-diffuse") \
        + _sphere((2.7, 0.4, 3.8), 0.4, "plastic") + _sphere((0.9, 0.5, 0.6), 0.5, "glossy") + _sphere((1.7, 1.2, 1.9), 1.2, "glass") \
        + _sphere((4.6, 1.9, 2.6), 0.45, "ggx") + _sphere((0.9, 2.6, 3.6), 0.5, "frosted") + _sphere((3.4, 3.6, 4.2), 0.35, "phong") \
        + _sphere((1.2, 4.4, 4.4), 0.3, "light", emitter="ball-light")
    # a small smooth-shaded mesh (interpolated, un-normalised vertex normals) hanging in the room
    c = np.float64([2.3, 3.3, 2.2])
    bv, bn, bf = param_surface(16, 12, lambda T, Pm: c + _sph(T, Pm, 0.55 + 0.08 * np.sin(3 * T) * np.sin(2 * Pm)))
    write_obj(os.path.join(root, "meshes", "test", "blob.obj"), "blob", bv, bf, normals=bn)
    boxes = _obj("../meshes/test/blob.obj", "phong")
    sensor = _sensor(256, 256, 10, 2, extra=['<boolean name="brdf_two_sides" value="true"/>', '<boolean name="enable_microfacet" value="true"/>'])
    _write_xml(os.path.join(root, "test", "allbxdf.xml"), [sensor, mats, lights, room + balls + boxes, _WORLD])


def _medium(kind, u_a, u_s, par, ior, pdf=None):
    out = [f'<medium type="{kind}">', f'<rgb name="u_a" value="{u_a}"/>', f'<rgb name="u_s" value="{u_s}"/>', f'<rgb name="par" value="{par}"/>']
    if pdf is not None:
        out.append(f'<rgb name="pdf" value="{pdf}"/>')
    return out + [f'<float name="ior" value="{ior}"/>', "</medium>"]


def write_media(root):
    """Coverage scenes of the volumetric integrator (renderer/vpt.py, homogeneous media): a glass ball filled with an HG medium, a
    null-surface box of Rayleigh 'smoke' that light and shadow rays pass through, a frosted (Lambertian-transmission) ball with a
    multi-HG medium, Lambertian walls, an area light and a point light (sample_light's two-draw branch).  media.xml fills the
    free space with a thin multi-HG medium, media-clear.xml leaves it transparent."""
    mats, room = _room("lambertian")
    mats += ['<bsdf type="det-refraction" id="milk-glass">', '<rgb name="k_d" value="#FAFAFA"/>'] + _medium("hg", "0.05, 0.1, 0.2", "0.9, 0.6, 0.4", "0.35", 1.45) + ["</bsdf>"]
    mats += ['<bsdf type="null" id="smoke">', '<rgb name="k_d" value="#FFFFFF"/>'] + _medium("rayleigh", "0.1", "0.7, 0.8, 1.1", "0.0", 1.0) + ["</bsdf>"]
    mats += ['<bsdf type="lambertian" id="wax">', '<rgb name="k_d" value="#E0D0B0"/>'] \
        + _medium("multi-hg", "0.02", "1.5, 1.2, 0.9", "0.8, -0.3, 0.1", 1.3, pdf="0.5, 0.3, 0.2") + ["</bsdf>"]
    mats += _brdf("lambertian", "box", k_d="#BCBCBC", k_g="1.0", k_s="0.0")
    lights = _AREA + ['<emitter type="point" id="pt">', '<rgb name="emission" value="6.0, 6.0, 5.0"/>', '<point name="center" x="1.0" y="4.2" z="1.2"/>', "</emitter>"]
    shapes = room + _sphere((1.6, 1.1, 2.0), 1.0, "milk-glass") + _sphere((4.2, 0.7, 1.4), 0.7, "wax") \
        + _obj(_M + "cbox_smallbox.obj", "smoke", translate=(2.2, 1.8, 1.6)) + _obj(_M + "cbox_largebox.obj", "box", euler=(0, 0, 0))
    sensor = _sensor(128, 128, 10, 1)
    fog = ['<world name="fog">', '<rgb name="skybox" value="0.0"/>', '<rgb name="ambient" value="0.0"/>'] \
        + _medium("multi-hg", "0.01", "0.06, 0.07, 0.09", "0.7, 0.2, -0.4", 1.0, pdf="0.6, 0.3, 0.1") + ["</world>"]
    _write_xml(os.path.join(root, "test", "media.xml"), [sensor, mats, lights, shapes, fog])
    _write_xml(os.path.join(root, "test", "media-clear.xml"), [sensor, mats, lights, shapes, _WORLD])


def write_big_xml(root):
    """XML of BASELINE configs 3-5; the OBJ files come from ensure_big_meshes()."""
    S = "../meshes/synth/"
    mats, room = _room("lambertian")
    _write_xml(os.path.join(root, "cbox", "bunny90k.xml"),
               [_sensor(1920, 1080, 16, 1, accelerator="bvh"), mats + _brdf("lambertian", "body", k_d="#BCBCBC"), _AREA,
                room + _obj(S + "bunny90k.obj", "body"), _WORLD])
    orb_mats = mats + _bsdf("det-refraction", "glass", "#FAFAFA", 1.5) + _brdf("microfacet", "ggx", k_d="#E0C080", roughness="0.3", ref_ior="1.0, 1.5, 0.0") \
        + _brdf("fresnel-blend", "fresnel", k_d="#CACACA", k_s="#333333", k_g='r="10" g="1000"')
    _write_xml(os.path.join(root, "cbox", "orb500k.xml"),
               [_sensor(1920, 1080, 24, 1, accelerator="bvh", extra=['<boolean name="enable_microfacet" value="true"/>']), orb_mats, _AREA,
                room + _obj(S + "orb_outer.obj", "glass") + _obj(S + "orb_mid.obj", "ggx") + _obj(S + "orb_inner.obj", "fresnel"), _WORLD])
    car_mats = mats + _brdf("mod-phong", "paint", k_d="#A02020", k_g="30.0", k_s="#505050") + _brdf("specular", "mirror", k_d="#DEDEDE")
    _write_xml(os.path.join(root, "cbox", "car290k.xml"),
               [_sensor(3840, 2160, 16, 1, accelerator="bvh"), car_mats, _AREA,
                room + _obj(S + "car290k.obj", "paint") + _obj(S + "mirror_patch.obj", "mirror"), _WORLD])


def ensure_small_scenes(root: str):
    """Writes the Cornell meshes + small XML scenes if missing. Returns root."""
    marker = os.path.join(root, "test", "allbxdf.xml")
    if not os.path.exists(marker):
        write_cornell_meshes(os.path.join(root, "meshes", "cornell"))
        write_cbox(root)
        write_balls_mono(root)
        write_allbxdf(root)
        write_big_xml(root)
    if not os.path.exists(os.path.join(root, "test", "media.xml")):
        write_media(root)
    return root


def ensure_big_meshes(root: str, which: Sequence[str] = ("bunny90k",)):
    """Generates the analytic meshes of BASELINE configs 3-5 on demand (tens of MB of OBJ text: not committed)."""
    d = os.path.join(root, "meshes", "synth")
    ensure_small_scenes(root)
    if "bunny90k" in which and not os.path.exists(os.path.join(d, "bunny90k.obj")):
        v, n, f = bunny90k_mesh()
        write_obj(os.path.join(d, "bunny90k.obj"), "bunny90k", v, f, normals=n)
    if "orb500k" in which and not os.path.exists(os.path.join(d, "orb_inner.obj")):
        for name, r, b in (("orb_outer", 1.3, 0.0), ("orb_mid", 0.95, 0.03), ("orb_inner", 0.6, 0.05)):
            v, n, f = orb_shell_mesh(r, b)
            write_obj(os.path.join(d, name + ".obj"), name, v, f, normals=n)
    if "car290k" in which and not os.path.exists(os.path.join(d, "car290k.obj")):
        v, n, f = car290k_mesh()
        write_obj(os.path.join(d, "car290k.obj"), "car290k", v, f, normals=n)
        pv = np.float64([(1.0, 0.002, 1.2), (1.0, 0.002, 4.4), (4.6, 0.002, 4.4), (4.6, 0.002, 1.2)])
        write_obj(os.path.join(d, "mirror_patch.obj"), "mirror_patch", pv, np.int64([(0, 1, 2), (0, 2, 3)]),
                  normals=np.float64([(0, 1, 0)]), face_normals=np.int64([(0, 0, 0), (0, 0, 0)]))
    return root


DEFAULT_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")

if __name__ == "__main__":
    import sys
    root = DEFAULT_ROOT
    if "--root" in sys.argv:
        root = sys.argv[sys.argv.index("--root") + 1]
    write_cornell_meshes(os.path.join(root, "meshes", "cornell"))
    write_cbox(root); write_balls_mono(root); write_allbxdf(root); write_media(root); write_big_xml(root)
    if "--big" in sys.argv:
        ensure_big_meshes(root, ("bunny90k", "orb500k", "car290k"))
    print("scenes written under", root)
