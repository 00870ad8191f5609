"""Host scene ingestion: the re-hosted parsers must produce the reference's 4-tuple
(parsers/xml_parser.py:246-289) for the shipped Cornell fixtures (SURVEY Appendix D numbers)."""
import os
import xml.etree.ElementTree as xet

import numpy as np
import pytest

from conftest import load_scene


def test_cbox_tuple(scene_root):
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml")
    assert a["primitives"].shape == (34, 3, 3) and a["primitives"].dtype == np.float32
    assert a["indices"] is None
    assert a["n_g"].shape == (34, 3) and a["n_s"].shape == (34, 3, 3) and a["uvs"].shape == (34, 3, 2)
    assert len(o) == 7 and [ob.tri_num for ob in o] == [2, 2, 2, 2, 2, 12, 12]
    assert len(e) == 1 and e[0].type == "point"
    np.testing.assert_allclose(e[0].intensity, [12, 12, 12], rtol=1e-6)          # 60 * 0.2
    np.testing.assert_allclose(e[0].pos, [2.779, 4.5, 3.0], rtol=1e-6)
    assert c["max_bounce"] == 12 and c["num_shadow_ray"] == 1 and c["use_rr"] and c["use_mis"]
    assert c["film"]["width"] == 512 and c["has_vertex_normal"]
    assert c["world"].medium.ior == 1.0
    # geometric normals are unit, floor faces +y
    np.testing.assert_allclose(np.linalg.norm(a["n_g"], axis=1), 1.0, atol=1e-6)
    np.testing.assert_allclose(a["n_g"][0], [0, 1, 0], atol=1e-6)
    # planar AABBs are padded by 2e-2 (parsers/obj_desc.py:17-21)
    np.testing.assert_allclose(o[0].aabb[:, 1], [-0.02, 0.02], atol=1e-7)


def test_balls_mono_tuple(scene_root):
    e, a, o, c = load_scene(scene_root, "csphere", "balls-mono.xml")
    assert a["primitives"].shape == (18, 3, 3)
    np.testing.assert_array_equal(a["indices"], [12, 13, 14, 15, 16, 17])
    assert [ob.type for ob in o] == [0] * 6 + [1] * 6
    assert abs(e[0].inv_area - 1.0 / 1.365) < 1e-6 and e[0].attached
    np.testing.assert_allclose(e[0].intensity, np.float32([70.0, 63.2, 60.3]) * np.float32(0.6), rtol=1e-6)
    types = [ob.bsdf.type for ob in o]
    assert types[6:] == ["specular", "fresnel-blend", "lambertian", "lambertian", "mod-phong", "det-refraction"]
    fb = o[7].bsdf.export()
    # k_g = (n_u, n_v, sqrt((n_u+1)(n_v+1)) / 8pi); missing b channel defaults to 0 (quirk 14)
    assert fb["k_g"][0] == 10 and fb["k_g"][1] == 1000
    assert abs(fb["k_g"][2] - np.sqrt(11 * 1001) / (8 * np.pi)) < 1e-5
    glass = o[11].bsdf.export()
    assert glass["kind"] == 1 and glass["type"] == 0 and glass["is_delta"] == 1 and abs(glass["ior"] - 1.5) < 1e-7
    mirror = o[6].bsdf.export()
    assert mirror["is_delta"] == 1 and mirror["type"] == 2
    # sphere primitive layout: (center, (r,r,r), 0)
    np.testing.assert_allclose(a["primitives"][12], [[4.5, 0.6, 1.1], [0.6, 0.6, 0.6], [0, 0, 0]], rtol=1e-6)


def test_camera_constants(scene_root):
    from adapt_b200._lib import pack_scene
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml")
    ps = pack_scene(e, a, o, c)
    assert abs(ps.host["focal"] - 716.80) < 0.01
    np.testing.assert_allclose(ps.host["cam_r"], np.eye(3), atol=1e-7)
    np.testing.assert_allclose(ps.host["cam_t"], [2.78, 2.73, -8.0], rtol=1e-6)
    d = ps.desc
    assert (d.n_prims, d.n_objects, d.n_emitters, d.width, d.height) == (34, 7, 1, 512, 512)
    assert d.accelerator == 0                                # no accelerator requested in cbox.xml


def test_rgb_parse_variants():
    from adapt_b200.parsers.general_parser import rgb_parse
    el = lambda **kw: xet.Element("rgb", {k: str(v) for k, v in kw.items()})     # noqa: E731
    np.testing.assert_allclose(rgb_parse(el(value="#FF8000")), [1.0, 128 / 255, 0.0], rtol=1e-6)
    np.testing.assert_allclose(rgb_parse(el(value="0.5")), [0.5] * 3)
    np.testing.assert_allclose(rgb_parse(el(value="1, 2, 3")), [1, 2, 3])
    np.testing.assert_allclose(rgb_parse(el(r="10", g="1000")), [10, 1000, 0])
    with pytest.raises(ValueError):
        rgb_parse(xet.Element("rgb"))
    with pytest.raises(ValueError):
        rgb_parse(None)


def test_obj_loader_quads_and_negative_indices(tmp_path):
    from adapt_b200.parsers.obj_loader import extract_obj_info
    p = tmp_path / "quad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf -4//1 -3//1 -2//1 -1//1\n")
    faces, normals, vns, uvs = extract_obj_info(str(p), verbose=False)
    assert faces.shape == (2, 3, 3) and uvs is None and vns.shape == (2, 3, 3)
    np.testing.assert_allclose(faces[0], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    np.testing.assert_allclose(faces[1], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    np.testing.assert_allclose(normals, [[0, 0, 1], [0, 0, 1]])


def test_first_material_only(tmp_path):
    from adapt_b200.parsers.obj_loader import extract_obj_info
    p = tmp_path / "two.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nusemtl a\nf 1 2 3\nusemtl b\nf 1 2 4\n")
    faces, *_ = extract_obj_info(str(p), verbose=False)
    assert faces.shape == (1, 3, 3)                          # obj_loader.py:35-37, quirk 6


def test_error_behaviour(tmp_path, scene_root):
    from adapt_b200.parsers.xml_parser import scene_parsing
    bad = tmp_path / "bad.xml"
    bad.write_text('<scene version="0.9"><sensor/></scene>')
    with pytest.raises(ValueError):
        scene_parsing(str(tmp_path), "bad.xml")
    src = open(os.path.join(scene_root, "cbox", "cbox.xml")).read()
    tex = tmp_path / "tex.xml"
    tex.write_text(src.replace("</scene>", '<texture id="t" tag="albedo"/></scene>'))
    with pytest.raises(AttributeError):          # image texture without a <string> child: same failure as bxdf/texture.py:59
        scene_parsing(str(tmp_path), "tex.xml")


def test_transform_semantics():
    """Rotation about the mesh centroid by right-multiplication; scale parsed but ignored (quirk 5)."""
    from adapt_b200.parsers.obj_loader import apply_transform
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(1)
    m = rng.normal(size=(5, 3, 3)).astype(np.float32)
    R = Rot.from_euler("zxy", (10, 20, 30), degrees=True).as_matrix()
    out, _ = apply_transform(m.copy(), None, R, np.float32([1, 2, 3]), np.float32([2, 2, 2]))
    c = m.mean(axis=1).mean(axis=0)
    np.testing.assert_allclose(out, (m - c) @ R + c + np.float32([1, 2, 3]), rtol=1e-5, atol=1e-5)


def test_tile_partition_covers_film():
    from adapt_b200.dist import tile_partition
    w, h = 100, 70
    for world in (1, 2, 3, 8):
        parts = [tile_partition(w, h, r, world) for r in range(world)]
        allp = np.concatenate(parts)
        assert len(allp) == w * h and len(np.unique(allp)) == w * h
    win = tile_partition(64, 64, 0, 1, window=(8, 24, 16, 48))
    ii, jj = win // 64, win % 64
    assert ii.min() == 8 and ii.max() == 23 and jj.min() == 16 and jj.max() == 47 and len(win) == 16 * 32


# ------------------------------------------------------------------------------------------------ textures
def test_texture_scene_parses_and_packs(scene_root):
    from conftest import load_scene
    e, a, o, c = load_scene(scene_root, "test", "textured.xml")
    packed = c["packed_textures"]
    assert set(packed) == {"albedo", "normal", "bump", "roughness"} and packed["roughness"] is None
    for key in ("albedo", "normal", "bump"):
        img = packed[key]
        assert img.dtype == np.float32 and img.shape[0] == img.shape[1] and img.shape[2] == 3
    floor = o[1].texture_group["albedo"]
    assert (floor.w, floor.h, floor.scale_u, floor.scale_v) == (48, 32, 2.0, 1.5)
    # the rectangle in the atlas holds the image (RGB, /255), bump maps with G and B swapped (bxdf/texture.py:77-79)
    import cv2
    raw = cv2.cvtColor(cv2.imread(os.path.join(scene_root, "textures", "tex_albedo.png")), cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0
    np.testing.assert_array_equal(packed["albedo"][floor.off_y:floor.off_y + 32, floor.off_x:floor.off_x + 48], raw)
    bump = o[7].texture_group["bump"]
    rawb = cv2.cvtColor(cv2.imread(os.path.join(scene_root, "textures", "tex_bump.png")), cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0
    np.testing.assert_array_equal(packed["bump"][bump.off_y:bump.off_y + bump.h, bump.off_x:bump.off_x + bump.w], rawb[..., [0, 2, 1]])
    # mesh uv coordinates reach array_info; objects without vt get zeros
    assert a["uvs"].shape == (a["primitives"].shape[0], 3, 2) and a["uvs"][2:4].any() and not a["uvs"][4:6].any()


def test_shelf_packer_places_without_overlap():
    from adapt_b200.parsers.texture_packing import shelf_pack
    rects = [(300, 200, 0), (500, 120, 1), (64, 64, 2), (400, 400, 3), (100, 700, 4)]
    placed = shelf_pack(rects, 1024)
    assert placed is not None and set(placed) == {0, 1, 2, 3, 4}
    occ = np.zeros((1024, 1024), np.int32)
    for w, h, rid in rects:
        x, y = placed[rid]
        assert x >= 0 and y >= 0 and x + w <= 1024 and y + h <= 1024
        occ[y:y + h, x:x + w] += 1
    assert occ.max() == 1
    assert shelf_pack([(800, 800, 0), (800, 800, 1)], 1024) is None and shelf_pack([(2000, 10, 0)], 1024) is None


def test_texture_errors(scene_root, tmp_path):
    import shutil
    from adapt_b200.parsers.xml_parser import scene_parsing
    src = open(os.path.join(scene_root, "test", "textured.xml")).read()
    d = tmp_path / "test"
    d.mkdir()
    shutil.copytree(os.path.join(scene_root, "meshes"), tmp_path / "meshes", ignore=shutil.ignore_patterns("synth"))
    shutil.copytree(os.path.join(scene_root, "textures"), tmp_path / "textures")
    (d / "missing.xml").write_text(src.replace("../textures/tex_bump.png", "../textures/nope.png"))
    with pytest.raises(ValueError):
        scene_parsing(str(d), "missing.xml")
    (d / "badref.xml").write_text(src.replace('<ref type="texture" id="dents" tag="bump"/>', '<ref type="texture" id="dents" tag="normal"/>', 1))
    with pytest.raises(KeyError):
        scene_parsing(str(d), "badref.xml")
    checker = '<texture id="chk" type="checkerboard" tag="albedo"><rgb name="c1" value="#000000"/><rgb name="c2" value="#FFFFFF"/></texture>'
    (d / "checker.xml").write_text(src.replace("<emitter type=\"area\"", checker + "<emitter type=\"area\"", 1)
                                   .replace('<ref type="texture" id="paint2" tag="albedo"/>', '<ref type="texture" id="chk" tag="albedo"/>', 1))
    with pytest.raises(NotImplementedError):
        scene_parsing(str(d), "checker.xml")


def test_microfacet_flag_does_not_leak_between_scenes(scene_root):
    """The reference compiles GGX out unless its source flag is flipped (bxdf/brdf.py:8); here a sensor key of the scene switches it on,
    and the next scene parsed by the same process starts from the reference's default (off) again."""
    from adapt_b200.bxdf import brdf
    load_scene(scene_root, "test", "allbxdf.xml", 8, 8)            # <boolean name="enable_microfacet" value="true"/>
    assert brdf.microfacet_enabled()
    load_scene(scene_root, "cbox", "cbox.xml", 8, 8)
    assert not brdf.microfacet_enabled()


def test_every_rank_owns_a_tile_and_empty_lists_are_refused(scene_root):
    """A 64 x 64 film has four 32 x 32 tiles: eight ranks need a smaller tile (auto_tile), and a handle must never be created from an
    empty pixel list (n_pixels = 0 means `whole film` to adapt_create: the reduce would count those pixels once per rank)."""
    from adapt_b200._lib import pack_scene
    from adapt_b200.dist import auto_tile, tile_partition
    assert tile_partition(64, 64, 5, 8).size == 0
    t = auto_tile(64, 64, 8)
    assert t == 16
    parts = [tile_partition(64, 64, r, 8, tile=t) for r in range(8)]
    assert all(p.size > 0 for p in parts) and np.array_equal(np.sort(np.concatenate(parts)), np.arange(64 * 64))
    assert auto_tile(1920, 1080, 8) == 32 and auto_tile(640, 480, 2, window=(100, 120, 100, 120)) == 16
    e, a, o, c = load_scene(scene_root, "cbox", "cbox.xml", 64, 64)
    with pytest.raises(ValueError):
        pack_scene(e, a, o, c, pixel_list=tile_partition(64, 64, 5, 8))
