"""Driver-level host logic that needs no GPU: CLI flags (reference parsers/opts.py:15-44), the image writer's
orientation (ti.tools.imwrite semantics) and the watermark stamp (utils/watermark.py:22-33)."""
import os

import numpy as np

from conftest import ROOT


def test_cli_flags_match_reference(tmp_path):
    from adapt_b200.parsers.opts import get_options
    o = get_options(argv=["--scene", "csphere", "--name", "balls-mono.xml", "--iter_num", "64", "--no_gui", "--no_watermark",
                          "--img_name", "x", "--save_iter", "16", "-l"])
    assert (o.scene, o.name, o.iter_num, o.no_gui, o.no_watermark, o.img_name, o.save_iter, o.load) == \
        ("csphere", "balls-mono.xml", 64, True, True, "x", 16, True)
    assert o.type == "vpt" and get_options(argv=[]).name == "complex.xml" and o.input_path == "./scenes/" and o.output_path == "./outputs/" and o.chkpt_path == "./checkpoint/"
    cfg = tmp_path / "run.conf"
    cfg.write_text("scene = cbox\nname = cbox.xml\niter_num = 8\nno_gui = true\n# comment\n")
    o = get_options(argv=["--config", str(cfg), "--iter_num", "12"])
    assert o.scene == "cbox" and o.no_gui and o.iter_num == 12            # command line wins over the file


def test_imwrite_orientation(tmp_path):
    import cv2
    import sys
    sys.path.insert(0, ROOT)
    from render import imwrite
    img = np.zeros((4, 3, 3), np.float32)          # (w=4, h=3), indexed [x, y], y up
    img[0, 2] = [1, 0, 0]                          # x = 0, top row -> red at file row 0, col 0
    img[3, 0] = [0, 0, 2.0]                        # x = 3, bottom row, clipped blue
    p = str(tmp_path / "o.png")
    imwrite(img, p)
    out = cv2.imread(p)[..., ::-1]
    assert out.shape == (3, 4, 3)
    assert out[0, 0].tolist() == [255, 0, 0] and out[2, 3].tolist() == [0, 0, 255]


def test_watermark_stamp():
    from adapt_b200.utils.watermark import apply_watermark, water_mark
    assert water_mark.shape[0] == 7 and set(np.unique(water_mark)) == {0.0, 1.0}

    class _Pix:
        def to_numpy(self):
            return np.full((256, 128, 3), 0.25, np.float32)

    class _R:
        pixels = _Pix(); do_crop = False; start_x = start_y = 0; end_x = 256; end_y = 128
    out = apply_watermark(_R(), 0.0, False, True)
    h, w = water_mark.shape
    np.testing.assert_allclose(out[-w - 1:-1, :h, 0], 0.25 + water_mark.T)
    plain = apply_watermark(_R(), 0.0, False, False)
    assert np.all(plain == 0.25)
    _R.do_crop = True; _R.start_x, _R.end_x, _R.start_y, _R.end_y = 10, 20, 30, 50
    assert apply_watermark(_R(), 0.0, False, True).shape == (20, 10, 3)    # reference crops [start_y:end_y, start_x:end_x]
