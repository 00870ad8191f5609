"""Regenerates tests/golden/oracle_small.npz from the CPU oracle (the reference itself cannot be
imported here: Taichi is not installed, so these are *oracle* vectors, not reference vectors).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("ADAPT_QUIET", "1")

# tag, scene dir, xml, film size, spp, seed
CASES = [
    ("cbox_64_16", "cbox", "cbox.xml", 64, 16, 0),
    ("mono_64_16", "csphere", "balls-mono.xml", 64, 16, 0),
    ("allbxdf_64_16", "test", "allbxdf.xml", 64, 16, 3),
]

if __name__ == "__main__":
    from adapt_b200._lib import pack_scene
    from adapt_b200.parsers.xml_parser import scene_parsing
    from adapt_b200.scenes import DEFAULT_ROOT, ensure_small_scenes
    from oracle.pt_oracle import OracleScene
    root = ensure_small_scenes(DEFAULT_ROOT)
    out = {}
    for tag, scene, name, size, spp, seed in CASES:
        e, a, o, c = scene_parsing(os.path.join(root, scene), name)
        c["film"]["width"] = size; c["film"]["height"] = size
        acc, _ = OracleScene(pack_scene(e, a, o, c, seed=seed)).render(spp)
        out[tag] = (acc / spp).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **out)
    print("wrote", list(out))
