import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("ADAPT_QUIET", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def scene_root():
    from adapt_b200.scenes import DEFAULT_ROOT, ensure_small_scenes
    return ensure_small_scenes(DEFAULT_ROOT)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pt_oracle
    pt_oracle.build()
    return pt_oracle.load()


def load_scene(scene_root, scene, name, width=None, height=None, **overrides):
    from adapt_b200.parsers.xml_parser import scene_parsing
    e, a, o, c = scene_parsing(os.path.join(scene_root, scene), name)
    if width:
        c["film"]["width"] = width
        c["film"]["height"] = height or width
    c.update(overrides)
    return e, a, o, c


def rel_l2(a, b):
    import numpy as np
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))
