#!/bin/bash
# GPU session that brings the volumetric integrator's first kernel up: memcheck on a tiny scene first (a hang or an
# out-of-bounds access must not take the box down), then the gated parity tests, then a short timing of a fog scene.
mkdir -p gpurun_out
cat > /tmp/vpt_tiny.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
os.environ.setdefault("ADAPT_QUIET", "1")
import numpy as np
from conftest import load_scene, rel_l2
from adapt_b200.scenes import DEFAULT_ROOT, ensure_small_scenes
from adapt_b200._lib import pack_scene
from adapt_b200.renderer.vanilla_renderer import Renderer
from oracle.pt_oracle import OracleScene
root = ensure_small_scenes(DEFAULT_ROOT)
for scene, name in (("cbox", "cbox.xml"), ("test", "media.xml")):
    e, a, o, c = load_scene(root, scene, name, 24, 24)
    r = Renderer(e, a, o, c, seed=5, integrator="vpt", pool_size=1024)
    r.render_batch(2)
    img = r.pixels.to_numpy()
    ref, _ = OracleScene(pack_scene(e, a, o, c, seed=5, integrator="vpt")).render(2)
    print(name, "rel L2 vs oracle", rel_l2(img, ref / 2), "stats", {k: v for k, v in r.stats().items() if k in ("paths", "rays_closest", "rays_shadow", "iterations")})
PY
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/vpt_tiny.py 2>&1 | tail -25 | tee gpurun_out/vpt_memcheck.log
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "MEMCHECK FAILED - not running the larger tests"; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_vpt.py -q -x --timeout 120 2>&1 | tail -15 | tee gpurun_out/pytest_vpt.log
timeout 300 python bench.py --integrator vpt --workload cbox --width 1024 --steps 3 --warmup 3 --spp-per-step 8 --cpu-budget 8 > gpurun_out/bench_vpt_cbox.json 2> gpurun_out/bench_vpt_cbox.err; tail -c 1500 gpurun_out/bench_vpt_cbox.json
