#!/bin/bash
# Round-2 session 1: vpt bring-up, compute-sanitizer on the shipped pt kernels, A/B of the two round-1 experiments
# (variants prebuilt in adapt_b200/lib/{postpone,colorred}/ so no box time goes into nvcc).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi

echo "=== vpt bring-up"
bash tools/vpt_round.sh 2>&1 | tail -60

echo "=== sanitizer on the pt kernels (tiny films)"
cat > /tmp/pt_tiny.py <<'PY'
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
for scene, name in (("cbox", "cbox.xml"), ("csphere", "balls-mono.xml"), ("test", "allbxdf.xml"), ("test", "textured.xml")):
    e, a, o, c = load_scene(root, scene, name, 24, 24)
    r = Renderer(e, a, o, c, seed=5, pool_size=1024)
    r.render_batch(2)
    img = r.pixels.to_numpy()
    ref, _ = OracleScene(pack_scene(e, a, o, c, seed=5)).render(2)
    print(name, "rel L2 vs oracle", rel_l2(img, ref / 2))
PY
for tool in memcheck racecheck synccheck initcheck; do
  echo "--- compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python /tmp/pt_tiny.py 2>&1 | tail -12 | tee gpurun_out/sanitizer_pt_$tool.log
done

echo "=== gpu tests"
timeout 600 python -m pytest tests -q -m gpu -x --timeout 120 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log

echo "=== A/B"
rm -f gpurun_out/ab.txt
P="ADAPT_B200_LIB=$PWD/adapt_b200/lib/postpone/libadapt_b200.so"
C="ADAPT_B200_LIB=$PWD/adapt_b200/lib/colorred/libadapt_b200.so"
for V in "$P" "$C"; do
  env $V timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 90 2>&1 | tail -2
done
bash tools/ab.sh "" "$P" "$P ADAPT_LEAF_T=12" "$C"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" "$P" "$C"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" "$P" "$C"
ls -la gpurun_out/
