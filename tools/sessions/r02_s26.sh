#!/bin/bash
# Round-2 session 26 (final library): smoke, whole GPU suite, both bench arms, launch list, compute-sanitizer on the device SAH builder
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -6 | tee gpurun_out/r02y_pytest_gpu.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02y_bench.json 2> gpurun_out/bench.err; tail -c 4500 gpurun_out/r02y_bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/r02y_bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/r02y_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r02y_launches.csv "bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 under ncu (serialised, cold-cache launches; shares are what matter)" | tee gpurun_out/r02y_launch_summary.txt
cat > /tmp/sah_small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); os.environ.setdefault("ADAPT_QUIET", "1")
from adapt_b200.parsers.xml_parser import scene_parsing
from adapt_b200.renderer.vanilla_renderer import Renderer
from adapt_b200.scenes import DEFAULT_ROOT, ensure_small_scenes
root = ensure_small_scenes(DEFAULT_ROOT)
for scene, name in (("cbox", "cbox.xml"), ("test", "allbxdf.xml")):
    e, a, o, c = scene_parsing(os.path.join(root, scene), name)
    c["film"]["width"] = c["film"]["height"] = 16
    r = Renderer(e, a, o, c, bvh_builder="sah_device", pool_size=1024)
    r.render_batch(1); r.pixels.to_numpy()
    r.update_geometry(a["primitives"], a["n_g"], a["n_s"]); r.render_batch(1); r.pixels.to_numpy()
    print(scene, name, r.bvh_export(arrays=False)["n_nodes"]); r.close()
PY
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python /tmp/sah_small.py > gpurun_out/r02y_sanitizer_sah_$tool.log 2>&1; tail -3 gpurun_out/r02y_sanitizer_sah_$tool.log
done
