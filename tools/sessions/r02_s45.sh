#!/bin/bash
# Round-2 session 45: adaptive second lane (default): batch-size scan against one lane, the GPU suite, vpt, bench line
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
for spp in 8 16 32 64 256; do bash tools/ab.sh "--spp-per-step $spp" "ADAPT_LANES=1"; done
bash tools/ab.sh "--workload orb500k --spp-per-step 16" "ADAPT_LANES=1"
bash tools/ab.sh "--workload orb500k --spp-per-step 128" "ADAPT_LANES=1"
bash tools/ab.sh "--integrator vpt --workload cbox --width 1024 --height 1024 --spp-per-step 64" "ADAPT_LANES=1"
cp gpurun_out/ab.txt gpurun_out/r02zp_ab_adaptive_lanes.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02zp_pytest_gpu.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02zp_bench.json 2> gpurun_out/bench.err; tail -c 500 gpurun_out/r02zp_bench.json; tail -2 gpurun_out/bench.err
