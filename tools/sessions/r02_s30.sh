#!/bin/bash
# Round-2 session 30: compressed 8-wide tree collapsed on the device from the device-SAH hierarchy
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
timeout 900 python -m pytest tests/test_gpu_lbvh.py -q -m gpu --timeout 300 2>&1 | tail -6 | tee gpurun_out/r03b_pytest_gpu_lbvh.txt
ADAPT_BVH_BUILDER=2 ADAPT_TRACE_MODE=3 timeout 1200 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_lbvh.py 2>&1 | tail -6 | tee gpurun_out/r03b_pytest_gpu_builder2_cw8.txt
BUILDERS=sah,sah_device ADAPT_TRACE_MODE=3 timeout 300 python tools/bvh_build_bench.py bunny90k orb500k car290k 2>&1 | tee gpurun_out/r03b_bvh_build_cw8.txt
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_BVH_BUILDER=2
ADAPT_TRACE_MODE=3 bash tools/ab.sh "" ADAPT_BVH_BUILDER=2
ADAPT_TRACE_MODE=3 bash tools/ab.sh "--workload car290k --spp-per-step 4" ADAPT_BVH_BUILDER=2
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" ADAPT_BVH_BUILDER=2
cp gpurun_out/ab.txt gpurun_out/r03b_ab_device_cw8.txt
