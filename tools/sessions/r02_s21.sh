#!/bin/bash
# Round-2 session 21: device SAH builder (bvh_builder = 2): GPU tests, build time and traversal time against the host SAH tree and the LBVH
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
timeout 900 python -m pytest tests/test_gpu_lbvh.py -q -m gpu --timeout 300 -s 2>&1 | grep -v "^$" | tail -15 | tee gpurun_out/r02t_pytest_gpu_lbvh.txt
export ADAPT_TRACE_MODE=1
bash tools/ab.sh "" ADAPT_BVH_BUILDER=1 ADAPT_BVH_BUILDER=2
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_BVH_BUILDER=1 ADAPT_BVH_BUILDER=2
bash tools/ab.sh "--workload car290k --spp-per-step 4" ADAPT_BVH_BUILDER=1 ADAPT_BVH_BUILDER=2
cp gpurun_out/ab.txt gpurun_out/r02t_ab_device_sah.txt
