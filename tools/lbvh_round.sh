#!/bin/bash
# GPU session for the device BVH builder: its tests (-s: the update test prints build / rebuild times), then build-time figures.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_lbvh.py -q -x -s --timeout 150 2>&1 | tail -15 | tee gpurun_out/pytest_lbvh.log
timeout 300 python tools/bvh_build_bench.py bunny90k orb500k 2>&1 | tail -6 | tee gpurun_out/bvh_build.txt
bash tools/ab.sh "" ADAPT_NODE_STEPS=2 ADAPT_NODE_STEPS=3
