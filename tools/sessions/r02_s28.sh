#!/bin/bash
# Round-2 session 28: the whole GPU suite with every handle's tree built by the device SAH builder (ADAPT_BVH_BUILDER=2; binary traversal)
mkdir -p gpurun_out
ADAPT_BVH_BUILDER=2 timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -12 | tee gpurun_out/r02z_pytest_gpu_builder2.txt
