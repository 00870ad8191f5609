#!/bin/bash
# Round-2 session 33: vpt trace as two launches (transmittance stream; closest-hit stream through the 56-register pt kernel) against the fused one
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
bash tools/ab.sh "--integrator vpt --workload cbox --width 1024 --height 1024 --spp-per-step 16" ADAPT_FUSE_TRACE_VPT=0
bash tools/ab.sh "--integrator vpt --workload media --width 1024 --height 1024 --spp-per-step 16" ADAPT_FUSE_TRACE_VPT=0
cp gpurun_out/ab.txt gpurun_out/r03d_ab_vpt_trace_split.txt
ADAPT_FUSE_TRACE_VPT=0 timeout 600 python -m pytest tests/test_gpu_vpt.py -q -m gpu --timeout 300 2>&1 | tail -3
