#!/bin/bash
# Round-2 session 24: BVH nodes / leaf records as a persisting L2 access-policy window of the render streams
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
bash tools/ab.sh "" ADAPT_L2_PERSIST=1 ADAPT_L2_PERSIST=2
ADAPT_TRACE_MODE=1 bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_L2_PERSIST=1 ADAPT_L2_PERSIST=2
cp gpurun_out/ab.txt gpurun_out/r02w_ab_l2_persist.txt
