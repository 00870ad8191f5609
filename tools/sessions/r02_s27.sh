#!/bin/bash
# Round-2 session 27: resident trace blocks per SM below the occupancy limit (the kernel is L1-capacity bound: fewer rays in flight?)
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
bash tools/ab.sh "" ADAPT_TRACE_BLOCKS_PER_SM=8 ADAPT_TRACE_BLOCKS_PER_SM=7 ADAPT_TRACE_BLOCKS_PER_SM=6
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_TRACE_BLOCKS_PER_SM=7 ADAPT_TRACE_BLOCKS_PER_SM=6 ADAPT_TRACE_BLOCKS_PER_SM=5
cp gpurun_out/ab.txt gpurun_out/r02z_ab_trace_blocks.txt
