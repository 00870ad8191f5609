#!/bin/bash
# Round-2 session 42: lanes x trace blocks per SM, three workloads
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
bash tools/ab.sh "--spp-per-step 256" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=4" "ADAPT_LANES=3 ADAPT_TRACE_BLOCKS_PER_SM=5" "ADAPT_LANES=3 ADAPT_TRACE_BLOCKS_PER_SM=9" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=5 ADAPT_POOL=4194304" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=5 ADAPT_POOL=16777216"
bash tools/ab.sh "--workload orb500k --spp-per-step 128" "ADAPT_LANES=2" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=7" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=5"
bash tools/ab.sh "--workload car290k --spp-per-step 32" "ADAPT_LANES=2" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=6"
cp gpurun_out/ab.txt gpurun_out/r02zm_ab_lanes_grid.txt
