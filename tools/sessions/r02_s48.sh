#!/bin/bash
# Round-2 session 48: k_trace at six resident blocks per SM while two lanes are active (default) against nine (before) and five
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
for spp in 32 64 256; do bash tools/ab.sh "--spp-per-step $spp" "ADAPT_TRACE_BLOCKS_2LANES=9" "ADAPT_TRACE_BLOCKS_2LANES=5"; done
bash tools/ab.sh "--workload orb500k --spp-per-step 256" "ADAPT_TRACE_BLOCKS_2LANES=9"
bash tools/ab.sh "--workload car290k --spp-per-step 32" "ADAPT_TRACE_BLOCKS_2LANES=9"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 256" "ADAPT_TRACE_BLOCKS_2LANES=9"
cp gpurun_out/ab.txt gpurun_out/r03i_ab_blocks_2lanes.txt
