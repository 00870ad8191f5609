#!/bin/bash
# Round-2 session 41: two pools on two streams with room left beside the persistent trace kernel (fewer trace blocks per SM), 256 spp per step
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
bash tools/ab.sh "--spp-per-step 256" "ADAPT_LANES=2" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=8" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=7" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=6" "ADAPT_LANES=2 ADAPT_TRACE_BLOCKS_PER_SM=5"
cp gpurun_out/ab.txt gpurun_out/r02zl_ab_lanes_room.txt
