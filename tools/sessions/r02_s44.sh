#!/bin/bash
# Round-2 session 44: two lanes at small batches (does the second pool's drain cost more than the overlap wins?)
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
bash tools/ab.sh "--spp-per-step 8" "ADAPT_LANES=2"
bash tools/ab.sh "--spp-per-step 32" "ADAPT_LANES=2"
bash tools/ab.sh "--spp-per-step 64" "ADAPT_LANES=2"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" "ADAPT_LANES=2"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 64" "ADAPT_LANES=2"
cp gpurun_out/ab.txt gpurun_out/r02zo_ab_lanes_small_batches.txt
