#!/bin/bash
# Round-2 session 56: where the second lane breaks even now (six trace blocks beside 128-thread logic blocks): 8 / 16 spp per synchronisation
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
bash tools/ab.sh "--spp-per-step 16" "ADAPT_LANES=2" "ADAPT_LANE_THRESHOLD=6"
bash tools/ab.sh "--spp-per-step 8" "ADAPT_LANES=2"
cp gpurun_out/ab.txt gpurun_out/r03q_ab_lane_threshold.txt
