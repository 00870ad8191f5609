#!/bin/bash
# Round-2 session 53: k_logic (one material group) in 64-thread blocks against the shipped 128
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=$PWD/adapt_b200/lib
bash tools/ab.sh "--spp-per-step 256" "ADAPT_B200_LIB=$L/lb64.so"
bash tools/ab.sh "--spp-per-step 32" "ADAPT_B200_LIB=$L/lb64.so"
cp gpurun_out/ab.txt gpurun_out/r03n_ab_logic_block64.txt
