#!/bin/bash
# Round-2 session 50: k_logic in 128-thread blocks (8 per SM) so that more of them fit beside the trace blocks of the other lane
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=$PWD/adapt_b200/lib
bash tools/ab.sh "--spp-per-step 256" "ADAPT_B200_LIB=$L/lb128.so" "ADAPT_B200_LIB=$L/lb128.so ADAPT_TRACE_BLOCKS_2LANES=5" "ADAPT_B200_LIB=$L/lb128.so ADAPT_TRACE_BLOCKS_2LANES=7" "ADAPT_B200_LIB=$L/lb128.so ADAPT_TRACE_BLOCKS_2LANES=4"
bash tools/ab.sh "--spp-per-step 32" "ADAPT_LANES=1" "ADAPT_LANES=1 ADAPT_B200_LIB=$L/lb128.so"
bash tools/ab.sh "--workload orb500k --spp-per-step 256" "ADAPT_B200_LIB=$L/lb128.so" "ADAPT_B200_LIB=$L/lb128.so ADAPT_TRACE_BLOCKS_2LANES=5"
cp gpurun_out/ab.txt gpurun_out/r03k_ab_logic_block128.txt
