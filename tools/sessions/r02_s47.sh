#!/bin/bash
# Round-2 session 47: load-based leaf prefetch in k_trace (pf3), L1::no_allocate pool loads in k_logic (na1) with two lanes and fewer
# resident trace blocks (logic and trace blocks of the two lanes share an SM's L1)
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=$PWD/adapt_b200/lib
bash tools/ab.sh "--spp-per-step 32" "ADAPT_LANES=1" "ADAPT_LANES=1 ADAPT_B200_LIB=$L/pf3.so" "ADAPT_LANES=1 ADAPT_B200_LIB=$L/na1.so"
bash tools/ab.sh "--spp-per-step 256" "ADAPT_B200_LIB=$L/pf3.so" "ADAPT_B200_LIB=$L/na1.so" "ADAPT_TRACE_BLOCKS_PER_SM=5" \
  "ADAPT_TRACE_BLOCKS_PER_SM=5 ADAPT_B200_LIB=$L/na1.so" "ADAPT_TRACE_BLOCKS_PER_SM=6 ADAPT_B200_LIB=$L/na1.so" \
  "ADAPT_TRACE_BLOCKS_PER_SM=5 ADAPT_POOL=16777216" "ADAPT_TRACE_BLOCKS_PER_SM=5 ADAPT_POOL=16777216 ADAPT_B200_LIB=$L/na1.so"
bash tools/ab.sh "--workload orb500k --spp-per-step 128" "ADAPT_B200_LIB=$L/na1.so" "ADAPT_TRACE_BLOCKS_PER_SM=6 ADAPT_B200_LIB=$L/na1.so"
cp gpurun_out/ab.txt gpurun_out/r03h_ab_pf3_noalloc.txt
