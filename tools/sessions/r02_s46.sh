#!/bin/bash
# Round-2 session 46: L1 prefetch variants of k_trace (leaf records when a lane lands on a leaf; the far child when it is pushed)
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=$PWD/adapt_b200/lib
V="ADAPT_B200_LIB=$L/pf1.so ADAPT_B200_LIB=$L/pf2.so ADAPT_B200_LIB=$L/far1.so"
bash tools/ab.sh "" $V "ADAPT_B200_LIB=$L/pf1.so ADAPT_LEAF_T=12"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" $V
bash tools/ab.sh "--workload car290k --spp-per-step 16" $V
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" $V
bash tools/ab.sh "--spp-per-step 256" ADAPT_B200_LIB=$L/pf1.so ADAPT_B200_LIB=$L/pf2.so
cp gpurun_out/ab.txt gpurun_out/r03g_ab_prefetch.txt
