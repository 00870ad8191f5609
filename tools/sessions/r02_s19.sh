#!/bin/bash
# Round-2 session 19: L1 policy of k_trace's tree fetches (evict_last nodes, no_allocate / evict_first leaf records) and the L1 / shared split
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=/root/repo/adapt_b200/lib
bash tools/ab.sh "" ADAPT_B200_LIB=$L/v0_base.so ADAPT_TRACE_CARVEOUT=0 ADAPT_B200_LIB=$L/v1_nodelast.so ADAPT_B200_LIB=$L/v2_primnoalloc.so ADAPT_B200_LIB=$L/v3_primfirst.so ADAPT_B200_LIB=$L/v4_nodelast_primnoalloc.so
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_TRACE_CARVEOUT=0
ADAPT_TRACE_MODE=1 bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_TRACE_CARVEOUT=0 ADAPT_B200_LIB=$L/v4_nodelast_primnoalloc.so ADAPT_B200_LIB=$L/v2_primnoalloc.so
cp gpurun_out/ab.txt gpurun_out/r02s_ab_l1_policy.txt
