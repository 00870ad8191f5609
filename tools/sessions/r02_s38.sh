#!/bin/bash
# Round-2 session 38: binary against 8-wide tree per scene size with the final kernels (where should the automatic choice switch?)
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
bash tools/ab.sh "" ADAPT_TRACE_MODE=3
bash tools/ab.sh "--workload car290k --spp-per-step 4" ADAPT_TRACE_MODE=3 "ADAPT_TRACE_MODE=3 ADAPT_REFILL=8 ADAPT_LEAF_T=4" "ADAPT_TRACE_MODE=3 ADAPT_REFILL=12"
cp gpurun_out/ab.txt gpurun_out/r02zj_ab_tree_choice.txt
