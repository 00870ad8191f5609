#!/bin/bash
# GPU session: node-steps / vote-scheme sweep of k_trace on the bench workloads, parity suite on the candidate configuration.
mkdir -p gpurun_out
R=$PWD/adapt_b200/lib/redux/libadapt_b200.so
bash tools/ab.sh "" ADAPT_NODE_STEPS=3 ADAPT_NODE_STEPS=4 ADAPT_NODE_STEPS=6 ADAPT_NODE_STEPS=8 "ADAPT_NODE_STEPS=4 ADAPT_LEAF_T=8" "ADAPT_NODE_STEPS=4 ADAPT_LEAF_T=16" \
    "ADAPT_B200_LIB=$R" "ADAPT_B200_LIB=$R ADAPT_NODE_STEPS=3" "ADAPT_B200_LIB=$R ADAPT_NODE_STEPS=4" "ADAPT_B200_LIB=$R ADAPT_NODE_STEPS=6"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_NODE_STEPS=4 "ADAPT_B200_LIB=$R ADAPT_NODE_STEPS=4"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" ADAPT_NODE_STEPS=4 "ADAPT_B200_LIB=$R ADAPT_NODE_STEPS=4"
ADAPT_B200_LIB=$R ADAPT_NODE_STEPS=4 timeout 400 python -m pytest tests -q -m gpu -x --timeout 120 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_redux4.log
