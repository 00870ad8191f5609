#!/bin/bash
# Round-2 session 14: ray binning experiment (ADAPT_RAY_BINS=1): parity, then A/B
mkdir -p gpurun_out
ADAPT_RAY_BINS=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_bins.log
rm -f gpurun_out/ab.txt
bash tools/ab.sh "" "ADAPT_RAY_BINS=1"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" "ADAPT_RAY_BINS=1" "ADAPT_RAY_BINS=1 ADAPT_TRACE_MODE=1" "ADAPT_TRACE_MODE=1"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" "ADAPT_RAY_BINS=1"
bash tools/ab.sh "--workload car290k --spp-per-step 4" "ADAPT_RAY_BINS=1"
