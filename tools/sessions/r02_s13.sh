#!/bin/bash
# Round-2 session 13: camera rays culled against the scene box in k_logic (ADAPT_CULL_PRIMARY=1): parity + A/B
mkdir -p gpurun_out
ADAPT_CULL_PRIMARY=1 timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_cull.log
rm -f gpurun_out/ab.txt
bash tools/ab.sh "" "ADAPT_CULL_PRIMARY=1"
bash tools/ab.sh "--spp-per-step 128" "ADAPT_CULL_PRIMARY=1"
bash tools/ab.sh "--workload orb500k --spp-per-step 64" "ADAPT_CULL_PRIMARY=1"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 64" "ADAPT_CULL_PRIMARY=1"
bash tools/ab.sh "--workload car290k --spp-per-step 16" "ADAPT_CULL_PRIMARY=1"
