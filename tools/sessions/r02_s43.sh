#!/bin/bash
# Round-2 session 43: three and four lanes; the GPU suite with two lanes
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=/root/repo/adapt_b200/lib
export ADAPT_B200_LIB=$L/v_lanes4.so
bash tools/ab.sh "--spp-per-step 256" "ADAPT_LANES=2" "ADAPT_LANES=3" "ADAPT_LANES=4" "ADAPT_LANES=3 ADAPT_TRACE_BLOCKS_PER_SM=6" "ADAPT_LANES=4 ADAPT_POOL=8388608"
bash tools/ab.sh "--workload orb500k --spp-per-step 128" "ADAPT_LANES=2" "ADAPT_LANES=3"
cp gpurun_out/ab.txt gpurun_out/r02zn_ab_lanes_34.txt
unset ADAPT_B200_LIB
ADAPT_LANES=2 timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02zn_pytest_gpu_lanes2.txt
