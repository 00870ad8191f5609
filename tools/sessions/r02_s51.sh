#!/bin/bash
# Round-2 session 51: k_logic in 128-thread blocks (default now; k_classify and k_logic_vpt stay at 256) against the 256-thread build,
# then the GPU suite and the default bench line with it
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=$PWD/adapt_b200/lib
bash tools/ab.sh "--spp-per-step 256" "ADAPT_B200_LIB=$L/lb256.so" "ADAPT_TRACE_BLOCKS_2LANES=5"
bash tools/ab.sh "--workload orb500k --spp-per-step 256" "ADAPT_B200_LIB=$L/lb256.so" "ADAPT_TRACE_BLOCKS_2LANES=5"
bash tools/ab.sh "--workload car290k --spp-per-step 32" "ADAPT_B200_LIB=$L/lb256.so"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 256" "ADAPT_B200_LIB=$L/lb256.so"
cp gpurun_out/ab.txt gpurun_out/r03l_ab_logic_block.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/r03l_pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/r03l_bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/r03l_bench.json; tail -2 gpurun_out/bench.err
