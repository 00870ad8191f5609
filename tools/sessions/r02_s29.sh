#!/bin/bash
# Round-2 session 29: k_logic index arithmetic without integer divisions (A/B against the previous library), parity tests
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=/root/repo/adapt_b200/lib
bash tools/ab.sh "" ADAPT_B200_LIB=$L/v_prev.so DEFAULT=2 ADAPT_B200_LIB=$L/v_prev.so
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_B200_LIB=$L/v_prev.so
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" ADAPT_B200_LIB=$L/v_prev.so
cp gpurun_out/ab.txt gpurun_out/r03a_ab_index_arith.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_golden.py tests/test_gpu_vpt.py -q -m gpu --timeout 300 2>&1 | tail -3
