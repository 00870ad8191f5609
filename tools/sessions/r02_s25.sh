#!/bin/bash
# Round-2 session 25: k_logic_vpt per material set and at 3 / 4 blocks per SM
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=/root/repo/adapt_b200/lib
bash tools/ab.sh "--integrator vpt --workload cbox --width 1024 --height 1024 --spp-per-step 16" ADAPT_VPT_SPECIALISE=0 ADAPT_B200_LIB=$L/v_vpt3.so ADAPT_B200_LIB=$L/v_vpt4.so
bash tools/ab.sh "--integrator vpt --workload media --width 1024 --height 1024 --spp-per-step 16" ADAPT_VPT_SPECIALISE=0 ADAPT_B200_LIB=$L/v_vpt3.so ADAPT_B200_LIB=$L/v_vpt4.so
cp gpurun_out/ab.txt gpurun_out/r02x_ab_vpt_logic.txt
timeout 600 python -m pytest tests/test_gpu_vpt.py -q -m gpu --timeout 300 2>&1 | tail -3
ADAPT_B200_LIB=$L/v_vpt4.so timeout 600 python -m pytest tests/test_gpu_vpt.py -q -m gpu --timeout 300 2>&1 | tail -3
