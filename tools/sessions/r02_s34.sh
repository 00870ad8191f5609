#!/bin/bash
# Round-2 session 34: k_transmit_vpt at 4 / 6 / 7 / 8 / 9 resident blocks per SM
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=/root/repo/adapt_b200/lib
bash tools/ab.sh "--integrator vpt --workload cbox --width 1024 --height 1024 --spp-per-step 16" ADAPT_B200_LIB=$L/v_tr7.so ADAPT_B200_LIB=$L/v_tr8.so ADAPT_B200_LIB=$L/v_tr9.so
bash tools/ab.sh "--integrator vpt --workload media --width 1024 --height 1024 --spp-per-step 16" ADAPT_B200_LIB=$L/v_tr7.so ADAPT_B200_LIB=$L/v_tr8.so ADAPT_B200_LIB=$L/v_tr9.so
cat gpurun_out/ab.txt >> gpurun_out/r03e_ab_vpt_transmit_blocks.txt
