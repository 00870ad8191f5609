#!/bin/bash
# Round-2 session 35: the 8-wide trace kernel at 8 (default) / 9 / 10 resident blocks per SM
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=/root/repo/adapt_b200/lib
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_B200_LIB=$L/v_cw9.so ADAPT_B200_LIB=$L/v_cw10.so
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" ADAPT_B200_LIB=$L/v_cw9.so ADAPT_B200_LIB=$L/v_cw10.so
cp gpurun_out/ab.txt gpurun_out/r03f_ab_cw8_blocks.txt
