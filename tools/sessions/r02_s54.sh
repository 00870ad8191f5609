#!/bin/bash
# Round-2 session 54: the one-group k_logic compiled for 10 / 12 resident blocks of 128 threads (48 / 40 registers) so that three / four of
# them fit beside the other lane's six trace blocks
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=$PWD/adapt_b200/lib
bash tools/ab.sh "--spp-per-step 256" "ADAPT_B200_LIB=$L/lmb5.so" "ADAPT_B200_LIB=$L/lmb6.so"
cp gpurun_out/ab.txt gpurun_out/r03o_ab_logic_regs.txt
