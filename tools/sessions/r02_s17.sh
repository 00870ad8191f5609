#!/bin/bash
# Round-2 session 17: where k_trace's instructions go (per source line: node step / leaf phase / scheduler / refill), bunny90k
mkdir -p gpurun_out
P="python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/prof_trace $P > gpurun_out/ncu_full.log 2>&1
python tools/ncu_hot.py gpurun_out/prof_trace.ncu-rep 40 > gpurun_out/r02q_hot_trace.txt 2>&1
python tools/ncu_extract.py gpurun_out/prof_trace.ncu-rep > gpurun_out/r02q_ncu_trace.txt 2>&1
ls -la gpurun_out/
head -30 gpurun_out/r02q_ncu_trace.txt
