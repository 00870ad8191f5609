#!/bin/bash
# ncu --set full capture of a steady-state k_logic launch for heavy-shading workloads
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 8 -c 1 -f -o gpurun_out/prof_logic_mono \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 4 --workload balls-mono --width 1024 --height 1024 > gpurun_out/ncu_logic_mono.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 8 -c 1 -f -o gpurun_out/prof_logic_orb \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 4 --workload orb500k > gpurun_out/ncu_logic_orb.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_closest -s 8 -c 1 -f -o gpurun_out/prof_closest_orb \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 4 --workload orb500k > gpurun_out/ncu_closest_orb.log 2>&1
ls -la gpurun_out/*.ncu-rep
