#!/bin/bash
# Round-2 session 20: k_logic per source line (bunny90k: the one-group kernel; orb500k: the class-list launches)
mkdir -p gpurun_out
P="python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 --also ''"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 6 -c 1 -f -o gpurun_out/prof_logic $P > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_logic|k_classify" -s 30 -c 5 -f -o gpurun_out/prof_logic_orb $P --workload orb500k > gpurun_out/ncu_full_orb.log 2>&1
ls -la gpurun_out
