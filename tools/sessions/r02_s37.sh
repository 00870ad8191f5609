#!/bin/bash
# Round-2 session 37: where k_logic_vpt's time goes (fog scene)
mkdir -p gpurun_out
P="python bench.py --integrator vpt --workload cbox --width 1024 --height 1024 --steps 1 --warmup 1 --no-cpu --spp-per-step 8 --also ''"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_logic_vpt -s 8 -c 1 -f -o gpurun_out/prof_logic_vpt $P > gpurun_out/ncu_full.log 2>&1
python tools/ncu_extract.py gpurun_out/prof_logic_vpt.ncu-rep > gpurun_out/r02zi_ncu_logic_vpt.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_logic_vpt.ncu-rep 0 0.008 > gpurun_out/r02zi_logic_vpt_lines.txt 2>&1
python tools/ncu_hot.py gpurun_out/prof_logic_vpt.ncu-rep 30 > gpurun_out/r02zi_hot_logic_vpt.txt 2>&1
rm -f gpurun_out/prof_logic_vpt.ncu-rep
cat gpurun_out/r02zi_ncu_logic_vpt.txt | head -34
