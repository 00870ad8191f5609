#!/bin/bash
# ncu --set full capture of steady-state k_closest / k_shadow launches for the given env (mode etc.)
mkdir -p gpurun_out
for mode in "$@"; do
ADAPT_TRACE_MODE=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_closest -s 6 -c 1 -f -o gpurun_out/prof_closest_m$mode \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_m$mode.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
