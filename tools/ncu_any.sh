#!/bin/bash
# Full ncu capture of one kernel on one workload: bash tools/ncu_any.sh <kernel-regex> <out-name> <bench args...>
k=$1; o=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_$o \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 4 "$@" > gpurun_out/ncu_$o.log 2>&1
tail -2 gpurun_out/ncu_$o.log | cut -c1-300
