#!/bin/bash
# Second round-end session (compile-time node steps as the default): GPU tests, bench line, ncu launch list, leaf-threshold A/B on the
# two other workloads, full capture of k_trace.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x --timeout 90 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_bench.log 2>&1
rm -f gpurun_out/ab.txt
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_LEAF_T=8
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" ADAPT_LEAF_T=8
rm -f gpurun_out/prof_trace.ncu-rep
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_full.log 2>&1
ls gpurun_out | head -30
