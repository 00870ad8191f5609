#!/bin/bash
# Round-end GPU session, most important first (the box time left may cut the tail): smoke, GPU tests, bench line, ncu launch list,
# full ncu capture of k_trace, the two other workloads, full capture of k_logic, a few knob A/Bs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi
timeout 200 python -m pytest tests -q -m gpu -x --timeout 90 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_bench.log 2>&1
rm -f gpurun_out/prof_*.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_full.log 2>&1
timeout 200 python bench.py --workload orb500k --steps 3 --warmup 3 --spp-per-step 16 --cpu-budget 10 > gpurun_out/bench_orb500k.json 2> gpurun_out/bench_orb500k.err; tail -c 1200 gpurun_out/bench_orb500k.json
timeout 120 python bench.py --workload balls-mono --width 1024 --steps 3 --warmup 3 --spp-per-step 16 --cpu-budget 8 > gpurun_out/bench_balls.json 2> gpurun_out/bench_balls.err; tail -c 1200 gpurun_out/bench_balls.json
rm -f gpurun_out/ab.txt
bash tools/ab.sh "" ADAPT_LEAF_T=8 ADAPT_LEAF_T=10 ADAPT_REFILL=12 ADAPT_REFILL=20 "ADAPT_B200_LIB=$PWD/adapt_b200/lib/ct4/libadapt_b200.so"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 6 -c 1 -f -o gpurun_out/prof_logic \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | head -40
