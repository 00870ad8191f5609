#!/bin/bash
# One GPU-box session: tests, bench, launch list, full ncu capture of the dominant kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi
timeout 300 python -m pytest tests -q -m gpu -x --timeout 120 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -8
timeout 240 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$1" == "sweep" ]; then bash tools/sweep.sh $2; fi
if [ "$1" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_closest -s 6 -c 2 -f -o gpurun_out/prof_closest \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 6 -c 1 -f -o gpurun_out/prof_logic \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 >> gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 6 -c 1 -f -o gpurun_out/prof_shadow \
    python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 >> gpurun_out/ncu_full.log 2>&1
timeout 200 python tools/dump_gpu.py | tail -1
ls -la gpurun_out/
fi
