#!/bin/bash
# Round-2 session 49: the library with the two-lane trace grid -- smoke, whole GPU suite, the default bench line, launch list
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/r03j_pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/r03j_bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/r03j_bench.json; tail -2 gpurun_out/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/r03j_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --also '' --spp-per-step 32 > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r03j_launches.csv 2>/dev/null | tee gpurun_out/r03j_launch_summary.txt | tail -12
