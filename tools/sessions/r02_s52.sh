#!/bin/bash
# Round-2 session 52: final library (k_logic 128 threads per block for one-group scenes, 256 for the class-list launches): smoke, GPU suite,
# default bench line, the three other workloads
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/r03m_pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/r03m_bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/r03m_bench.json; tail -2 gpurun_out/bench.err
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 256"
bash tools/ab.sh "--workload car290k --spp-per-step 32"
cp gpurun_out/ab.txt gpurun_out/r03m_ab_other_workloads.txt
