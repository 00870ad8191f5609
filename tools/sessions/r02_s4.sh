#!/bin/bash
# Round-2 session 4: lanes + bulk-tile k_logic: tests, A/B, the full bench line of both arms
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 300 -s 2>&1 | grep -E "passed|failed|error|rel L2|Error|assert" | tail -30 | tee gpurun_out/pytest_gpu.log
rm -f gpurun_out/ab.txt
N="ADAPT_B200_LIB=$PWD/adapt_b200/lib/nobulk/libadapt_b200.so"
bash tools/ab.sh "" "ADAPT_LANES=1" "$N" "$N ADAPT_LANES=1" "ADAPT_TRACE_BLOCKS_PER_SM=8" "ADAPT_TRACE_BLOCKS_PER_SM=7" "ADAPT_TRACE_BLOCKS_PER_SM=6" "ADAPT_POOL=4194304"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" "ADAPT_LANES=1" "$N" "ADAPT_TRACE_BLOCKS_PER_SM=7" "ADAPT_POOL=4194304"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" "ADAPT_LANES=1" "$N" "ADAPT_TRACE_BLOCKS_PER_SM=7"
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 5000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
