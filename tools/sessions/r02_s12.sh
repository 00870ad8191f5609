#!/bin/bash
# Round-2 session 12: whole GPU suite on the final library (refit path included), both bench arms
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -s 2>&1 | grep -E "passed|failed|error|refit|Error|assert" | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
