#!/bin/bash
# Shipped defaults (leaf threshold 8): GPU tests + the bench line.
mkdir -p gpurun_out
timeout 120 python -m pytest tests -q -m gpu -x --timeout 90 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
timeout 100 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
