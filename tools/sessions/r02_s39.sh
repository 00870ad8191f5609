#!/bin/bash
# Round-2 session 39: GPU suite after the tree-choice change (car290k now through the 8-wide tree), car290k bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02zk_pytest_gpu.txt
timeout 300 python bench.py --workload car290k --steps 3 --warmup 3 --spp-per-step 16 --no-cpu --also '' 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('car290k', round(d['value'], 1), 'Mrays/s e2e', round(d['e2e']['value'], 1), d['stage_ms_per_step'], d['run']['bvh'])" | tee gpurun_out/r02zk_bench_car290k.txt
