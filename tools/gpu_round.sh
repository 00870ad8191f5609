#!/bin/bash
# One GPU-box session: smoke, tests, bench (+ A/B variants), launch list, full ncu captures of the two kernels.
#   bash tools/gpu_round.sh [ncu] [ab] [big] [vpt]
#     vpt       bring-up of the volumetric kernels (tools/vpt_round.sh)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi
timeout 400 python -m pytest tests -q -m gpu -x --timeout 90 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 240 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for arg in "$@"; do
case $arg in
ab)
  for v in "ADAPT_REFILL=8" "ADAPT_REFILL=12" "ADAPT_LEAF_T=6" "ADAPT_LEAF_T=10" "ADAPT_TRACE_BLOCKS_PER_SM=8" "ADAPT_TRACE_BLOCKS_PER_SM=10" "ADAPT_BVH_MAX_LEAF=3" "ADAPT_POOL=8388608"; do
    echo "== $v"; env $v timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'Mrays/s', round(d['ms_per_step'], 2), 'ms/step', {k: round(v, 2) for k, v in d['stage_ms_per_step'].items()})"
  done | tee gpurun_out/ab.txt ;;
big)
  timeout 400 python bench.py --workload orb500k --steps 3 --warmup 3 --spp-per-step 16 --cpu-budget 10 > gpurun_out/bench_orb500k.json 2> gpurun_out/bench_orb500k.err; tail -c 2500 gpurun_out/bench_orb500k.json; tail -3 gpurun_out/bench_orb500k.err
  timeout 300 python bench.py --workload balls-mono --width 1024 --steps 3 --warmup 3 --spp-per-step 16 --cpu-budget 8 > gpurun_out/bench_balls.json 2> gpurun_out/bench_balls.err; tail -c 2500 gpurun_out/bench_balls.json; tail -3 gpurun_out/bench_balls.err ;;
vpt)
  bash tools/vpt_round.sh ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_bench.log 2>&1
  rm -f gpurun_out/prof_*.ncu-rep
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace \
      python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 6 -c 1 -f -o gpurun_out/prof_logic \
      python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 >> gpurun_out/ncu_full.log 2>&1 ;;
esac
done
ls -la gpurun_out/
