#!/bin/bash
# Tuning sweep on the GPU box: each line = one bench run with one knob changed. Output: gpurun_out/sweep.txt
mkdir -p gpurun_out
OUT=gpurun_out/sweep.txt
: > $OUT
run() {   # label, env..., -- extra args
  label=$1; shift
  envs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done; shift
  res=$(env "${envs[@]}" timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --spp-per-step 16 "$@" 2>/dev/null | tail -1)
  python - "$label" "$res" >> $OUT <<'PY'
import json, sys
label, res = sys.argv[1], sys.argv[2]
try:
    j = json.loads(res)
    s = j["stage_ms_per_step"]
    print(f"{label:34s} {j['value']:8.1f} Mrays/s  spp/s {j['spp_per_s']:7.1f}  ms/step {j['ms_per_step']:7.2f}  logic {s['logic']:6.2f} shadow {s['shadow']:6.2f} closest {s['closest']:6.2f}  e2e {j['e2e']['value']:8.1f}")
except Exception as ex:
    print(f"{label:34s} FAILED {ex} {res[:200]}")
PY
  tail -1 $OUT
}
L=adapt_b200/lib
W=${WORKLOAD:-bunny90k}
run "base(mode1)" -- --workload $W
run "trace_mode0" ADAPT_TRACE_MODE=0 -- --workload $W
run "blocks_per_sm=4" ADAPT_TRACE_BLOCKS_PER_SM=4 -- --workload $W
run "blocks_per_sm=9" ADAPT_TRACE_BLOCKS_PER_SM=9 -- --workload $W
run "pool=1M" ADAPT_POOL=1048576 -- --workload $W
run "pool=4M" ADAPT_POOL=4194304 -- --workload $W
run "leaf=2" ADAPT_BVH_MAX_LEAF=2 -- --workload $W
run "leaf=8" ADAPT_BVH_MAX_LEAF=8 -- --workload $W
run "logic lb3" ADAPT_B200_LIB=$L/libadapt_b200_lb3.so -- --workload $W
run "logic 128x4" ADAPT_B200_LIB=$L/libadapt_b200_lb128x4.so -- --workload $W
run "logic 128x5" ADAPT_B200_LIB=$L/libadapt_b200_lb128x5.so -- --workload $W
if [ "$1" == "big" ]; then
run "orb500k base" -- --workload orb500k
run "orb500k mode0" ADAPT_TRACE_MODE=0 -- --workload orb500k
fi
