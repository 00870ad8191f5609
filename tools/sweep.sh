#!/bin/bash
# Tuning sweep on the GPU box: each line = one bench run with one knob changed. Output: gpurun_out/sweep.txt
mkdir -p gpurun_out
OUT=gpurun_out/sweep.txt
: > $OUT
run() {   # label, env..., -- extra args
  label=$1; shift
  envs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done; shift
  res=$(env "${envs[@]}" timeout 90 python bench.py --steps 3 --warmup 2 --no-cpu --spp-per-step 16 "$@" 2>/dev/null | tail -1)
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
run "bunny default" -- --workload $W
run "bunny no-cull" ADAPT_CULL_PRIMARY=0 -- --workload $W
run "bunny mode1 (no vote, no popcull)" ADAPT_TRACE_MODE=1 -- --workload $W
if [ "$1" == "big" ]; then
run "orb500k default" -- --workload orb500k
run "orb500k no-cull" ADAPT_CULL_PRIMARY=0 -- --workload orb500k
run "balls-mono 1024 default" -- --workload balls-mono --width 1024 --height 1024
run "car290k default" -- --workload car290k --spp-per-step 4
fi
