#!/bin/bash
# A/B bench lines: bash tools/ab.sh "<bench args>" VAR=VAL [VAR=VAL ...]   (first run = defaults)
args=$1; shift
# A/B tables are taken at 32 spp per step (the round-1 figures) unless the arguments say otherwise
case "$args" in *--spp-per-step*) ;; *) args="$args --spp-per-step 32" ;; esac
mkdir -p gpurun_out
for v in "DEFAULT=1" "$@"; do
  echo -n "== $args | $v : "
  env $v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --also '' $args 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'Mrays/s', round(d['ms_per_step'], 2), 'ms/step', {k: round(v, 2) for k, v in d['stage_ms_per_step'].items()}, 'e2e', round(d['e2e']['value'], 1), d.get('run', {}).get('bvh'))"
done | tee -a gpurun_out/ab.txt
