#!/bin/bash
# Round-2 session 32: the device SAH builder as the default builder: smoke, whole GPU suite, bench line
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -6 | tee gpurun_out/r03c_pytest_gpu.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r03c_bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r03c_bench.json').read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'bvh', d['run']['bvh'], 'also', d['also']['orb500k']['value'], d['also']['orb500k']['e2e']['value'], d['also']['orb500k'].get('bvh'))
PY
tail -3 gpurun_out/bench.err
