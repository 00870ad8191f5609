#!/bin/bash
# Round-2 session 23: device SAH builder with block-private bins: tests, build times, per-kernel launch times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lbvh.py -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02v_pytest_gpu_lbvh.txt
BUILDERS=sah_device timeout 300 python tools/bvh_build_bench.py bunny90k orb500k car290k 2>&1 | tee gpurun_out/r02v_bvh_build.txt
BUILDERS=sah_device timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/build_launches.csv python tools/bvh_build_bench.py orb500k > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/build_launches.csv "orb500k, 5 builds with the device SAH builder under ncu (serialised launches)" | tee gpurun_out/r02v_build_launch_summary.txt
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/build_launches.csv')) if len(r) > 10]
hdr = rows[0]; iK = hdr.index('Kernel Name'); iV = hdr.index('Metric Value'); iU = hdr.index('Metric Unit')
seq = [(r[iK].split('(')[0].split('::')[-1], float(r[iV].replace(',', '')) / (1000 if r[iU] == 'ns' else 1)) for r in rows[1:]]
# per level of the first build: bin / scatter times
lv = 0; out = []
for k, t in seq:
    if k == 'k_sah_bin': out.append([lv, t, 0.0]); 
    if k == 'k_sah_scatter': out[-1][2] = t; lv += 1
    if k == 'k_fit': break
print('level: bin us / scatter us'); print(' '.join(f'{l}:{b:.0f}/{s:.0f}' for l, b, s in out))
PY
