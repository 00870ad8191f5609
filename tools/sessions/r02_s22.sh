#!/bin/bash
# Round-2 session 22: device SAH builder, second version (vertex means, SAH-decided small leaves, bins fetched in one round trip)
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
timeout 900 python -m pytest tests/test_gpu_lbvh.py -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02u_pytest_gpu_lbvh.txt
timeout 300 python tools/bvh_build_bench.py bunny90k orb500k car290k 2>&1 | tee gpurun_out/r02u_bvh_build.txt
BUILDERS=sah_device timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/build_launches.csv python tools/bvh_build_bench.py orb500k > /dev/null 2>&1
python - <<'PY' | tee gpurun_out/r02u_build_launch_summary.txt
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/build_launches.csv')) if len(r) > 10]
hdr = rows[0]; iK = hdr.index('Kernel Name'); iV = hdr.index('Metric Value'); iU = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iV].replace(',', '')); v = v / 1000 if r[iU] == 'ns' else v
    k = r[iK][:60]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('orb500k, 5 builds with the device SAH builder under ncu (serialised launches): total', round(tot / 1000, 2), 'ms')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print(f'{k:62s} launches {n:5d}  total {t / 1000:8.3f} ms  avg {t / n:8.1f} us')
PY
export ADAPT_TRACE_MODE=1
bash tools/ab.sh "" ADAPT_BVH_BUILDER=1 ADAPT_BVH_BUILDER=2
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_BVH_BUILDER=2
bash tools/ab.sh "--workload car290k --spp-per-step 4" ADAPT_BVH_BUILDER=2
cp gpurun_out/ab.txt gpurun_out/r02u_ab_device_sah.txt
