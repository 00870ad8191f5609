#!/bin/bash
# Round-2 session 2: A/B of the compile-time trace variants (prebuilt under adapt_b200/lib/<name>/) + ncu recapture of the defaults.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi
rm -f gpurun_out/ab.txt
VARS=()
for name in "$@"; do
  V="ADAPT_B200_LIB=$PWD/adapt_b200/lib/$name/libadapt_b200.so"
  echo "== parity tests on variant $name"
  env $V timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 90 2>&1 | tail -2
  VARS+=("$V")
done
bash tools/ab.sh "" "${VARS[@]}"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" "${VARS[@]}"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" "${VARS[@]}"
if [ "$NCU" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_bench.log 2>&1
  rm -f gpurun_out/prof_*.ncu-rep
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace \
      python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 > gpurun_out/ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 6 -c 1 -f -o gpurun_out/prof_logic \
      python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8 >> gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out/
