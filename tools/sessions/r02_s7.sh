#!/bin/bash
# Round-2 session 7: the shipped defaults -- full bench lines of both arms, batch-size scan, ncu captures with per-launch ray counts,
# the staged-top-of-tree variant under ncu (L1 / L2 sector counts for the N1 table)
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
rm -f gpurun_out/ab.txt
for s in 32 64 256; do bash tools/ab.sh "--spp-per-step $s"; done
bash tools/ab.sh "--workload orb500k --spp-per-step 64"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 64"
bash tools/ab.sh "--workload car290k --spp-per-step 16"
# ---- profiles (never a bench number)
P="python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_bench.log 2>&1
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/iter_log_*.txt
ADAPT_ITER_LOG=gpurun_out/iter_log_trace.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace $P > gpurun_out/ncu_full.log 2>&1
ADAPT_ITER_LOG=gpurun_out/iter_log_logic.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 6 -c 1 -f -o gpurun_out/prof_logic $P >> gpurun_out/ncu_full.log 2>&1
ADAPT_ITER_LOG=gpurun_out/iter_log_trace_orb_cw8.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace_orb_cw8 $P --workload orb500k >> gpurun_out/ncu_full.log 2>&1
ADAPT_TRACE_MODE=1 ADAPT_ITER_LOG=gpurun_out/iter_log_trace_orb_bin.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace_orb_bin $P --workload orb500k >> gpurun_out/ncu_full.log 2>&1
ADAPT_B200_LIB=$PWD/adapt_b200/lib/top256/libadapt_b200.so ADAPT_ITER_LOG=gpurun_out/iter_log_trace_top256.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace_top256 $P >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
