#!/bin/bash
# Round-2 session 8: ncu captures of the shipped defaults with per-launch ray counts, summarised ON THE BOX (the reports exceed the 64 MiB
# that come back): launch list, k_trace / k_logic on bunny90k, k_trace on orb500k (8-wide and binary), the staged-top-of-tree variant
mkdir -p gpurun_out
P="python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_bench.log 2>&1
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/iter_log_*.txt
cap() {  # name, kernel regex, count, extra env / args ...
  local name=$1 kern=$2 cnt=$3; shift 3
  env ADAPT_ITER_LOG=gpurun_out/iter_log_$name.txt "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 6 -c $cnt -f -o gpurun_out/prof_$name $P $EXTRA >> gpurun_out/ncu_full.log 2>&1
}
EXTRA="" cap trace k_trace 2 X=1
EXTRA="" cap logic k_logic 1 X=1
EXTRA="--workload orb500k" cap trace_orb_cw8 k_trace 2 X=1
EXTRA="--workload orb500k" cap trace_orb_bin k_trace 2 ADAPT_TRACE_MODE=1
EXTRA="" cap trace_top256 k_trace 2 ADAPT_B200_LIB=$PWD/adapt_b200/lib/top256/libadapt_b200.so
python tools/profile_summary.py r02h > gpurun_out/profile_summary.log 2>&1
python tools/ncu_hot.py gpurun_out/prof_trace.ncu-rep 30 > gpurun_out/r02h_hot_trace.txt 2>&1
python tools/ncu_hot.py gpurun_out/prof_trace_orb_cw8.ncu-rep 30 > gpurun_out/r02h_hot_trace_orb_cw8.txt 2>&1
mkdir -p gpurun_out/profiles && cp profiles/r02h_* profiles/ncu_summary.json gpurun_out/profiles/
rm -f gpurun_out/prof_trace_top256.ncu-rep gpurun_out/prof_trace_orb_bin.ncu-rep gpurun_out/prof_trace_orb_cw8.ncu-rep gpurun_out/prof_logic.ncu-rep
ls -la gpurun_out gpurun_out/profiles; du -sh gpurun_out
