#!/bin/bash
# Round-2 session 36 (final library): smoke, whole GPU suite, both bench arms, launch list, full ncu captures of k_trace / k_logic (bunny90k)
# and of k_trace on orb500k with the ray counts of the captured launches, summarised ON THE BOX (profiles/ncu_summary.json refreshed),
# and a full capture of the device builder's heaviest kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
P="python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 240 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_bench.log 2>&1
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/iter_log_*.txt
cap() {  # name, kernel regex, count, extra env ...
  local name=$1 kern=$2 cnt=$3; shift 3
  env ADAPT_ITER_LOG=gpurun_out/iter_log_$name.txt "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 6 -c $cnt -f -o gpurun_out/prof_$name $P $EXTRA >> gpurun_out/ncu_full.log 2>&1
}
EXTRA="" cap trace k_trace 2 X=1
EXTRA="" cap logic k_logic 1 X=1
EXTRA="--workload orb500k" cap trace_orb_cw8 k_trace 2 X=1
python tools/profile_summary.py r02zh > gpurun_out/profile_summary.log 2>&1
python tools/ncu_hot.py gpurun_out/prof_trace.ncu-rep 30 > gpurun_out/r02zh_hot_trace.txt 2>&1
BUILDERS=sah_device timeout 600 ncu --set full --clock-control none -k regex:"k_sah_bin|k_sah_scatter|k_sah_split|k_cw8_emit" -s 24 -c 8 -f -o gpurun_out/prof_builder python tools/bvh_build_bench.py orb500k > /dev/null 2>&1
python tools/ncu_extract.py gpurun_out/prof_builder.ncu-rep > gpurun_out/r02zh_ncu_builder.txt 2>&1
mkdir -p gpurun_out/profiles && cp profiles/r02zh_* profiles/ncu_summary.json gpurun_out/profiles/
rm -f gpurun_out/prof_*.ncu-rep
ls gpurun_out gpurun_out/profiles; du -sh gpurun_out
