#!/bin/bash
# Round-2 session 6: launch thread (async adapt_render), the cheaper 8-wide step, full GPU suite
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python tools/async_probe.py bunny90k 2>&1 | tail -6 | tee gpurun_out/async_probe.txt
rm -f gpurun_out/ab.txt
S1="ADAPT_B200_LIB=$PWD/adapt_b200/lib/cw8s1/libadapt_b200.so"
M="ADAPT_TRACE_MODE=3"
for W in "" "--workload orb500k --spp-per-step 16" "--workload balls-mono --width 1024 --spp-per-step 16" "--workload car290k --spp-per-step 4"; do
  bash tools/ab.sh "$W" "$M" "$M ADAPT_REFILL=8" "$M ADAPT_LEAF_T=4" "$M ADAPT_REFILL=8 ADAPT_LEAF_T=4" "$M ADAPT_REFILL=12 ADAPT_LEAF_T=6" "$M $S1" "$M $S1 ADAPT_REFILL=8 ADAPT_LEAF_T=4"
done
