#!/bin/bash
# Round-2 session 5: the compressed 8-wide BVH (ADAPT_TRACE_MODE=3) -- parity, then A/B against the binary tree; and the intersect-stage failure of session 4
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED OR HUNG - aborting session"; exit 1; fi
echo "== intersect test alone"
timeout 300 python -m pytest "tests/test_gpu_parity.py::test_intersect_stage_parity_small" -q -x --timeout 120 2>&1 | tail -4
echo "== test_gpu_parity.py alone"
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 2>&1 | tail -6
echo "== mode 3: all GPU tests"
ADAPT_TRACE_MODE=3 timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_cw8.log
rm -f gpurun_out/ab.txt
L1="ADAPT_LANES=1"
for W in "" "--workload orb500k --spp-per-step 16" "--workload balls-mono --width 1024 --spp-per-step 16"; do
  bash tools/ab.sh "$W" "$L1" "$L1 ADAPT_TRACE_MODE=3" "ADAPT_TRACE_MODE=3" \
     "$L1 ADAPT_TRACE_MODE=3 ADAPT_B200_LIB=$PWD/adapt_b200/lib/cw8s1/libadapt_b200.so" "$L1 ADAPT_TRACE_MODE=3 ADAPT_B200_LIB=$PWD/adapt_b200/lib/cw8s3/libadapt_b200.so" \
     "$L1 ADAPT_TRACE_MODE=3 ADAPT_B200_LIB=$PWD/adapt_b200/lib/cw8b7/libadapt_b200.so" "$L1 ADAPT_TRACE_MODE=3 ADAPT_B200_LIB=$PWD/adapt_b200/lib/cw8b6/libadapt_b200.so" \
     "$L1 ADAPT_TRACE_MODE=3 ADAPT_LEAF_T=12" "$L1 ADAPT_TRACE_MODE=3 ADAPT_LEAF_T=4" "$L1 ADAPT_TRACE_MODE=3 ADAPT_REFILL=8" \
     "$L1 ADAPT_B200_LIB=$PWD/adapt_b200/lib/nobulk/libadapt_b200.so" "$L1 ADAPT_TRACE_MODE=3 ADAPT_B200_LIB=$PWD/adapt_b200/lib/nobulk/libadapt_b200.so"
done
