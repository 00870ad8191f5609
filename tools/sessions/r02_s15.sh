#!/bin/bash
# Round-2 session 15: pool layout (RNG state in spare words; records for class-list scenes): GPU suite in both layouts, A/B
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
ADAPT_POOL_AOS=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vpt.py tests/test_reference_golden.py -q -m gpu --timeout 300 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_aos1.log
ADAPT_POOL_AOS=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_golden.py -q -m gpu --timeout 300 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_aos0.log
rm -f gpurun_out/ab.txt
bash tools/ab.sh "" "ADAPT_POOL_AOS=1"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" "ADAPT_POOL_AOS=0"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" "ADAPT_POOL_AOS=0"
bash tools/ab.sh "--workload car290k --spp-per-step 4" "ADAPT_POOL_AOS=0"
