#!/bin/bash
# Round-2 session 3: (a) overlap probe -- two handles on two streams; (b) coherent-ray rate (primary rays only)
mkdir -p gpurun_out
for b in 16 6 5 4; do
  ADAPT_TRACE_BLOCKS_PER_SM=$b timeout 300 python tools/overlap_probe.py bunny90k 32 2>&1 | grep handles
done | tee gpurun_out/overlap_probe.txt
ADAPT_TRACE_BLOCKS_PER_SM=5 timeout 300 python tools/overlap_probe.py orb500k 16 2>&1 | grep handles | tee -a gpurun_out/overlap_probe.txt
rm -f gpurun_out/ab.txt
bash tools/ab.sh "--max-bounce 1" 
bash tools/ab.sh "--max-bounce 2"
bash tools/ab.sh "--max-bounce 4"
