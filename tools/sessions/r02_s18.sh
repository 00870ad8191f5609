#!/bin/bash
# Round-2 session 18: k_trace per-lane state diet (ray origin / direction / uv in shared memory, hit as a leaf-order index) -> more
# resident warps; branch-free child selection.  Variant libraries built by tools/build_variant.py.
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
L=/root/repo/adapt_b200/lib
bash tools/ab.sh "" ADAPT_B200_LIB=$L/v0_base.so ADAPT_B200_LIB=$L/v2_smem.so ADAPT_B200_LIB=$L/v3_smem10.so ADAPT_B200_LIB=$L/v4_smem10bf.so ADAPT_B200_LIB=$L/v5_bf.so ADAPT_B200_LIB=$L/v6_smem64x21.so ADAPT_B200_LIB=$L/v7_smem12.so
export ADAPT_TRACE_MODE=1
bash tools/ab.sh "--workload orb500k --spp-per-step 16" ADAPT_B200_LIB=$L/v0_base.so ADAPT_B200_LIB=$L/v3_smem10.so ADAPT_B200_LIB=$L/v4_smem10bf.so ADAPT_B200_LIB=$L/v6_smem64x21.so
cp gpurun_out/ab.txt gpurun_out/r02r_ab_trace_state.txt
