#!/bin/bash
# Round-2 session 11: k_logic variants (early loads, launch bounds) + the staged-top-of-tree variant under ncu for the N1 table
mkdir -p gpurun_out
rm -f gpurun_out/ab.txt
V() { echo "ADAPT_B200_LIB=$PWD/adapt_b200/lib/$1/libadapt_b200.so"; }
for name in early lb5; do env $(V $name) timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 90 2>&1 | tail -1; done
bash tools/ab.sh "" "$(V early)" "$(V lb5)" "$(V lb6)" "$(V earlylb5)"
bash tools/ab.sh "--workload orb500k --spp-per-step 16" "$(V early)" "$(V lb5)" "$(V earlylb5)"
bash tools/ab.sh "--workload balls-mono --width 1024 --spp-per-step 16" "$(V early)" "$(V lb5)" "$(V earlylb5)"
P="python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8"
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/iter_log_*.txt
env ADAPT_ITER_LOG=gpurun_out/iter_log_trace_top256.txt $(V top256) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 2 -f -o gpurun_out/prof_trace_top256 $P > gpurun_out/ncu_full.log 2>&1
python tools/ncu_extract.py gpurun_out/prof_trace_top256.ncu-rep > gpurun_out/r02k_ncu_trace_top256.txt 2>&1
ncu -i gpurun_out/prof_trace_top256.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
want = [h for h in hdr if any(k in h for k in ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'smsp__inst_executed.sum', 'gpu__time_duration.sum', 'smsp__sass_inst_executed_op_shared_ld.sum'))]
for r in rows[2:]:
    print({h: r[hdr.index(h)] for h in want})
" > gpurun_out/r02k_top256_sectors.txt 2>&1
rm -f gpurun_out/prof_trace_top256.ncu-rep
cat gpurun_out/r02k_top256_sectors.txt | head -5
