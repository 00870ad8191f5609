"""Prints the key metrics of an .ncu-rep capture (run here, no GPU needed): python tools/ncu_extract.py gpurun_out/prof_closest.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum', 'smsp__sass_inst_executed_op_global_ld.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('---- kernel', r[hdr.index('Kernel Name')][:50], 'id', r[0])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f'  {w:75s} {r[i]:>18s} {units[i]}')
    st = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in stall), reverse=True)[:8]
    for v, h in st:
        print(f'  stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:40s} {v:10.3f}')
