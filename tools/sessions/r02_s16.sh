#!/bin/bash
# Round-2 session 16: where k_logic's time goes (source-level stall sites + instruction mix), bunny90k and orb500k
mkdir -p gpurun_out
P="python bench.py --steps 1 --warmup 1 --no-cpu --spp-per-step 8"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_logic -s 6 -c 1 -f -o gpurun_out/prof_logic $P > gpurun_out/ncu_full.log 2>&1
python tools/ncu_hot.py gpurun_out/prof_logic.ncu-rep 60 > gpurun_out/r02p_hot_logic.txt 2>&1
python tools/ncu_extract.py gpurun_out/prof_logic.ncu-rep > gpurun_out/r02p_ncu_logic.txt 2>&1
ncu -i gpurun_out/prof_logic.ncu-rep --page source --csv > gpurun_out/r02p_logic_source.csv 2>/dev/null
python - <<'PY' > gpurun_out/r02p_logic_opmix.txt 2>&1
import csv, io, collections
txt = open('gpurun_out/r02p_logic_source.csv').read()
rows = list(csv.reader(io.StringIO(txt)))
hdr = [r for r in rows if 'Source' in r and 'Instructions Executed' in r][0]
iS, iE, iT = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
mix = collections.Counter(); thr = collections.Counter()
for r in rows:
    if len(r) != len(hdr) or r is hdr: continue
    try: e = int(r[iE]); t = int(r[iT])
    except ValueError: continue
    op = r[iS].strip().split()[0] if r[iS].strip() else '?'
    if op.startswith('@'): op = r[iS].strip().split()[1]
    op = op.split('.')[0]
    mix[op] += e; thr[op] += t
tot = sum(mix.values())
print('warp instructions executed', tot)
for op, e in mix.most_common(30):
    print(f'{op:12s} {e:12d} {e / tot:6.1%}  lanes/instr {thr[op] / max(e, 1):5.1f}')
PY
gzip -f gpurun_out/r02p_logic_source.csv
rm -f gpurun_out/prof_logic.ncu-rep
head -50 gpurun_out/r02p_hot_logic.txt; cat gpurun_out/r02p_logic_opmix.txt
