"""Per CUDA source line: share of warp instructions, lanes per instruction and share of stall samples of one kernel in an .ncu-rep.
python tools/ncu_lines.py report.ncu-rep [kernel-index] [min-share]   (runs in the dev container, no GPU needed)"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; kernels = []; agg = None
for r in rows:
    if len(r) == 2 and r[0] == 'Function Name':
        if agg is None or fname != r[1]:
            fname = r[1]; agg = collections.OrderedDict(); kernels.append((fname, agg))
        continue
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r; iE = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iS = hdr.index('Warp Stall Sampling (All Samples)'); continue
    if hdr and len(r) == len(hdr) and r[0] != '':
        try: l = int(r[0]); e = int(r[iE] or 0); t = int(r[iT] or 0); s = int(r[iS] or 0)
        except ValueError: continue
        if e == 0 and s == 0: continue
        k = (cur, l); o = agg.get(k, (0, 0, 0, ''))
        agg[k] = (o[0] + e, o[1] + t, o[2] + s, r[1].strip()[:105])
print('kernels:', [k[0][:60] for k in kernels])
fname, agg = kernels[kidx]
tot = sum(v[0] for v in agg.values()); tots = max(1, sum(v[2] for v in agg.values()))
print(fname[:120]); print('total warp inst', tot, 'stall samples', tots)
for (f, l), (e, t, s, src) in agg.items():
    if e / tot > thr or s / tots > thr:
        print(f'{f:14s}{l:4d} {e / tot:6.2%} lanes {t / max(e, 1):5.1f} stall {s / tots:6.2%}  {src}')
