"""Top stall sites of a kernel in an .ncu-rep (SASS view): python tools/ncu_hot.py gpurun_out/prof_logic.ncu-rep [N]
Prints the N instructions with the most warp-stall samples, with the dominant stall reason and the preceding few
SASS lines for context (run in the dev container, no GPU needed)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
txt = '"Kernel Name"' + blocks[1]
lines = txt.split('\n')
rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
hdr = rows[0]; rows = [r for r in rows[1:] if len(r) == len(hdr)]
iS = hdr.index('Warp Stall Sampling (All Samples)'); iSrc = hdr.index('Source')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS] or 0) for r in rows)
print('kernel:', lines[0][:120]); print('total samples', tot, 'instructions', len(rows))
agg = {h: sum(int(r[i] or 0) for r in rows) for i, h in stall_cols}
print('by reason:', ', '.join(f'{h[6:]}={v / tot:.1%}' for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
order = sorted(range(len(rows)), key=lambda k: -int(rows[k][iS] or 0))[:N]
for k in sorted(order):
    r = rows[k]
    top = max(stall_cols, key=lambda c: int(r[c[0]] or 0))
    print(f'{int(r[iS]) / tot:6.2%} idx {k:5d} {top[1][6:]:14s} {r[iSrc].strip()[:90]}')
