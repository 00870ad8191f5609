"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py launches.csv [title]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; iK = hdr.index('Kernel Name'); iV = hdr.index('Metric Value'); iU = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iV].replace(',', '')); v = v / 1000 if r[iU] == 'ns' else v
    a = agg.setdefault(r[iK][:60], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print((sys.argv[2] + ': ' if len(sys.argv) > 2 else '') + f'total {tot / 1000:.2f} ms')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{k:62s} launches {n:5d}  total {t / 1000:8.3f} ms  avg {t / n:8.1f} us')
