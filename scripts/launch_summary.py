"""Compact an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_summary.py in.csv out.txt"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg, order = OrderedDict(), []
tot = 0.0
for r in rows:
    name = re.sub(r'\(.*', '', r[4]).replace('void ', '').strip()[:90]
    ns = float(r[-1].replace(',', ''))
    tot += ns
    a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
    a[0] += 1
    a[1] += ns
with open(sys.argv[2], 'w') as f:
    f.write(f'# launches: {len(rows)}  total device time: {tot/1e6:.3f} ms  (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n')
    f.write(f'# {"kernel":90s} {"count":>6s} {"total ms":>12s} {"share %":>8s}  block / grid of first launch\n')
    for name, (n, ns, blk, grd) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'{name:92s} {n:6d} {ns/1e6:12.3f} {100*ns/tot:8.2f}  {blk} {grd}\n')
print(open(sys.argv[2]).read()[:3000])
