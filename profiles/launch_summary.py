"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: python profiles/launch_summary.py <launches.csv> [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
col = {a: i for i, a in enumerate(rows[0])}
agg, tot, n = collections.OrderedDict(), 0.0, 0
for r in rows[1:]:
    if r[col['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = r[col['Kernel Name']]
    m = re.search(r'<(\w+(?:<[^>]*>)?)', name)
    short = m.group(1) if m else name.split('(')[0]
    v = float(r[col['Metric Value']].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}[r[col['Metric Unit']]]
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; n += 1
print(f"{n} launches, {tot:.3f} ms of kernel time")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for k, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{k:40s} {c:3d} {ms:8.4f} ms {100 * ms / tot:5.1f}%")
