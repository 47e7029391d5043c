"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python tools/launch_summary.py <launches.csv> <out.txt>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[0] == 'ID':
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    agg[d['Kernel Name'][:70]][0] += 1
    agg[d['Kernel Name'][:70]][1] += v
tot = sum(v[1] for v in agg.values())
lines = [f'{"kernel":70s} {"launches":>8s} {"total us":>12s} {"avg us":>10s} {"share":>7s}']
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f'{k:70s} {v[0]:8d} {v[1] / 1e3:12.1f} {v[1] / v[0] / 1e3:10.1f} {v[1] / tot:7.3f}')
open(sys.argv[2], 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
