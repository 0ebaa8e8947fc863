#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and
the launches of the last proof group in order.  usage: launch_summary.py launches.csv [--group]"""
import csv, re, sys, collections
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = []
for x in csv.DictReader(lines):
    if x.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(x['Metric Value'].replace(',', ''))
    u = x['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    name = re.sub(r'\(.*', '', x['Kernel Name']).replace('void <unnamed>::', '').replace('<unnamed>::', '')
    rows.append((int(x['ID']), name, v, x['Grid Size'], x['Block Size']))
agg = collections.OrderedDict()
tot = sum(r[2] for r in rows)
for r in rows:
    a = agg.setdefault(r[1], [0, 0.0]); a[0] += 1; a[1] += r[2]
print('total %.1f us over %d launches' % (tot, len(rows)))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-58s n=%4d total=%10.1f us avg=%9.1f share=%5.1f%%' % (k[:58], c, t, t / c, 100 * t / tot))
if '--group' in sys.argv:
    idx = [i for i, x in enumerate(rows) if x[1].startswith('extras')]
    s = idx[-1] - 14
    print('--- last proof group, launch order ---')
    for x in rows[s:s + 80]:
        print(x[0], x[1][:50], '%.1f us' % x[2], x[3], x[4])
