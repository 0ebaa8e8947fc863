#!/usr/bin/env python
"""Hot SASS regions of an ncu report: ncu -i X.ncu-rep --page source --csv > src.csv; sass_regions.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = []
for r in rows[2:]:
    try: ie = float(r[ix['Instructions Executed']])
    except Exception: continue
    data.append((r[0], r[1], ie, float(r[ix['Thread Instructions Executed']] or 0), float(r[ix['# Samples']] or 0)))
tot = sum(d[2] for d in data); ts = sum(d[4] for d in data)
print('total warp-instr %.0f, sass instrs %d, samples %.0f' % (tot, len(data), ts))
def op(s):
    p = s.split()
    return p[1] if p[0].startswith('@') and len(p) > 1 else p[0]
prev = None; start = 0; ops = collections.Counter(); out = []; smp = 0; thr = 0
for i, d in enumerate(data + [(None, 'END', -1, 0, 0)]):
    k = round(d[2])
    if k != prev:
        if prev is not None: out.append((data[start][0], prev, i - start, dict(ops.most_common(5)), smp, thr))
        prev = k; start = i; ops = collections.Counter(); smp = 0; thr = 0
    if d[0] is not None:
        ops[op(d[1])] += 1; smp += d[4]; thr += d[3]
for o in out:
    if o[1] * o[2] / tot > 0.01 or o[4] / max(ts, 1) > 0.02:
        print('%s count %8d len %4d instr-share %5.1f%% sample-share %5.1f%% lanes %.1f %s' % (o[0][-6:], o[1], o[2], 100 * o[1] * o[2] / tot, 100 * o[4] / max(ts, 1), o[5] / max(o[1] * o[2], 1), o[3]))
