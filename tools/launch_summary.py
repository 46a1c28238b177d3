#!/usr/bin/env python
"""Per-kernel mean duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
kn, mv = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[kn][:56], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
last = int(sys.argv[2]) if len(sys.argv) > 2 else len(seq)
agg = collections.OrderedDict()
for k, v in seq[-last:]:
    agg.setdefault(k, []).append(v)
for k, v in agg.items():
    print(f"{k:58s} n={len(v):3d} mean={sum(v) / len(v) / 1000:9.2f} us")
