#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv). usage: launch_summary.py file.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, bi, vi, mi = hdr.index('Kernel Name'), hdr.index('Block Size'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg = OrderedDict()
for r in rows[1:]:
    if r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    unit = r[hdr.index('Metric Unit')]
    us = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
    agg.setdefault((r[ki][:70], r[bi]), []).append(us)
tot = sum(sum(v) for v in agg.values())
for (k, b), v in agg.items():
    print(f'{k:70s} block {b:14s} n={len(v):4d} mean={sum(v) / len(v):8.1f} us share={100 * sum(v) / tot:5.1f}%')
