#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / total / share."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
skip = sys.argv[2:] or []
agg = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    k = r[ki]
    if any(s in k for s in skip):
        continue
    k = k.replace('qtet::<unnamed>::', '').split('(')[0][:48]
    agg.setdefault(k, []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':50s} {'n':>4s} {'mean_us':>10s} {'total_us':>11s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:50s} {len(v):4d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/1e3:11.1f} {100*sum(v)/tot:5.1f}%")
