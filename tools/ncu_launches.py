#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total, average, share.
usage: ncu_launches.py launches.csv [markdown]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
name_i, val_i, unit_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[val_i].replace(",", ""))
    u = r[unit_i]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(r[name_i].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t / 1e3:.3f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
