#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  launch_table.py file.csv [top]"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
        continue
    unit = row['Metric Unit']
    us = v / 1000 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000)
    short = re.split(r'\(', row['Kernel Name'].replace('void ', '').replace('<unnamed>::', '').replace('(anonymous namespace)::', ''))[0][:64]
    agg[short][0] += 1
    agg[short][1] += us
tot = sum(v[1] for v in agg.values())
print(f'{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time')
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{k:<66s} n={c:5d}  {t / 1e3:8.2f} ms ({100 * t / tot:4.1f}%)  avg {t / c:7.2f} us')
