#!/usr/bin/env python
"""Per-source-line summary of an ncu report:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv ; ncu_lines.py f.csv [kernel#]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks = []
cur = None
for r in rows:
    if r and r[0] == 'Function Name':
        cur = dict(name=r[1], rows=[]); blocks.append(cur); continue
    if cur is None:
        continue
    if r and r[0] == 'Line No':
        cur['hdr'] = r; continue
    if 'hdr' in cur and r and r[0] not in ('', 'File Path') and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
b = blocks[which]
h = b['hdr']
ismp, iex = h.index('# Samples'), h.index('Instructions Executed')
def f(v):
    try:
        return float(v)
    except ValueError:
        return 0.
tot, tots = sum(f(r[iex]) for r in b['rows']), sum(f(r[ismp]) for r in b['rows'])
print(b['name'], f'| {len(blocks)} kernels | inst {tot:.3g} | samples {tots:.0f}')
for r in sorted(b['rows'], key=lambda r: -f(r[ismp]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print(f'{f(r[ismp]) / tots * 100:5.1f}% smp {f(r[iex]) / tot * 100:5.1f}% inst  L{r[0]:>4} {r[1].strip()[:120]}')
