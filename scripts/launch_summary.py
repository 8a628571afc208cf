#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share, and (optionally) one pass in order."""
import collections
import csv
import sys


def main(path, show_pass=None):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    kn, mv, gs = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
    seq = [(r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', ''), float(r[mv].replace(',', '')) / 1e3, r[gs]) for r in data]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, _ in seq:
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f'total {tot / 1e3:.1f} ms over {len(seq)} launches')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f'{v[1] / tot * 100:6.2f}%  n={v[0]:5d}  avg={v[1] / v[0]:9.1f}us  {k[:80]}')
    if show_pass is not None:
        idxs = [i for i, s in enumerate(seq) if s[0].startswith('assemble_tokens')]
        a, b = idxs[show_pass], idxs[show_pass + 1]
        t = 0.
        for i in range(a - 4, b - 4):
            n, d, g = seq[i]
            t += d
            print(f'{i - a:4d} {d:9.1f}us {g:>16s} {n[:60]}')
        print(f'pass total {t / 1e3:.2f} ms')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
