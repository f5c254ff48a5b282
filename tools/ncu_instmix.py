#!/usr/bin/env python
"""Executed-instruction mix of the first kernel in an .ncu-rep, plus the hottest SASS ranges.  usage: ncu_instmix.py rep [elements/32]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
ex = [int(r[idx['Instructions Executed']]) for r in data]
tot = sum(ex)
print('total warp-instr', tot)
c = collections.Counter()
for r, e in zip(data, ex):
    op = [o for o in r[1].strip().split() if not o.startswith('@')][0].split('.')[0]
    c[op] += e
for k, v in c.most_common(24):
    print(f'{k:12s} {v:10d} {100*v/tot:5.1f}%' + (f'  per elem {v/per:.2f}' if per else ''))
# contiguous regions with the same executed count (basic blocks), sorted by total executed
blocks = []
start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or ex[i] != ex[start]:
        blocks.append((ex[start] * (i - start), ex[start], start, i))
        start = i
print('hottest straight-line regions (total executed, per-instr executed, #instr, first..last):')
for t, e, a, b in sorted(blocks, reverse=True)[:14]:
    print(f'  {t:10d} {e:8d} x {b-a:4d}   {data[a][1].strip()[:50]:50s} .. {data[b-1][1].strip()[:40]}')
