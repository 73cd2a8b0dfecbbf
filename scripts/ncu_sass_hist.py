#!/usr/bin/env python
"""Opcode histogram (executed warp-instructions, stall samples) from `ncu --page source --csv` of one kernel.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > k.csv ; python scripts/ncu_sass_hist.py k.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
names = rows[hdr]
ci, cs, cx = names.index('Source'), names.index('Warp Stall Sampling (All Samples)'), names.index('Instructions Executed')
ex, st = collections.Counter(), collections.Counter()
tot = 0
for r in rows[hdr + 1:]:
    if len(r) <= max(ci, cs, cx): continue
    toks = r[ci].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0]
    n = int(r[cx] or 0); s = int(r[cs] or 0)
    ex[op] += n; st[op] += s; tot += n
ts = sum(st.values())
print(f'total executed warp-instructions {tot}, stall samples {ts}')
for op, n in ex.most_common(25):
    print(f'{op:10s} exec {n:12d} {100*n/tot:5.1f}%   stall {100*st[op]/max(ts,1):5.1f}%')
