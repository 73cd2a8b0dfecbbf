#!/usr/bin/env python
"""Per CUDA source line: executed warp-instructions and stall samples, from
   ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > f.csv
   python scripts/ncu_line_hist.py f.csv <function-substring> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = {}
cur_file = cur_fn = None
hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1]; continue
    if r[0] == 'Function Name': cur_fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or cur_fn is None or want not in cur_fn: continue
    if r[2] != '-':      # SASS row
        continue
    try:
        ln = int(r[0]); ex = int(r[hdr.index('Instructions Executed')] or 0); st = int(r[hdr.index('Warp Stall Sampling (All Samples)')] or 0)
    except ValueError:
        continue
    key = (cur_file.split('/')[-1], ln, r[1].strip()[:110])
    a = out.setdefault(key, [0, 0]); a[0] += ex; a[1] += st
tot = sum(v[0] for v in out.values()); ts = sum(v[1] for v in out.values())
print(f'{want}: total exec {tot}, stall samples {ts}')
for (f, ln, src), (ex, st) in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100*ex/max(tot,1):5.1f}% exec {100*st/max(ts,1):5.1f}% stall  {f}:{ln}  {src}')
