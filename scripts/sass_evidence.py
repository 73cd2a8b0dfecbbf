#!/usr/bin/env python
"""SASS evidence per kernel family: counts of the Blackwell-specific mnemonics in libcasmtr_b200.so (cuobjdump -sass).
    python scripts/sass_evidence.py > profiles/TAG_sass.md
Template instances of one kernel are merged (max over instances)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'casmtr_b200', 'lib', 'libcasmtr_b200.so')
PAT = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|SYNCS|FFMA2|LDGSTS|CREDUX|REDUX|HMMA|UTCCP)\b')
txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
fam = collections.defaultdict(lambda: collections.Counter())
cur, cnt = None, None
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        if cur:
            for k, v in cnt.items():
                fam[cur][k] = max(fam[cur][k], v)
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
        cur = re.sub(r'^void ', '', name).split('<')[0].split('(')[0]
        cnt = collections.Counter()
        continue
    m = PAT.search(line)
    if m and cur:
        cnt[m.group(1)] += 1
if cur:
    for k, v in cnt.items():
        fam[cur][k] = max(fam[cur][k], v)
cols = ['UTCHMMA', 'LDTM', 'UTMALDG', 'SYNCS', 'UTCBAR', 'LDGSTS', 'FFMA2', 'CREDUX']
print('# SASS mnemonics per kernel (cuobjdump -sass casmtr_b200/lib/libcasmtr_b200.so, sm_100a; max over template instances)\n')
print('UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load, SYNCS = mbarrier, UTCBAR = tcgen05.commit,')
print('LDGSTS = cp.async, FFMA2 = packed 2 x fp32 FMA, CREDUX = redux.sync.{max,min}.f32\n')
print('| kernel | ' + ' | '.join(cols) + ' |')
print('|---|' + '---|' * len(cols))
for k in sorted(fam):
    if any(fam[k][c] for c in cols):
        print(f'| `{k}` | ' + ' | '.join(str(fam[k][c] or '') for c in cols) + ' |')
