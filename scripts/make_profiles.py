#!/usr/bin/env python
"""Turn one scripts/gpu_round.sh pass (gpurun_out/TAG_*) into the committed evidence under profiles/:
    python scripts/make_profiles.py r01z
  profiles/TAG_ncu_summary.md   one table row + a stall / pipe digest per captured kernel (ncu --set full, --clock-control none)
  profiles/TAG_bench.json, TAG_bench_ref.json, TAG_launches.csv, TAG_gpu.txt   copied as measured
  profiles/traffic.json         DRAM bytes (read + write) per launch and kernel kind  -> bench.py roofline.traffic
  profiles/issue.json           issue-slot utilisation, L2->SM bytes, warp instructions per kernel kind -> bench.py breakdown[]
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1]
OUT, PROF = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
KIND = [('transpose_vec', 'layout'), ('qtatt_coarse', 'qt_coarse'), ('quad_cta', 'qt_fine_mid'), ('quad_attention_kernel', 'qt_fine_last'),
        ('cascade_att_tile', 'cascade_att'), ('quad_attention_list', 'cascade_fallback'), ('cascade_match_tile', 'cascade_match'),
        ('cascade_match_cell', 'cascade_match_fallback'), ('coarse_rowstats', 'coarse_match'), ('pool2_tokens', 'layout'), ('coarse_prep', 'coarse_prep'),
        ('fine_window_gather', 'fine_window_gather'), ('fine_match', 'fine_match'), ('extract_mask', 'extract'),
        ('relative_pe_kernel', 'relative_pe'), ('score5d_bwd', 'score5d_bwd'), ('value_agg_bwd', 'value_agg_bwd'), ('score3d_bwd', 'score3d_bwd')]


def num(x):
    try:
        return float(x.replace(',', ''))
    except (ValueError, AttributeError):
        return None


UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}     # -> bytes / microseconds


def rows_of(path):
    """ncu's raw page scales every column to a unit of its choice (second header row): bring bytes and times back to base units."""
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        rec = {}
        for n, u, v in zip(names, units, r):
            x = num(v)
            rec[n] = x * UNIT[u] if (x is not None and u in UNIT) else (x if x is not None else v)
        out.append(rec)
    return out


raws = [os.path.join(OUT, f'{TAG}_{r}_raw.csv') for r in ('qtatt', 'cascade', 'match', 'widen', 'w2_relative_pe_kernel', 'w2_score5d_bwd_kernel', 'w2_value_agg_bwd_kernel',
                                                             'w2_score3d_bwd_kernel')]
raws = [p for p in raws if os.path.exists(p)]
buf = io.StringIO()
buf.write(f'# {TAG} -- B200, `ncu --set full --clock-control none --import-source on`, one launch per kernel\n'
          f'(`bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline`; scripts/gpu_round.sh, scripts/make_profiles.py).\n'
          'Times under ncu are cold-cache and serialised: use them for shares and per-kernel counters, not as bench values.\n\n')
buf.write(subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), *raws], capture_output=True, text=True).stdout)
buf.write('\n```\n')
buf.write(subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_stalls.py'), *raws], capture_output=True, text=True).stdout)
buf.write('```\n')
open(os.path.join(PROF, f'{TAG}_ncu_summary.md'), 'w').write(buf.getvalue())

traffic, issue = {}, {}
for p in raws:
    for r in rows_of(p):
        kind = next((k for pat, k in KIND if pat in r['Kernel Name']), None)
        if kind is None:
            continue
        dram = (r.get('dram__bytes_read.sum') or 0) + (r.get('dram__bytes_write.sum') or 0)
        rec = {'issue_active_pct': round(r.get('smsp__issue_active.avg.pct_of_peak_sustained_active') or 0, 1),
               'l2_bytes': int((r.get('lts__t_sectors.sum') or 0) * 32), 'warp_instructions': int(r.get('smsp__inst_executed.sum') or 0),
               'ncu_us': round(r.get('gpu__time_duration.sum') or 0, 2)}
        if kind == 'layout' and kind in traffic:       # two layout launches are captured (QTAttB call, cascade call): keep the mean
            traffic[kind] = int((traffic[kind] + dram) / 2)
            continue
        if kind in traffic:
            continue
        traffic[kind], issue[kind] = int(dram), rec
json.dump(traffic, open(os.path.join(PROF, 'traffic.json'), 'w'), indent=1)
json.dump(issue, open(os.path.join(PROF, 'issue.json'), 'w'), indent=1)
for suffix in ('bench.json', 'bench_ref.json', 'launches.csv', 'gpu.txt'):
    src = os.path.join(OUT, f'{TAG}_{suffix}')
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, f'{TAG}_{suffix}'))
print(json.dumps(issue, indent=1))
