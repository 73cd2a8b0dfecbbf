#!/usr/bin/env python
"""Stall-reason / throughput digest of `ncu --page raw --csv` exports:  python scripts/ncu_stalls.py a_raw.csv [b_raw.csv ...]"""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fmaheavy.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active']
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    names, units = rows[hdr], rows[hdr + 1]
    for r in rows[hdr + 2:]:
        rec = dict(zip(names, r)); un = dict(zip(names, units))
        st = {k.replace('smsp__pcsamp_warps_issue_stalled_', ''): float(v.replace(',', '')) for k, v in rec.items()
              if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued') and v not in ('', 'n/a')}
        tot = sum(st.values()) or 1
        print('##', rec['Kernel Name'][:90], '| grid', rec.get('launch__grid_size'), 'block', rec.get('launch__block_size'))
        print('   stalls:', ' | '.join(f'{k} {100 * v / tot:.0f}%' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
        for k in KEYS:
            if k in rec and rec[k] not in ('', 'n/a'):
                print(f'   {k} = {rec[k]} {un.get(k, "")}')
