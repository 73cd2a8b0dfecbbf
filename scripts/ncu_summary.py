#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with the metrics the roofline needs.
    python scripts/ncu_summary.py gpurun_out/x_raw.csv [more.csv ...] > profiles/x_summary.md
"""
import csv
import sys

COLS = [('gpu__time_duration.sum', 'us', 1e-3), ('dram__bytes_read.sum', 'rd_MB', 1e-6), ('dram__bytes_write.sum', 'wr_MB', 1e-6),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 1), ('lts__t_bytes.sum', 'L2_MB', 1e-6),
        ('lts__t_sector_hit_rate.pct', 'L2hit%', 1), ('l1tex__t_bytes.sum', 'L1_MB', 1e-6),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 1), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 1),
        ('launch__registers_per_thread', 'regs', 1), ('launch__grid_size', 'grid', 1), ('launch__block_size', 'block', 1),
        ('smsp__inst_executed.sum', 'inst_M', 1e-6), ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma%', 1),
        ('sm__inst_executed_pipe_lsu.sum', 'lsu_M', 1e-6), ('launch__occupancy_limit_registers', 'lim_reg', 1),
        ('launch__occupancy_limit_shared_mem', 'lim_smem', 1)]


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return None


def main():
    print('| kernel | ' + ' | '.join(c[1] for c in COLS) + ' |')
    print('|---|' + '---|' * len(COLS))
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
        names, units = rows[hdr], rows[hdr + 1]
        for r in rows[hdr + 2:]:
            if len(r) != len(names):
                continue
            rec = dict(zip(names, r))
            un = dict(zip(names, units))
            out = []
            for key, _, mul in COLS:
                v = num(rec.get(key, ''))
                if v is None:
                    out.append('-')
                    continue
                u = un.get(key, '')
                if key.startswith('gpu__time'):
                    v = v * {'ns': 1e-3, 'us': 1, 'usecond': 1, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3, 'second': 1e6}.get(u, 1e-3)
                elif 'bytes' in key:
                    v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1e-6)
                else:
                    v = v * mul
                out.append(f'{v:.4g}')
            kn = rec['Kernel Name'][:60]
            print(f'| {kn} | ' + ' | '.join(out) + ' |')


if __name__ == '__main__':
    main()
