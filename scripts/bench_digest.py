#!/usr/bin/env python
"""Print the few numbers of a bench.py JSON line one wants to see in a terminal tail."""
import json
import sys


def main(path):
    try:
        with open(path) as fh:
            line = [l for l in fh.read().splitlines() if l.startswith('{')][-1]
        d = json.loads(line)
    except Exception as e:      # noqa: BLE001
        print(f'{path}: no JSON line ({e})')
        return
    cfg = d.get('config', {})
    print(f"== {path}: {cfg.get('workload')}")
    print(f"   value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  [{d.get('execution')}]  eager {d.get('ms_per_step_eager', 0):.3f} ms  "
          f"launches/step {d.get('gpu_launches_per_step')}  matches {d.get('matches_per_step')}  n_gpus {d.get('n_gpus')}")
    if d.get('e2e'):
        e = d['e2e']
        print(f"   e2e {e['value']:.1f} pairs/s  h2d {e.get('h2d_bytes_per_step', 0) / 1e6:.0f} MB/step  link {e.get('h2d_link_gbps_measured')} GB/s  achieved {e.get('h2d_gbps_achieved')}")
    q = d.get('qtatt_call_roofline')
    if q:
        print(f"   QTAttB call-equivalent {q['us_per_call_equivalent']} us  hbm {q['frac_of_hbm_peak']}  simt {q['frac_of_fp32_simt_peak']}  {q['us_by_kernel']}")
        if q.get('graph_replay'):
            print(f"   QTAttB calls alone as one graph: {q['graph_replay']}")
    r = d.get('roofline')
    if r:
        print(f"   roofline {r['kernel']}: {r['achieved']} GB/s frac {r['frac']}  {r['ms_per_launch']} ms/launch")
    for row in d.get('breakdown', []):
        print(f"     {row['kind']:18s} {row['launches_per_step']:5.1f}/step  {1e3 * row['ms_per_launch']:8.1f} us  share {row['share']:.3f}  frac_hbm {row.get('frac_of_hbm_peak')}")
    for s in d.get('batch_sweep') or []:
        if 'error' in s:
            print(f"   sweep P={s['pairs_per_step']}: {s['error']}")
            continue
        q = s['qtatt_call_roofline']
        print(f"   sweep P={s['pairs_per_step']}: {s['value']:.1f} pairs/s  {s['ms_per_step']:.3f} ms/step  QTAttB {q['us_per_call_equivalent']} us/call  hbm {q['frac_of_hbm_peak']} simt {q['frac_of_fp32_simt_peak']}  {q['us_by_kernel']}")
    for k in ('gpu_torch_baseline', 'gpu_reference_kernels_baseline', 'cpu_baseline'):
        if d.get(k):
            b = d[k]
            print(f"   {k}: " + (f"{b['value']:.3f} pairs/s p10 {b.get('p10')} p90 {b.get('p90')} {b.get('speedup_eager_vs_eager', '')}" if 'value' in b else str(b)[:200]))
    if d.get('allgather_ms') is not None:
        print(f"   allgather_ms {d['allgather_ms']:.4f}")
    print(f"   clocks {d.get('clocks')}")


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
