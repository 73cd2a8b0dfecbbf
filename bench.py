#!/usr/bin/env python
"""bench.py -- image-pairs/s of the CasMTR coarse-to-fine matching hot path at 832x832 on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host cores
    torchrun ... bench.py --gpus N ...                             # one rank per GPU, pairs sharded, NCCL

Workload = BASELINE.json configs[1]: CasMTR-4c outdoor, 832x832, batch 1 per GPU: 12 QTAttB + 4 CascadeQTAttB +
CascadeMatching (2 sparse correlations, softmax/argmax, 5x5 NMS, extraction) + CascadeFineMatching per pair
(casmtr_b200/pipeline.py).  A step = one pass of that sequence over one batch of synthetic feature maps.

  value      pairs/s with the step's inputs resident in HBM (every call has its own input buffers; one step
             touches ~0.9 GB > the 126 MB L2, so nothing is served from a previous step's cache lines).  Timed passes
             of K steps each: an eager single-stream pass in which the library brackets every kernel with CUDA events
             (breakdown / roofline, `value_eager_instrumented`), the same eager step without those events
             (`ms_per_step_eager`, and `ms_per_step_eager_overlap` / `value_eager` with the library's side-stream
             transposes on), and the same step replayed as a CUDA graph with the two independent directions of each
             layer on two streams; `value` is the fastest (`execution`)
  e2e        pairs/s through the same module API with the inputs in pinned HOST memory: H2D of every input
             and D2H of the match list inside the timed region (copy stream overlapped with compute)
  roofline   the dominant kernel: algorithmic bytes per launch / its mean device time, CUDA events recorded by
             the library around each of its launches inside the timed region (casmtr_profile_*)
  cpu_baseline  the CPU oracle (a port of the reference algorithm, oracle/) timed on this host's cores on a
             bounded sample: ONE full-size call of each kind, scaled by the per-pair call counts

Prints ONE JSON line (rank 0).  The only place this file touches oracle/ is the cpu_baseline / --impl reference leg.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'image_pairs_per_sec_832x832_hot_path'
UNIT = 'pairs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=832, help='square image size (multiple of 32)')
    ap.add_argument('--pairs', type=int, default=1, help='image pairs per GPU per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time the eager single-stream path only')
    ap.add_argument('--cpu-budget-s', type=float, default=150.0, help='wall-clock bound of the reference arm')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.rows, self.proc = [], None
        sel = str(device_index)
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            sel = uuid if uuid.startswith('GPU-') else 'GPU-' + uuid
        except Exception:
            pass
        self.cmd = ['nvidia-smi', '-i', sel, f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-lms', '100']

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def start(self):
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()               # the exact process we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return None
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(smax), 'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------- CPU reference leg
def cpu_reference_sample(wl, host, threads, sync=None, ref_kernels=False):
    """One full-size call of each kind through the oracle (reference algorithm in plain torch ops, fp32); returns
    (pairs_per_s extrapolated with the per-pair call counts, seconds spent, detail dict).  `host` on the CPU = the CPU
    baseline (all host threads); the same tensors on the GPU (sync = torch.cuda.synchronize) = the "PyTorch on the same
    GPU" baseline, which is what the reference's own pure-PyTorch QTAttB (quadtree_attention_smart.py) amounts to."""
    from oracle import cascade as ocas, fine as ofine, qtatt as oqt       # the checker, used here as the baseline
    torch.set_num_threads(threads)
    qt_fn = lambda c: oqt.qtatt_b(c['q'], c['k'], c['v'], c['weight'], wl.topks, wl.nh8)
    cas_fn = lambda c: oqt.cascade_qtatt_b(c['q'], c['k'], c['v'], c['topk_pos'], None, wl.nh4)
    saved_score3d = ocas.ops.score3d
    if ref_kernels:     # the reference's own data flow around its own CUDA extension kernels (oracle/_ref, built unmodified)
        from oracle import ref_path
        qt_fn = lambda c: ref_path.qtatt_b(c['q'], c['k'], c['v'], c['weight'], wl.topks, wl.nh8)
        cas_fn = lambda c: ref_path.cascade_qtatt_b(c['q'], c['k'], c['v'], c['topk_pos'], wl.nh4)
        ocas.ops.score3d = ref_path.score3d
    t = {}
    _pc = time.perf_counter

    class _Clock:
        def __call__(self):
            if sync is not None:
                sync()
            return _pc()
    time_now = _Clock()
    with torch.no_grad():
        c = host['qt'][0]
        t0 = time_now()
        qt_fn(c)
        t['qtatt_b'] = time_now() - t0
        c = host['cas'][0]
        t0 = time_now()
        _, idx01 = cas_fn(c)
        t['cascade_qtatt_b'] = time_now() - t0
        c1 = host['cas'][1]
        idx10 = oqt.quad_to_raster(oqt.cascade_window_idx(c1['topk_pos'], wl.h4, wl.w4).reshape(wl.B, 1, -1, 1, 100)
                                   .expand(wl.B, 1, -1, 4, 100), wl.h4 // 2, wl.w4 // 2).reshape(wl.B, wl.h4 * wl.w4, 100).contiguous()
        m = host['match']
        t0 = time_now()
        o = ocas.cascade_match(m['feat0'], m['feat1'], idx01, idx10, None, None, 1.0)
        r = ocas.extract_matches(o['next_conf01'], o['next_idx01'], o['next_idx10'], (wl.h4, wl.w4), (wl.h4, wl.w4), (wl.H, wl.W),
                                 test_thr=0.2, border_rm=2, nms_window=5, pre_confs=[(m['pre_conf'], wl.h8, wl.w8)],
                                 pre_thrs=[0.2], double_check=True)
        t['cascade_matching'] = time_now() - t0
        M = min(r['mconf'].shape[0], host['fine']['feat_f0'].shape[0])
        t0 = time_now()
        ofine.fine_match(host['fine']['feat_f0'][:M], host['fine']['feat_f1'][:M], r['mkpts1_c'][:M].float(), wl.H / wl.hf)
        t['fine_matching'] = time_now() - t0
    ocas.ops.score3d = saved_score3d
    per_batch = wl.qt_calls * t['qtatt_b'] + wl.cas_calls * t['cascade_qtatt_b'] + t['cascade_matching'] + t['fine_matching']
    return wl.B / per_batch, sum(t.values()), {k: round(v, 4) for k, v in t.items()} | {'matches': int(M)}


SAMPLE_DESC = ('one full-size call of each kind (QTAttB, CascadeQTAttB, CascadeMatching+NMS+extract, FineMatching) through the '
               'CPU oracle, seconds per pair = 12*t_qt + 4*t_cas + t_match + t_fine')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from casmtr_b200 import pipeline
    wl = pipeline.Workload(args.size, args.size, pairs=args.pairs)
    host = pipeline.make_host_inputs(wl, seed=1234)
    threads = os.cpu_count() or 1
    runs, t_start = [], time.perf_counter()
    for it in range(args.warmup + args.steps):          # as many of the W+K samples as the wall-clock budget allows
        elapsed = time.perf_counter() - t_start
        if it > 0 and elapsed + elapsed / it > args.cpu_budget_s:
            break
        runs.append(cpu_reference_sample(wl, host, threads))
    done_w = min(args.warmup, len(runs) - 1)             # the first ones are warm-up, at least one is timed
    vals = [r[0] for r in runs[done_w:]]
    done_k, detail = len(vals), runs[-1][2]
    value = statistics.median(vals)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': done_k, 'warmup': done_w,
        'ms_per_step': 1000.0 * wl.B / value, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': config_dict(wl, args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': SAMPLE_DESC,
                         'seconds_per_call': detail},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': ('reference algorithm = CPU oracle port (torch CPU fp32, all host threads); the reference is Python + CUDA '
                 'extensions and /root/reference does not exist on the GPU box; requested steps/warmup are cut to fit '
                 f'--cpu-budget-s={args.cpu_budget_s:.0f}'),
    }
    print(json.dumps(line), flush=True)
    return 0


def config_dict(wl, n_gpus):
    return {'workload': wl.name, 'image': [wl.H, wl.W], 'pairs_per_gpu': wl.B, 'global_pairs': wl.B * n_gpus,
            'calls_per_pair': {'QTAttB': wl.qt_calls, 'CascadeQTAttB': wl.cas_calls, 'CascadeMatching': 1, 'CascadeFineMatching': 1},
            'topks': wl.topks, 'parallelism': f'pairs sharded over {n_gpus} GPU(s), all-gather of the match list',
            'cache': 'inputs larger than L2 (per-call buffers, ~0.9 GB per step)'}


# ---------------------------------------------------------------------------------------------- this repo's arm
def run_ours(args):
    import torch.distributed as dist
    from casmtr_b200 import dist as cdist
    from casmtr_b200 import functional as F
    from casmtr_b200 import pipeline

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; this arm has no CPU fallback (use --impl reference for the CPU leg)')
    torch.cuda.set_device(local)
    numa = cdist.bind_to_gpu_numa(local) if world > 1 else None     # before any pinned allocation (first touch)
    dev = torch.device('cuda', local)
    # stdout carries exactly one JSON line: library chatter (NCCL's version banner ...) goes to stderr until then
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world

    wl = pipeline.Workload(args.size, args.size, pairs=args.pairs)
    host = pipeline.make_host_inputs(wl, seed=1234 + 1000 * rank)      # HostFedRunner packs its own pinned blocks
    hp = pipeline.HotPath(wl).to(dev)
    hp.load_level_weights(host)
    dev_in = pipeline.tree_map(lambda t: t.to(dev), host)
    pair_offset = rank * wl.B

    cap = wl.fine_cap                                   # static per-rank capacity of the match-list all-gather

    pending = []

    def finish(out):
        # multi-GPU: the only exchange of the path, one pack kernel + one fixed-size NCCL all-gather.  It is asynchronous:
        # the next pair's kernels do not depend on it and overlap it; drain() orders the compute stream after the last one
        if world == 1:
            return out
        blocks, work = cdist.gather_matches_device(out, pair_offset, cap, async_op=True)
        pending[:] = [work]
        return blocks

    def drain():
        if pending:
            pending[-1].wait()          # collectives complete in order on the NCCL stream
            pending.clear()

    def step():
        return finish(hp(dev_in))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between a barrier + synchronize on both sides, CUDA events on the launch stream, max over ranks."""
        sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        drain()                         # the timed region ends after the last all-gather
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    for _ in range(max(args.warmup, 3)):
        out = step()
    drain()
    if world > 1:
        out = cdist.unpack_gathered(out)
    n_matches = int(out['mconf'].shape[0])
    sampler = ClockSampler(local)
    sampler.start()

    # ---- instrumented pass: K eager steps, the library brackets every kernel launch with a CUDA event pair on the launch
    # stream -> per-kernel device time (roofline, breakdown) and the launch count
    # (one stream, the library's own side-stream overlap of the transposes off: every kernel is timed alone, the times add up)
    F.profile_collect()
    prev_overlap = F.set_overlap(False)
    l0 = F.launch_count()
    F.profile_enable(True)
    ms_eager = timed(step, args.steps)
    F.profile_enable(False)
    launches = (F.launch_count() - l0) // args.steps
    prof = F.profile_collect()
    # the same eager step without the per-kernel event pairs (they serialise the launches and defeat the programmatic dependent
    # launch): once with the side-stream overlap off, once with it on (the library default)
    for _ in range(3):
        step()
    ms_eager_plain = timed(step, args.steps)
    F.set_overlap(True)
    for _ in range(3):
        step()
    ms_eager_overlap = timed(step, args.steps)
    F.set_overlap(prev_overlap)

    # ---- primary pass: the same step replayed as a CUDA graph, the two directions of every layer (independent in the
    # reference model, transformer.py:300) forked onto two streams; the graph ends at the path's one host sync (the match
    # count), the fine stage and the multi-GPU all-gather follow eagerly.  Same kernels, same results.
    ms_eager_best = min(ms_eager_plain, ms_eager_overlap)
    graph_info, ms_step, mode = None, ms_eager_best, 'eager (one stream)'
    if not args.no_graph:
        try:
            gr = pipeline.GraphRunner(hp, dev_in, two_streams=True, whole_step=True)

            def gstep():
                return finish(gr.step())
            for _ in range(max(args.warmup, 3)):
                gout = gstep()
            drain()
            gout = cdist.unpack_gathered(gout) if world > 1 else pipeline.trim_result(gout)      # the host reads the count here
            assert int(gout['mconf'].shape[0]) == n_matches, 'graph replay changed the match list'
            ms_graph = timed(gstep, args.steps)
            graph_info = {'ms_per_step': ms_graph, 'matches': n_matches}
            if ms_graph < ms_eager_best:
                ms_step, mode = ms_graph, 'CUDA graph replay of the whole step (no host sync inside), layer directions on two streams'
            del gr
        except Exception as e:      # noqa: BLE001  (capture not possible: the eager number stands)
            graph_info = {'error': str(e)[:300]}
    clocks = sampler.stop()
    value = wl.B * n_gpus / (ms_step / 1000.0)

    # ---- e2e: same module path, inputs from pinned host memory, result read back
    e2e = None
    if not args.no_e2e:
        runner = pipeline.HostFedRunner(hp, host, dev)
        for _ in range(3):
            runner.step()
        sync_all()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            res = runner.step()
            if world > 1:
                cdist.gather_matches({k: v.to(dev) for k, v in res.items() if k != 'expec_f'}, pair_offset, cap=cap)
        e1.record()
        sync_all()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        e_ms = max(e0.elapsed_time(e1), wall_ms)         # the D2H read ends on the host: take the later clock
        if world > 1:
            t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {'value': wl.B * n_gpus / (e_ms / args.steps / 1000.0), 'unit': UNIT,
               'h2d_bytes_per_step': int(runner.h2d_bytes), 'd2h_bytes_per_step': int(runner.d2h_bytes),
               'ms_per_step': e_ms / args.steps}
        del runner
        # what bounds it: the host->device link.  Measured live with a plain pinned 256 MB copy (all ranks at once, like the run)
        try:
            hb = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
            db = torch.empty_like(hb, device=dev)
            link = 0.0
            for nbytes in (32 << 20, 64 << 20, 256 << 20):          # the runner moves 44-72 MB blocks; keep the best size
                for _ in range(2):
                    db[:nbytes].copy_(hb[:nbytes], non_blocking=True)
                sync_all()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                for _ in range(8):
                    db[:nbytes].copy_(hb[:nbytes], non_blocking=True)
                c1.record()
                sync_all()
                link = max(link, 8 * nbytes / (c0.elapsed_time(c1) / 1e3) / 1e9)
            e2e['h2d_link_gbps_measured'] = round(link, 1)
            e2e['h2d_gbps_achieved'] = round(runner_gbps(e2e), 1)
            e2e['h2d_link_frac'] = round(runner_gbps(e2e) / link, 3)
            del hb, db
        except Exception as e:      # noqa: BLE001
            e2e['h2d_link_error'] = str(e)[:200]

    # ---- per-kernel breakdown and the roofline of the dominant kernel
    peak, peak_src = peaks()
    kernel_ms = sum(ms for ms, _ in prof.values())
    breakdown = []
    for kind, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        if n == 0:
            continue
        per = ms / n
        alg = wl.bytes_kernel(kind)
        row = {'kind': kind, 'launches_per_step': n / args.steps, 'ms_per_launch': round(per, 5), 'share': round(ms / kernel_ms, 4)}
        extra = ncu_stats().get(kind)
        if extra:       # from the committed ncu --set full capture of the same kernels (profiles/issue.json)
            row['ncu_issue_active_pct'] = extra['issue_active_pct']
            row['ncu_l2_to_sm_bytes'] = extra['l2_bytes']
        if alg:
            row['alg_bytes_per_launch'] = alg
            row['gbps'] = round(alg / per / 1e6, 1)
            row['frac_of_hbm_peak'] = round(alg / per / 1e6 / peak, 4)
        breakdown.append(row)
    dom = next((r for r in breakdown if 'gbps' in r), None)
    roofline = None
    if dom:
        roofline = {'bound': 'hbm', 'kernel': dom['kind'], 'achieved': dom['gbps'], 'peak': peak, 'unit': 'GB/s',
                    'frac': dom['frac_of_hbm_peak'], 'traffic': ncu_traffic(dom['kind']), 'peak_source': peak_src,
                    'alg_bytes_per_launch': dom['alg_bytes_per_launch'], 'ms_per_launch': dom['ms_per_launch'],
                    'note': ('dominant kernel by device time.  The kernels of this path are bound by instruction issue / load '
                             'latency, not by HBM or L2 bandwidth (gathers re-read each row ~16x from L2, measured L2 gather peak '
                             '15-20 TB/s): see breakdown[].ncu_issue_active_pct and DESIGN.md section 4'),
                    'ncu_issue_active_pct': dom.get('ncu_issue_active_pct')}
    qt_ms = sum(prof[k][0] for k in ('qt_coarse', 'qt_fine_mid', 'qt_fine_last')) / args.steps / wl.qt_calls
    lay = prof['layout']
    lay_per_launch = lay[0] / max(lay[1], 1)
    qt_call_ms = qt_ms + lay_per_launch                    # one layout launch per QTAttB call
    qtatt_call = {'alg_bytes': wl.bytes_qtatt_call(), 'ms': round(qt_call_ms, 5),
                  'gbps': round(wl.bytes_qtatt_call() / qt_call_ms / 1e6, 1),
                  'frac_of_hbm_peak': round(wl.bytes_qtatt_call() / qt_call_ms / 1e6 / peak, 4)}
    # the other bound SURVEY 8d asks for: fp32 FMA throughput of the SIMT pipes (top-k selection rules out reduced precision)
    S, L1, L0 = wl.h8 * wl.w8 // 16, wl.h8 * wl.w8 // 4, wl.h8 * wl.w8
    qt_flops = wl.B * (4.0 * S * S * wl.C8 + 4.0 * wl.C8 * (L1 * 4 * wl.topks[0] + L0 * 4 * wl.topks[1]))
    props = torch.cuda.get_device_properties(dev)
    simt_peak = props.multi_processor_count * 128 * 2 * (clocks.get('sm_mhz') or 1965.0) * 1e6 / 1e12      # TFLOP/s at the clock under load
    qtatt_call.update({'alg_flops': qt_flops, 'tflops': round(qt_flops / qt_call_ms / 1e9, 2), 'fp32_simt_peak_tflops': round(simt_peak, 1),
                       'frac_of_fp32_simt_peak': round(qt_flops / qt_call_ms / 1e9 / simt_peak, 4)})

    # ---- SURVEY section 8f "next" #1, reported beside the hot path (not part of `value`): dense coarse matching statistics
    # of one pair at the 1/8 grid on the tensor cores
    next_rows = None
    if rank == 0:
        try:
            g = torch.Generator().manual_seed(7)
            L8 = wl.h8 * wl.w8
            cf0 = torch.randn(wl.B, L8, wl.C8, generator=g).to(dev)
            cf1 = (0.8 * cf0.cpu()[:, torch.randperm(L8, generator=g)] + 0.6 * torch.randn(wl.B, L8, wl.C8, generator=g)).to(dev)
            for _ in range(3):
                F.coarse_match_forward(cf0, cf1, 0.1)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                F.coarse_match_forward(cf0, cf1, 0.1)
            c1.record()
            torch.cuda.synchronize()
            cms = c0.elapsed_time(c1) / 10
            alg = 2 * 2.0 * wl.B * L8 * L8 * wl.C8                      # both directions, fp32-equivalent FLOPs
            tpeak = tensor_peak()
            next_rows = {'coarse_matching': {
                'ms_per_call': cms, 'alg_tflops': alg / cms / 1e9, 'issued_tf32_tflops': 3 * alg / cms / 1e9,
                'roofline': {'bound': 'tensor', 'achieved': alg / cms / 1e9, 'peak': tpeak, 'unit': 'TFLOP/s', 'frac': alg / cms / 1e9 / tpeak,
                             'note': 'fp32-accurate path = 3 TF32 MMAs per product at half the bf16 rate: ceiling = peak / 6'},
                'what': 'CoarseMatching next_idx/next_conf (both directions) at the 1/8 grid, tcgen05 kind::tf32 3-term split'}}
            del cf0, cf1
        except Exception as e:      # noqa: BLE001
            next_rows = {'coarse_matching': {'error': str(e)[:300]}}
    if rank == 0:   # "next" #2 / #4: token-major QTAtt entry (pyramid inside) and the fine-window gather, timed alone
        def _time(fn, n=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(n):
                fn()
            c1.record()
            torch.cuda.synchronize()
            return c0.elapsed_time(c1) / n
        try:
            call = dev_in['qt'][0]
            tok = [t[0].flatten(2).transpose(1, 2).contiguous() for t in (call['q'], call['k'], call['v'])]
            t_tok = _time(lambda: F.qtatt_tokens_forward(tok[0], tok[1], tok[2], (wl.h8, wl.w8), (wl.h8, wl.w8), wl.topks, wl.nh8, weight=call['weight']))
            t_pyr = _time(lambda: F.qtatt_forward(call['q'], call['k'], call['v'], wl.topks, wl.nh8, weight=call['weight']))
            next_rows['quadtree_attention_tokens'] = {
                'ms_per_call_tokens_entry': t_tok, 'ms_per_call_nchw_pyramid_entry': t_pyr,
                'what': 'QTAttB at 1/8 from token-major level-0 q/k/v with the avg-pool pyramid built inside (casmtr_qtatt_tokens_fwd) '
                        'vs the NCHW pyramid lists of the reference API (casmtr_qtatt_fwd); the reference additionally pays 6 avg_pool2d launches upstream'}
            del tok
            cc = dev_in['cas'][0]
            centre = cc['topk_pos'][:, :, 12]                                   # window centre -> an index with the same window
            nidx = (centre[..., 0] * (wl.w4 // 2) + centre[..., 1]).contiguous()
            t_pos = _time(lambda: F.cascade_qtatt_forward(cc['q'], cc['k'], cc['v'], cc['topk_pos'], None, wl.nh4))
            t_idx = _time(lambda: F.cascade_qtatt_forward(cc['q'], cc['k'], cc['v'], nidx, None, wl.nh4))
            next_rows['cascade_window_fusion'] = {
                'ms_per_call_topk_pos': t_pos, 'ms_per_call_next_idx': t_idx,
                'what': 'CascadeQTAttB at 1/4 fed with the expanded window positions [B,L/4,25,2] (reference API) vs with next_idx [B,L/4] '
                        '(window expansion of get_window_warp_idx fused into the kernels, casmtr_cascade_qtatt_window_fwd)'}
            # "next" #3, second half: the indoor config's relative position bias, as a tensor (get_relative_pe drop-in) vs computed
            # inside the attention kernels from the two embedding tables
            g = torch.Generator().manual_seed(8)
            h8, w8 = wl.h4 // 2, wl.w4 // 2
            pe = F.RelativePE(torch.randn(22, wl.nh4, generator=g).to(dev), torch.randn(22, wl.nh4, generator=g).to(dev), 10, nidx, (h8, w8), w8)
            t_rp = _time(lambda: F.relative_pe(pe, cc['topk_pos'], (wl.h4, wl.w4)))
            rp = F.relative_pe(pe, cc['topk_pos'], (wl.h4, wl.w4))
            t_ten = _time(lambda: F.cascade_qtatt_forward(cc['q'], cc['k'], cc['v'], nidx, rp, wl.nh4))
            t_fus = _time(lambda: F.cascade_qtatt_forward(cc['q'], cc['k'], cc['v'], nidx, pe, wl.nh4))
            next_rows['relative_pe'] = {
                'ms_bias_tensor_kernel': t_rp, 'ms_per_call_bias_tensor_input': t_ten, 'ms_per_call_bias_fused': t_fus,
                'bias_tensor_bytes': rp.numel() * 4,
                'what': 'CascadeQTAttB at 1/4 with the relative position bias of the indoor config: get_relative_pe materialised by one kernel '
                        '(casmtr_relative_pe_fwd; ~25 torch ops in the reference) and read by the attention kernels, vs computed inside them '
                        'from the two embedding tables (casmtr_cascade_qtatt_relpe_fwd)'}
            del rp
            # "next" #4, second half: backward of the three op-level drop-ins at the shapes of the hot path, next to the
            # reference's own backward kernels (oracle/_ref, built unmodified) where they are available
            g = torch.Generator().manual_seed(10)
            L2q, L2k = (wl.h8 // 2) * (wl.w8 // 2), wl.h8 * wl.w8                 # last QTAttB level: parents, keys
            q5 = torch.randn(wl.B, L2q, 4, wl.nh8, 32, generator=g).to(dev)
            k5 = torch.randn(wl.B, L2k, wl.nh8, 32, generator=g).to(dev)
            i5 = torch.randint(0, L2k, (wl.B, L2q, 4 * wl.topks[1], wl.nh8), generator=g).to(dev)
            go5 = torch.randn(wl.B, L2q, 4, 4 * wl.topks[1], wl.nh8, generator=g).to(dev)
            s5 = torch.rand(wl.B, 4 * L2q, 4 * wl.topks[1], wl.nh8, generator=g).to(dev)
            i5v = i5.view(wl.B, L2q, 1, -1, wl.nh8).expand(-1, -1, 4, -1, -1).reshape(wl.B, 4 * L2q, -1, wl.nh8).contiguous()
            gov = torch.randn(wl.B, 4 * L2q, wl.nh8, 32, generator=g).to(dev)
            L4 = wl.h4 * wl.w4
            q3, k3 = torch.randn(wl.B, L4, wl.C4, generator=g).to(dev), torch.randn(wl.B, L4, wl.C4, generator=g).to(dev)
            i3 = torch.randint(0, L4, (wl.B, L4, 100), generator=g).to(dev)
            go3 = torch.randn(wl.B, L4, 100, generator=g).to(dev)
            ob = {'ms_score5d_bwd': _time(lambda: F.score5d_backward(go5, q5, k5, i5), n=10),
                  'ms_value_agg_bwd': _time(lambda: F.value_agg_backward(gov, s5, k5, i5v), n=10),
                  'ms_score3d_bwd': _time(lambda: F.score3d_backward(go3, q3, k3, i3), n=10),
                  'shapes': {'score5d': [wl.B, L2q, L2k, wl.nh8, 32, 4 * wl.topks[1]], 'value_agg': [wl.B, 4 * L2q, 4 * wl.topks[1], wl.nh8, L2k, 32],
                             'score3d': [wl.B, L4, L4, wl.C4, 100]},
                  'what': 'backward of score5d / value_agg at the last QTAttB level and of score3d at the 1/4 cascade level (random indices), '
                          'this library vs the reference kernels (scalar atomicAdd per element) on the same GPU'}
            try:
                from oracle import build_ref
                if all(build_ref.built(n) for n in ('score_computation_cuda', 'value_aggregation_cuda', 'fast_score_computation')):
                    r5, rv, r3 = (build_ref.load(n) for n in ('score_computation_cuda', 'value_aggregation_cuda', 'fast_score_computation'))
                    ob['ms_score5d_bwd_reference'] = _time(lambda: r5.score_backward(go5, q5, k5, i5), n=5)
                    ob['ms_value_agg_bwd_reference'] = _time(lambda: rv.value_aggregation_backward(gov, s5, k5, i5v, torch.zeros_like(s5), torch.zeros_like(k5)), n=5)
                    ob['ms_score3d_bwd_reference'] = _time(lambda: r3.score_backward(go3, q3, k3, i3), n=3)
            except Exception as e:      # noqa: BLE001
                ob['reference_error'] = str(e)[:200]
            next_rows['op_backward'] = ob
            del q5, k5, i5, go5, s5, i5v, gov, q3, k3, i3, go3
            M = max(n_matches, 1)
            g = torch.Generator().manual_seed(9)
            ff = torch.randn(wl.B, 64, wl.hf, wl.wf, generator=g).to(dev)
            bi = torch.zeros(M, dtype=torch.int64, device=dev)
            ii = torch.randint(0, wl.h4 * wl.w4, (M,), generator=g).to(dev)
            t_g = _time(lambda: F.fine_window_gather(ff, bi, ii, wl.w4, wl.hf // wl.h4, 5))
            t_u = _time(lambda: torch.nn.functional.unfold(ff, (5, 5), stride=wl.hf // wl.h4, padding=2).reshape(wl.B, 64, 25, -1).permute(0, 3, 2, 1)[bi, ii], n=5)
            next_rows['fine_preprocess'] = {'ms_per_map_gather': t_g, 'ms_per_map_unfold_select_torch': t_u, 'matches': M,
                                            'what': 'CascadeFinePreprocess window crop of one fine map: gather kernel vs the reference formulation (F.unfold + select) on the same GPU'}
            del ff
        except Exception as e:      # noqa: BLE001
            next_rows['next_rows_error'] = str(e)[:300]
    host_ms = None
    if True:        # host-side cost of enqueueing one step (no device wait): how launch-bound the path is
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hp_out = None
        for i, call in enumerate(dev_in['qt']):
            hp.run_qt(i, call)
        for i, call in enumerate(dev_in['cas']):
            hp.run_cas(i, call)
        host_ms = (time.perf_counter() - t0) * 1000.0
        torch.cuda.synchronize()
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': config_dict(wl, n_gpus), 'e2e': e2e, 'gpu_launches': int(launches) * args.steps,
        'gpu_launches_per_step': int(launches), 'clocks': clocks, 'roofline': roofline, 'qtatt_call_roofline': qtatt_call,
        'execution': mode, 'ms_per_step_eager_instrumented': ms_eager, 'ms_per_step_eager': ms_eager_plain, 'ms_per_step_eager_overlap': ms_eager_overlap, 'value_eager_instrumented': wl.B * n_gpus / (ms_eager / 1000.0),
        'value_eager': wl.B * n_gpus / (ms_eager_overlap / 1000.0),
        'cuda_graph': graph_info, 'numa_binding': numa, 'next_rows': next_rows, 'kernel_ms_per_step': round(kernel_ms / args.steps, 4), 'host_enqueue_ms_attention_calls': round(host_ms, 3), 'breakdown': breakdown, 'matches_per_step': n_matches,
    }
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:        # the same algorithm as plain torch CUDA ops on this GPU (extra context, not part of the contract)
            cpu_reference_sample(wl, dev_in, os.cpu_count() or 1, sync=torch.cuda.synchronize)
            v, spent, detail = cpu_reference_sample(wl, dev_in, os.cpu_count() or 1, sync=torch.cuda.synchronize)
            line['gpu_torch_baseline'] = {'value': v, 'unit': UNIT, 'seconds_per_call': detail,
                                          'what': 'oracle (plain torch ops, the formulation of the reference\'s own pure-PyTorch '
                                                  'QTAttB) on the same GPU, one call of each kind scaled by the call counts'}
        except Exception as e:      # noqa: BLE001  (out of memory etc.: the figure is optional)
            line['gpu_torch_baseline'] = {'error': str(e)[:200]}
        try:        # the reference build on the same B200 = its data flow around its own extension kernels
            from oracle import ref_path
            if ref_path.available():
                cpu_reference_sample(wl, dev_in, os.cpu_count() or 1, sync=torch.cuda.synchronize, ref_kernels=True)
                v, spent, detail = cpu_reference_sample(wl, dev_in, os.cpu_count() or 1, sync=torch.cuda.synchronize, ref_kernels=True)
                line['gpu_reference_kernels_baseline'] = {
                    'value': v, 'unit': UNIT, 'seconds_per_call': detail,
                    'speedup_eager_vs_eager': round(line['value_eager'] / v, 1),
                    'what': "the reference's GPU path on this B200: its QTAttB / CascadeQTAttB / ScoreComputation data flow (oracle/ref_path.py) "
                            'calling its own three CUDA extensions built unmodified for sm_100a (oracle/_ref), eager, one call of each kind '
                            'scaled by the call counts; matching post-processing and fine matching as plain torch CUDA ops'}
            else:
                line['gpu_reference_kernels_baseline'] = {'unavailable': 'oracle/_ref not built'}
        except Exception as e:      # noqa: BLE001
            line['gpu_reference_kernels_baseline'] = {'error': str(e)[:200]}
        torch.cuda.empty_cache()
        runs, t_cpu = [], time.perf_counter()                 # bounded sample: ~10 s of host work, median of the repeats
        while len(runs) < 2 or (time.perf_counter() - t_cpu < 10.0 and len(runs) < 16):
            runs.append(cpu_reference_sample(wl, host, os.cpu_count() or 1))
        runs = runs[1:]                                       # the first repeat warms the allocator and the thread pool
        v = statistics.median(r[0] for r in runs)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': os.cpu_count() or 1, 'kind': 'port',
                                'sample': SAMPLE_DESC + f'; median of {len(runs)} repeats', 'seconds_per_call': runs[-1][2],
                                'seconds_spent': round(time.perf_counter() - t_cpu, 2)}
    else:
        line['cpu_baseline'] = None
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def runner_gbps(e2e):
    """host->device bytes per second per GPU of the e2e run."""
    return e2e['h2d_bytes_per_step'] / (e2e['ms_per_step'] / 1e3) / 1e9


def tensor_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['bf16_tflops'])
    return 1590.0


def ncu_stats():
    path = os.path.join(ROOT, 'profiles', 'issue.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def ncu_traffic(kind):
    """dram bytes per launch of the kernel from the committed ncu --set full capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(kind)
    return None


if __name__ == '__main__':
    a = parse()
    sys.exit(run_reference(a) if a.impl == 'reference' else run_ours(a))
