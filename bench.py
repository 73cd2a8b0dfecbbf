#!/usr/bin/env python
"""bench.py -- image-pairs/s of the CasMTR coarse-to-fine matching hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (C ABI), CasMTR-4c 832x832 batch 1
    python bench.py --config 2c --pairs 8                          # BASELINE.json configs[2]
    python bench.py --config indoor --size 640x480 --pairs 4       # configs[3] (4 pairs per GPU; 8 ranks under torchrun)
    python bench.py --size 1024 --global-pairs 16                  # configs[4]: 16 pairs split over the ranks (strong scaling)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host cores
    torchrun ... bench.py --gpus N ...                             # one rank per GPU, pairs sharded, NCCL

Default workload = BASELINE.json configs[1]: CasMTR-4c outdoor, 832x832, batch 1 per GPU: 12 QTAttB + 4 CascadeQTAttB +
CascadeMatching (2 sparse correlations, softmax/argmax, 5x5 NMS, extraction) + CascadeFineMatching per pair
(casmtr_b200/pipeline.py).  The two directions of every layer are stacked on the batch dimension (one launch set per layer),
the attention layers are fed token-major (`--entry nchw` feeds the reference's NCHW pyramid lists instead).
A step = one pass of that sequence over one batch of synthetic feature maps.

  value      pairs/s with the step's inputs resident in HBM (every call has its own input buffers; one step touches ~0.7 GB
             per pair > the 126 MB L2, so nothing is served from a previous step's cache lines).  Timed passes of K steps each:
             an eager pass in which the library brackets every kernel with CUDA events (breakdown / roofline), the same eager
             step without those events, and the step replayed as ONE CUDA graph; `value` is the fastest (`execution`)
  e2e        pairs/s through the same module API with the inputs in pinned HOST memory: H2D of every input and D2H of the
             match list inside the timed region (copy stream, double-buffered device blocks)
  roofline   the dominant kernel: algorithmic bytes per launch / its mean device time, CUDA events recorded by the library
             around each of its launches inside the timed region (casmtr_profile_*)
  batch_sweep  (N = 1, default workload) the same measurement at 2 / 4 / 8 pairs per step
  cpu_baseline  the CPU oracle (a port of the reference algorithm, oracle/) running the SAME full call sequence on this host's
             cores, a bounded number of whole steps

Prints ONE JSON line (rank 0).  The only places this file touches oracle/ are the cpu_baseline / gpu_*_baseline legs and the
--impl reference arm.
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

UNIT = 'pairs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='4c', choices=['4c', '2c', 'indoor'])
    ap.add_argument('--size', default='832', help='square image size, or WIDTHxHEIGHT (multiples of 32)')
    ap.add_argument('--pairs', type=int, default=1, help='image pairs per GPU per step (weak scaling)')
    ap.add_argument('--global-pairs', type=int, default=0, help='total pairs per step, split over the ranks (strong scaling)')
    ap.add_argument('--entry', default='tokens', choices=['tokens', 'nchw'])
    ap.add_argument('--simt-coarse', action='store_true', help='dense coarsest QTAtt level on the fp32 SIMT kernel (A/B)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-baselines', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time the eager path only')
    ap.add_argument('--no-sweep', action='store_true', help='skip the batch sweep')
    ap.add_argument('--no-next-rows', action='store_true', help='skip the widening rows timed beside the path')
    ap.add_argument('--cpu-budget-s', type=float, default=150.0, help='wall-clock bound of the reference arm')
    a = ap.parse_args()
    if 'x' in a.size:
        w, h = a.size.split('x')
        a.width, a.height = int(w), int(h)
    else:
        a.width = a.height = int(a.size)
    return a


def metric_name(a):
    if a.config == '4c' and a.width == a.height:
        return f'image_pairs_per_sec_{a.width}x{a.height}_hot_path'
    return f'image_pairs_per_sec_{a.width}x{a.height}_hot_path_{a.config}'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p['hbm_gbs']), float(p.get('bf16_tflops', 1590.0)), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, 'fallback (B200_PROFILING.md)'


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


# ---------------------------------------------------------------------------------------------- reference / baseline legs
def reference_steps(wl, inp, n_warm, n_steps, budget_s, impl='oracle', sync=None):
    """Whole steps (the full call sequence, every call on its own inputs) through oracle/hotpath.py.  Runs n_warm + n_steps of
    them unless the wall-clock budget runs out first; returns (seconds per timed step [list], warm-ups done, per-kind seconds of
    the last step, matches)."""
    from oracle import hotpath             # the checker, used here as the baseline
    secs, t_start, kinds, n_match = [], time.perf_counter(), {}, 0
    for it in range(n_warm + n_steps):
        elapsed = time.perf_counter() - t_start
        if it >= 2 and elapsed + elapsed / it > budget_s:
            break
        kinds = {}
        if sync is not None:
            sync()
        t0 = time.perf_counter()
        out = hotpath.run_step(wl, inp, impl=impl, sync=sync, times=kinds)
        if sync is not None:
            sync()
        secs.append(time.perf_counter() - t0)
        n_match = int(out['mconf'].shape[0])
    warm = min(n_warm, len(secs) - 1)
    return secs[warm:], warm, {k: round(v, 4) for k, v in kinds.items()}, n_match


REF_SAMPLE = ('whole steps of the full call sequence ({calls}) on the synthetic inputs of the GPU arm, CPU oracle port '
              '(torch CPU fp32, all host threads)')


def make_workload(args, world):
    from casmtr_b200 import pipeline
    pairs = args.pairs
    scaling = 'weak'
    if args.global_pairs:
        assert args.global_pairs % world == 0, '--global-pairs must divide by the number of ranks'
        pairs, scaling = args.global_pairs // world, 'strong'
    return pipeline.Workload(args.height, args.width, pairs=pairs, config=args.config, entry=args.entry), scaling


def config_dict(wl, n_gpus, scaling):
    return {'workload': wl.name, 'config': wl.config, 'image': [wl.H, wl.W], 'pairs_per_gpu': wl.P, 'global_pairs': wl.P * n_gpus,
            'calls_per_pair': wl.calls_per_pair,
            'launch_batching': f'the 2 directions of every layer x {wl.P} pair(s) stacked on the batch dimension (B = {wl.B} per call)',
            'entry': ('token-major level-0 q/k/v, pyramid built inside (QuadtreeAttention boundary)' if wl.entry == 'tokens'
                      else 'NCHW pyramid lists (QTAttB.forward boundary)'),
            'topks': wl.topks, 'parallelism': f'pairs sharded over {n_gpus} GPU(s) ({scaling} scaling), all-gather of the match list',
            'cache': 'inputs larger than L2 (per-call buffers, >= 0.6 GB per step)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return 0
    from casmtr_b200 import pipeline
    wl, scaling = make_workload(args, world)
    host = pipeline.make_host_inputs(wl, seed=1234)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    secs, warm, kinds, n_match = reference_steps(wl, host, args.warmup, args.steps, args.cpu_budget_s)
    sec = statistics.median(secs)
    value = wl.P / sec
    line = {
        'impl': 'reference', 'metric': metric_name(args), 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(secs), 'warmup': warm,
        'ms_per_step': 1000.0 * sec, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': config_dict(wl, args.gpus, scaling),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': REF_SAMPLE.format(calls=wl.calls_per_pair) + f'; median of {len(secs)} timed steps after {warm} warm-up',
                         'seconds_per_kind_last_step': kinds, 'matches': n_match},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'extrapolated': False,
        'note': ('reference algorithm = CPU oracle port (torch CPU fp32, all host threads) executing every call of the step; the '
                 'reference itself is Python + CUDA extensions and /root/reference does not exist on the GPU box.  `steps` / `warmup` '
                 f'are what actually ran: requested {args.steps} / {args.warmup}, cut only if --cpu-budget-s={args.cpu_budget_s:.0f} runs out'),
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------- this repo's arm
def percentiles(xs):
    xs = sorted(xs)
    pick = lambda q: xs[min(len(xs) - 1, max(0, int(round(q * (len(xs) - 1)))))]
    return {'median': statistics.median(xs), 'p10': pick(0.10), 'p90': pick(0.90), 'n': len(xs)}


def measure(hp, dev_in, wl, steps, warmup, finish, drain, timed, graph=True):
    """Instrumented eager pass (per-kernel device time), plain eager pass, whole-step CUDA graph.  Returns a dict."""
    from casmtr_b200 import functional as F
    from casmtr_b200 import pipeline
    r = {}

    def step():
        return finish(hp(dev_in))
    for _ in range(max(warmup, 3)):
        out = step()
    drain()
    r['out'] = out
    F.profile_collect()
    prev_overlap = F.set_overlap(False)         # one stream: every kernel is timed alone, the times add up
    l0 = F.launch_count()
    F.profile_enable(True)
    r['ms_eager_instrumented'] = timed(step, steps)
    F.profile_enable(False)
    r['launches_per_step'] = (F.launch_count() - l0) // steps
    r['prof'] = F.profile_collect()
    F.set_overlap(prev_overlap)
    for _ in range(3):
        step()
    r['ms_eager'] = timed(step, steps)
    r['ms_step'], r['mode'] = r['ms_eager'], 'eager (one stream)'
    r['graph'] = None
    if graph:
        try:
            gr = pipeline.GraphRunner(hp, dev_in)

            def gstep():
                return finish(gr.step())
            for _ in range(max(warmup, 3)):
                gout = gstep()
            drain()
            r['gout'] = gout
            ms_graph = timed(gstep, steps)
            r['graph'] = {'ms_per_step': ms_graph}
            if ms_graph < r['ms_step']:
                r['ms_step'], r['mode'] = ms_graph, 'CUDA graph replay of the whole step (one stream, no host sync inside)'
            del gr
        except Exception as e:      # noqa: BLE001  (capture not possible: the eager number stands)
            r['graph'] = {'error': str(e)[:300]}
    return r


def breakdown_rows(prof, wl, steps, peak):
    kernel_ms = sum(ms for ms, _ in prof.values())
    rows = []
    for kind, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        if n == 0:
            continue
        per = ms / n
        alg = wl.bytes_kernel(kind)
        row = {'kind': kind, 'launches_per_step': n / steps, 'ms_per_launch': round(per, 5), 'share': round(ms / kernel_ms, 4)}
        extra = ncu_stats().get(kind)
        if extra:       # from the committed ncu --set full capture of the same kernels (profiles/issue.json)
            row['ncu_issue_active_pct'] = extra.get('issue_active_pct')
            row['ncu_l2_to_sm_bytes'] = extra.get('l2_bytes')
        if alg:
            row['alg_bytes_per_launch'] = alg
            row['gbps'] = round(alg / per / 1e6, 1)
            row['frac_of_hbm_peak'] = round(alg / per / 1e6 / peak, 4)
        rows.append(row)
    return rows, kernel_ms


def qtatt_call_stats(prof, wl, steps, peak, simt_peak):
    """One QTAttB call-equivalent (one direction of one pair): device time of the QTAtt kernels (+ their share of the layout /
    pooling launches) divided by the calls a step makes, against the HBM and the fp32-SIMT roofs."""
    calls = wl.qt_layers * wl.B
    ms = sum(prof[k][0] for k in ('qt_coarse', 'qt_fine_mid', 'qt_fine_last')) / steps
    lay_n = prof['layout'][1] / steps
    lay_qt = prof['layout'][0] / steps * min(1.0, wl.qt_layers / max(lay_n, 1.0)) if wl.entry == 'nchw' else None
    # token entry: the pooling launches are the only `layout` launches the attention layers make
    lay_ms = prof['layout'][0] / steps if wl.entry == 'tokens' else lay_qt
    call_ms = (ms + lay_ms) / calls
    by = wl.bytes_qtatt_call()
    fl = wl.flops_qtatt_call()
    return {'us_per_call_equivalent': round(1e3 * call_ms, 2), 'calls_per_step': calls, 'alg_bytes': by,
            'gbps': round(by / call_ms / 1e6, 1), 'frac_of_hbm_peak': round(by / call_ms / 1e6 / peak, 4),
            'alg_flops': fl, 'tflops': round(fl / call_ms / 1e9, 2), 'fp32_simt_peak_tflops': round(simt_peak, 1),
            'frac_of_fp32_simt_peak': round(fl / call_ms / 1e9 / simt_peak, 4),
            'us_by_kernel': {k: round(1e3 * prof[k][0] / steps / calls, 2) for k in ('qt_coarse', 'qt_fine_mid', 'qt_fine_last')} |
                            {'layout_pool': round(1e3 * lay_ms / calls, 2)}}


def qtatt_graph_stats(hp, dev_in, wl, timed, steps, peak, simt_peak):
    """All QTAttB calls of a step (no cascade, no matching) captured as one CUDA graph and replayed: device time per call-equivalent."""
    import torch

    def calls():
        for i, call in enumerate(dev_in['qt']):
            hp.run_qt(i, call)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(3):
            calls()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        calls()
    for _ in range(3):
        g.replay()
    ms = timed(g.replay, steps)
    n = wl.qt_layers * wl.B
    call_ms = ms / n
    by, fl = wl.bytes_qtatt_call(), wl.flops_qtatt_call()
    return {'us_per_call_equivalent': round(1e3 * call_ms, 2), 'ms_all_calls': round(ms, 4), 'calls': n,
            'frac_of_hbm_peak': round(by / call_ms / 1e6 / peak, 4), 'frac_of_fp32_simt_peak': round(fl / call_ms / 1e9 / simt_peak, 4)}


def run_ours(args):
    import torch.distributed as dist
    from casmtr_b200 import _lib
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

    wl, scaling = make_workload(args, world)
    qt_flags = _lib.QT_SIMT_COARSE if args.simt_coarse else 0
    host = pipeline.make_host_inputs(wl, seed=1234 + 1000 * rank)      # HostFedRunner packs its own pinned blocks
    hp = pipeline.HotPath(wl, qt_flags=qt_flags).to(dev)
    hp.load_level_weights(host)
    dev_in = pipeline.tree_map(lambda t: t.to(dev), host)
    pair_offset = rank * wl.P
    cap = wl.fine_cap                                   # static per-rank capacity of the match-list all-gather
    pending = []

    def finish(out):
        # multi-GPU: the only exchange of the path, one pack kernel + one fixed-size NCCL all-gather.  It is asynchronous:
        # the next step's kernels do not depend on it and overlap it; drain() orders the compute stream after the last one
        if world == 1:
            return out
        blocks, work = cdist.gather_matches_device(out, pair_offset, cap, async_op=True)
        pending[:] = [work]
        return blocks

    def drain():
        if pending:
            pending[-1].wait()          # collectives complete in order on the NCCL stream
            pending.clear()

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

    sampler = ClockSampler(local)
    sampler.start()
    m = measure(hp, dev_in, wl, args.steps, args.warmup, finish, drain, timed, graph=not args.no_graph)
    clocks = sampler.stop()
    out = cdist.unpack_gathered(m['out']) if world > 1 else m['out']
    n_matches = int(out['mconf'].shape[0])
    if 'gout' in m:
        gout = cdist.unpack_gathered(m['gout']) if world > 1 else pipeline.trim_result(m['gout'])      # the host reads the count here
        assert int(gout['mconf'].shape[0]) == n_matches, 'graph replay changed the match list'
        m['graph']['matches'] = n_matches
    ms_step = m['ms_step']
    value = wl.P * n_gpus / (ms_step / 1000.0)

    # the all-gather alone (multi-GPU): pack kernel + NCCL all-gather of the fixed-size block, device time, max over ranks
    allgather_ms = None
    if world > 1:
        res = hp(dev_in)
        allgather_ms = timed(lambda: cdist.gather_matches_device(res, pair_offset, cap), 10)

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
        e2e = {'value': wl.P * n_gpus / (e_ms / args.steps / 1000.0), 'unit': UNIT,
               'h2d_bytes_per_step': int(runner.h2d_bytes), 'd2h_bytes_per_step': int(runner.d2h_bytes),
               'ms_per_step': e_ms / args.steps,
               'how': 'HostFedRunner: one pinned block per call, copy stream, two device buffers per call (the upload of step n+1 overlaps step n)'}
        del runner
        # what bounds it: the host->device link.  Measured live with a plain pinned copy (all ranks at once, like the run)
        try:
            hb = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
            db = torch.empty_like(hb, device=dev)
            link = 0.0
            for nbytes in (32 << 20, 64 << 20, 256 << 20):          # the runner moves 30-70 MB blocks; keep the best size
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
    peak, tpeak, peak_src = peaks()
    prof = m['prof']
    breakdown, kernel_ms = breakdown_rows(prof, wl, args.steps, peak)
    dom = next((r for r in breakdown if 'gbps' in r), None)
    roofline = None
    if dom:
        roofline = {'bound': 'hbm', 'kernel': dom['kind'], 'achieved': dom['gbps'], 'peak': peak, 'unit': 'GB/s',
                    'frac': dom['frac_of_hbm_peak'], 'traffic': ncu_traffic(dom['kind']), 'peak_source': peak_src,
                    'alg_bytes_per_launch': dom['alg_bytes_per_launch'], 'ms_per_launch': dom['ms_per_launch'],
                    'note': ('dominant kernel by device time, algorithmic bytes of ONE launch (= %d stacked calls).  The gather kernels of '
                             'this path re-read each 128-byte row several times from L2 and are bound by instruction issue / load '
                             'latency rather than HBM: see breakdown[] and DESIGN.md section 4' % wl.B),
                    'ncu_issue_active_pct': dom.get('ncu_issue_active_pct')}
    props = torch.cuda.get_device_properties(dev)
    simt_peak = props.multi_processor_count * 128 * 2 * ((clocks or {}).get('sm_mhz') or 1965.0) * 1e6 / 1e12      # TFLOP/s at the clock under load
    qtatt_call = qtatt_call_stats(prof, wl, args.steps, peak, simt_peak)
    # the same calls alone as ONE CUDA graph (no per-kernel event pairs, launch chaining on): the call-equivalent a step really pays
    if not args.no_graph and n_gpus == 1 and rank == 0:
        try:
            qtatt_call['graph_replay'] = qtatt_graph_stats(hp, dev_in, wl, timed, args.steps, peak, simt_peak)
        except Exception as e:      # noqa: BLE001
            qtatt_call['graph_replay'] = {'error': str(e)[:200]}

    # ---- batch sweep: the same step at 2 / 4 / 8 pairs per launch set
    batch_sweep = None
    if rank == 0 and n_gpus == 1 and not args.no_sweep and not args.global_pairs and wl.P == 1:
        batch_sweep = []
        for P in (2, 4, 8):
            try:
                wl2 = pipeline.Workload(args.height, args.width, pairs=P, config=args.config, entry=args.entry)
                host2 = pipeline.make_host_inputs(wl2, seed=4321)
                hp2 = pipeline.HotPath(wl2, qt_flags=qt_flags).to(dev)
                hp2.load_level_weights(host2)
                dev2 = pipeline.tree_map(lambda t: t.to(dev), host2)
                del host2
                ks = max(5, args.steps // 2)
                m2 = measure(hp2, dev2, wl2, ks, 3, lambda o: o, lambda: None, timed, graph=not args.no_graph)
                rows2, kms2 = breakdown_rows(m2['prof'], wl2, ks, peak)
                batch_sweep.append({'pairs_per_step': P, 'value': P / (m2['ms_step'] / 1e3), 'ms_per_step': m2['ms_step'], 'execution': m2['mode'],
                                    'ms_per_step_eager': m2['ms_eager'], 'kernel_ms_per_step': round(kms2 / ks, 4),
                                    'qtatt_call_roofline': qtatt_call_stats(m2['prof'], wl2, ks, peak, simt_peak),
                                    'breakdown': [{k: r[k] for k in ('kind', 'ms_per_launch', 'share', 'gbps', 'frac_of_hbm_peak') if k in r} for r in rows2]})
                del hp2, dev2, m2
                torch.cuda.empty_cache()
            except Exception as e:      # noqa: BLE001
                batch_sweep.append({'pairs_per_step': P, 'error': str(e)[:300]})

    next_rows = None
    if rank == 0 and not args.no_next_rows:
        next_rows = time_next_rows(wl, dev_in, dev, tpeak, n_matches)

    host_ms = None
    if True:        # host-side cost of enqueueing one step's attention calls (no device wait): how launch-bound the path is
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, call in enumerate(dev_in['qt']):
            hp.run_qt(i, call)
        host_ms = (time.perf_counter() - t0) * 1000.0
        torch.cuda.synchronize()
    line = {
        'metric': metric_name(args), 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': config_dict(wl, n_gpus, scaling), 'e2e': e2e, 'gpu_launches': int(m['launches_per_step']) * args.steps,
        'gpu_launches_per_step': int(m['launches_per_step']), 'clocks': clocks, 'roofline': roofline, 'qtatt_call_roofline': qtatt_call,
        'execution': m['mode'], 'ms_per_step_eager_instrumented': m['ms_eager_instrumented'], 'ms_per_step_eager': m['ms_eager'],
        'value_eager': wl.P * n_gpus / (m['ms_eager'] / 1000.0), 'cuda_graph': m['graph'], 'allgather_ms': allgather_ms, 'numa_binding': numa,
        'batch_sweep': batch_sweep, 'next_rows': next_rows, 'kernel_ms_per_step': round(kernel_ms / args.steps, 4),
        'host_enqueue_ms_qtatt_calls': round(host_ms, 3), 'breakdown': breakdown, 'matches_per_step': n_matches,
        'coarse_level': 'fp32 SIMT kernel (--simt-coarse)' if args.simt_coarse else 'tcgen05 kernel where the shape allows (see DESIGN.md)',
    }
    if rank == 0 and n_gpus == 1 and not args.no_gpu_baselines:
        # the same algorithm on this GPU (extra context, not part of the contract): whole steps, CUDA-synchronised wall clock
        for key, impl, what in (
                ('gpu_torch_baseline', 'oracle', "oracle (plain torch ops, the formulation of the reference's own pure-PyTorch QTAttB) on the same GPU"),
                ('gpu_reference_kernels_baseline', 'ref_kernels',
                 "the reference's GPU path on this B200: its QTAttB / CascadeQTAttB / ScoreComputation data flow (oracle/ref_path.py) calling its "
                 'own three CUDA extensions built unmodified for sm_100a (oracle/_ref), eager; matching post-processing and fine matching as plain torch CUDA ops')):
            try:
                if impl == 'ref_kernels':
                    from oracle import ref_path
                    if not ref_path.available():
                        line[key] = {'unavailable': 'oracle/_ref not built'}
                        continue
                secs, warm, kinds, nm = reference_steps(wl, dev_in, 2, 12, 60.0, impl=impl, sync=torch.cuda.synchronize)
                pc = percentiles([wl.P / s for s in secs])
                line[key] = {'value': pc['median'], 'unit': UNIT, 'p10': pc['p10'], 'p90': pc['p90'], 'steps': pc['n'], 'warmup': warm,
                             'seconds_per_kind_last_step': kinds, 'matches': nm, 'what': what + '; whole steps, every call executed'}
                if impl == 'ref_kernels':
                    line[key]['speedup_eager_vs_eager'] = {'median': round(line['value_eager'] / pc['median'], 1),
                                                           'worst_case': round(line['value_eager'] / pc['p90'], 1)}
                    line[key]['speedup_graph_vs_eager'] = round(value / pc['median'], 1)
            except Exception as e:      # noqa: BLE001  (out of memory etc.: the figure is optional)
                line[key] = {'error': str(e)[:200]}
            torch.cuda.empty_cache()
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        t_cpu = time.perf_counter()
        secs, warm, kinds, nm = reference_steps(wl, host, 1, 8, 25.0)      # bounded sample: <= ~25 s of host work
        v = wl.P / statistics.median(secs)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                'sample': REF_SAMPLE.format(calls=wl.calls_per_pair) + f'; median of {len(secs)} whole steps after {warm} warm-up',
                                'seconds_per_kind_last_step': kinds, 'matches': nm, 'seconds_spent': round(time.perf_counter() - t_cpu, 2)}
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


def time_next_rows(wl, dev_in, dev, tpeak, n_matches):
    """SURVEY section 8f "next" rows timed beside the hot path (not part of `value`)."""
    from casmtr_b200 import functional as F
    rows = {}

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
    try:        # n1: dense coarse matching statistics of the step's pairs at the 1/8 grid on the tensor cores
        g = torch.Generator().manual_seed(7)
        L8 = wl.h8 * wl.w8
        cf0 = torch.randn(wl.P, L8, wl.C8, generator=g).to(dev)
        cf1 = (0.8 * cf0.cpu()[:, torch.randperm(L8, generator=g)] + 0.6 * torch.randn(wl.P, L8, wl.C8, generator=g)).to(dev)
        cms = _time(lambda: F.coarse_match_forward(cf0, cf1, 0.1), n=10)
        alg = 2 * 2.0 * wl.P * L8 * L8 * wl.C8                      # both directions, fp32-equivalent FLOPs
        rows['coarse_matching'] = {
            'ms_per_call': cms, 'alg_tflops': alg / cms / 1e9, 'issued_tf32_tflops': 3 * alg / cms / 1e9,
            'roofline': {'bound': 'tensor', 'achieved': alg / cms / 1e9, 'peak': tpeak, 'unit': 'TFLOP/s', 'frac': alg / cms / 1e9 / tpeak,
                         'note': 'fp32-accurate path = 3 TF32 MMAs per product at half the bf16 rate: ceiling = peak / 6'},
            'what': 'CoarseMatching next_idx/next_conf (both directions) at the 1/8 grid, tcgen05 kind::tf32 3-term split'}
        del cf0, cf1
    except Exception as e:      # noqa: BLE001
        rows['coarse_matching'] = {'error': str(e)[:300]}
    try:        # n4: the fine-window gather
        M = max(n_matches, 1)
        g = torch.Generator().manual_seed(9)
        s = wl.stages[-1]
        stride = max(wl.hf // s['h'], 1)
        ff = torch.randn(wl.P, 64, wl.hf, wl.wf, generator=g).to(dev)
        bi = torch.zeros(M, dtype=torch.int64, device=dev)
        ii = torch.randint(0, s['h'] * s['w'], (M,), generator=g).to(dev)
        t_g = _time(lambda: F.fine_window_gather(ff, bi, ii, s['w'], stride, 5))
        rows['fine_preprocess'] = {'ms_per_map_gather': t_g, 'matches': M,
                                   'what': 'CascadeFinePreprocess window crop of one fine map (gather kernel instead of F.unfold + select)'}
        del ff
    except Exception as e:      # noqa: BLE001
        rows['fine_preprocess'] = {'error': str(e)[:300]}
    return rows


def runner_gbps(e2e):
    """host->device bytes per second per GPU of the e2e run."""
    return e2e['h2d_bytes_per_step'] / (e2e['ms_per_step'] / 1e3) / 1e9


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
