"""The coarse-to-fine matching hot path of one CasMTR forward, as the sequence of module calls the
reference model makes between its backbone and its match list (SURVEY.md §3.1, §8d "units of work"):

    CasMTR-4c outdoor (reference src/model/cascade_model_stage3.py:139-178)
      12 x QTAttB.forward            1/8 grid, C=256, 8 heads, topks [32,16,8]   (6 layers x 2 maps)
       4 x CascadeQTAttB.forward     1/4 grid, C=128, 4 heads, 5x5 window -> K=100 (2 cross layers x 2 dirs)
       1 x CascadeMatching.forward   1/4 grid: 2 sparse correlations + softmax/argmax + NMS(5) + extraction
       1 x CascadeFineMatching.forward  [M,25,64] windows -> sub-pixel keypoints

Everything between those calls in the reference (1x1 convs, MLPs, self-attention blocks, the
backbone, F.unfold of the fine map) is out of scope, so the calls are fed with
synthetic feature maps of the right shapes (casmtr_b200/synth.py).  bench.py, the full-size GPU
tests and smoke() all drive the path through this one class, via the public module API.
"""
import os

import torch
import torch.nn as nn

from . import functional as F
from . import synth
from .cascade_matching import CascadeMatching
from .fine_matching import CascadeFineMatching
from .modules.quadtree_attention import CascadeQTAttB, QTAttB

MATCH_CFG_4C_OUTDOOR = {        # configs/model_configs/outdoor/loftr_ds_quadtree_cas_twins_large_stage3.py via get_match_config(., 0)
    'thr': 0.0101, 'test_thr': 0.2, 'pre_thr': [0.2], 'border_rm': 2, 'double_check': True,
    'train_pad_num_gt_min': 4096, 'match_type': 'softmax', 'dsmax_temperature': 1.0}
CAS_CFG_4C_OUTDOOR = {
    'propagation': 'window', 'dilated': 1, 'detector_mode': None, 'grid_size': 4,
    'post_config': {'method': 'maxpool_nms', 'window_size': 5, 'topk': None, 'rt': None, 'rd': None}}


class Workload:
    """Shapes of BASELINE.json configs[1] (CasMTR-4c outdoor) at a given square/rect image size."""

    def __init__(self, height=832, width=832, pairs=1, qt_calls=12, cas_calls=4, topks=(32, 16, 8)):
        assert height % 32 == 0 and width % 32 == 0, 'image size must be a multiple of 32 (1/8 grid with a 3-level pyramid)'
        self.H, self.W, self.B = height, width, pairs
        self.h8, self.w8 = height // 8, width // 8
        self.h4, self.w4 = height // 4, width // 4
        self.hf, self.wf = height // 2, width // 2
        self.qt_calls, self.cas_calls, self.topks = qt_calls, cas_calls, list(topks)
        self.C8, self.nh8, self.C4, self.nh4, self.Cf = 256, 8, 128, 4, 64
        self.window, self.fine_ww = 5, 25
        self.fine_cap = max(64, (self.h4 * self.w4 // 4)) * pairs         # windows pre-generated for FineMatching

    @property
    def name(self):
        return f'CasMTR-4c outdoor {self.H}x{self.W} batch={self.B} per GPU, coarse->1/4 cascade + NMS + fine'

    # ---- algorithmic (compulsory) bytes per CALL; fp32 features, int64 indices at the API edge
    def bytes_qtatt_call(self):
        L = [self.h8 * self.w8 // (4 ** i) for i in range(3)]
        return 4 * self.C8 * (3 * sum(L) + L[0]) * self.B

    def bytes_cascade_att_call(self):
        L = self.h4 * self.w4
        return (3 * L * self.C4 * 4 + (L // 4) * 25 * 2 * 8 + L * self.C4 * 4 + L * 100 * 8) * self.B

    def bytes_cascade_match_call(self):
        L = self.h4 * self.w4
        return (2 * L * self.C4 * 4 + 2 * L * 100 * 8 + L * 100 * 4 + 2 * L * 12) * self.B

    # ---- per-LAUNCH bytes of the individual kernels (DESIGN.md "kernels" table)
    def bytes_kernel(self, kind):
        L0, L1, L2 = [self.h8 * self.w8 // (4 ** i) for i in range(3)]
        C, nh, B = self.C8, self.nh8, self.B
        if kind == 'qt_fine_last':       # q,k,v,out at L0 + parent message at L1 + parent top-k list (int32)
            return B * (4 * C * (4 * L0 + L1) + 4 * L1 * nh * self.topks[1])
        if kind == 'qt_fine_mid':        # same at L1/L2, plus the emitted top-k (idx int32 + score fp32)
            return B * (4 * C * (4 * L1 + L2) + 4 * L2 * nh * self.topks[0] + 8 * L1 * nh * self.topks[1])
        if kind == 'qt_coarse':
            return B * (4 * C * 4 * L2 + 8 * L2 * nh * self.topks[0])
        if kind == 'cascade_att':
            return self.bytes_cascade_att_call()
        if kind == 'cascade_match':
            return self.bytes_cascade_match_call()
        return None


def make_host_inputs(wl, seed=1234, pin=False):
    """All inputs of one step (one batch of wl.B pairs) as CPU tensors, grouped per call."""
    g = torch.Generator().manual_seed(seed)
    B = wl.B
    maybe_pin = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    host = {'qt': [], 'cas': []}
    for i in range(wl.qt_calls):
        qs, ks, vs, wt = synth.qtatt_inputs(B, wl.C8, wl.h8, wl.w8, 3, seed=seed + 1 + i)
        host['qt'].append({'q': [maybe_pin(t) for t in qs], 'k': [maybe_pin(t) for t in ks], 'v': [maybe_pin(t) for t in vs],
                           'weight': wt})
    c = synth.cascade_inputs(B, wl.C4, wl.h4, wl.w4, seed=seed + 100, max_shift=8)
    for i in range(wl.cas_calls):
        rev = i % 2 == 1                               # call order per cross layer: 0->1 then 1->0
        q = (c['feat1'] if rev else c['feat0']) + 0.05 * torch.randn(B, wl.C4, wl.h4, wl.w4, generator=g)
        k = (c['feat0'] if rev else c['feat1']) + 0.05 * torch.randn(B, wl.C4, wl.h4, wl.w4, generator=g)
        v = torch.randn(B, wl.C4, wl.h4, wl.w4, generator=g)
        host['cas'].append({'q': maybe_pin(q), 'k': maybe_pin(k), 'v': maybe_pin(v),
                            'topk_pos': maybe_pin((c['topk_pos10'] if rev else c['topk_pos01']).contiguous())})
    host['match'] = {'feat0': maybe_pin(c['feat0'].flatten(2).transpose(1, 2).contiguous()),
                     'feat1': maybe_pin(c['feat1'].flatten(2).transpose(1, 2).contiguous()),
                     'pre_conf': maybe_pin(c['pre_conf01'].contiguous())}
    f0, f1 = synth.fine_inputs(wl.fine_cap, wl.fine_ww, wl.Cf, seed=seed + 200)
    host['fine'] = {'feat_f0': maybe_pin(f0), 'feat_f1': maybe_pin(f1)}
    return host


def tree_map(fn, x):
    if isinstance(x, torch.Tensor):
        return fn(x)
    if isinstance(x, dict):
        return {k: tree_map(fn, v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [tree_map(fn, v) for v in x]
    return x


def tree_bytes(x):
    n = [0]
    tree_map(lambda t: n.__setitem__(0, n[0] + t.numel() * t.element_size()), x)
    return n[0]


class HotPath(nn.Module):
    """The module sequence of one forward, built from the drop-in classes (reference constructors)."""

    def __init__(self, wl):
        super().__init__()
        self.wl = wl
        self.qt = nn.ModuleList([QTAttB(wl.nh8, wl.C8 // wl.nh8, scale=3, topks=wl.topks) for _ in range(wl.qt_calls)])
        self.cas = nn.ModuleList([CascadeQTAttB(wl.nh4, wl.C4 // wl.nh4, dilated=1) for _ in range(wl.cas_calls)])
        self.matching = CascadeMatching(MATCH_CFG_4C_OUTDOOR, CAS_CFG_4C_OUTDOOR, stage=1)
        self.fine = CascadeFineMatching('4c')
        self.eval()

    def load_level_weights(self, host):
        with torch.no_grad():
            for m, call in zip(self.qt, host['qt']):
                m.weight.copy_(call['weight'])

    def data_dict(self, dev_in):
        wl = self.wl
        return {'bs': wl.B, 'hw0_i': (wl.H, wl.W), 'hw1_i': (wl.H, wl.W),
                'hw0_8c': (wl.h8, wl.w8), 'hw1_8c': (wl.h8, wl.w8), 'hw0_4c': (wl.h4, wl.w4), 'hw1_4c': (wl.h4, wl.w4),
                'hw0_f': (wl.hf, wl.wf), 'hw1_f': (wl.hf, wl.wf),
                'stage_8c': {'next_conf_c01': dev_in['match']['pre_conf']}}

    # the individual calls, so that a host-fed runner can interleave copies with them
    def run_qt(self, i, call):
        return self.qt[i](call['q'], call['k'], call['v'])

    def run_cas(self, i, call):
        return self.cas[i](call['q'], call['k'], call['v'], call['topk_pos'], None)

    def run_match(self, dev_in, idx01, idx10):
        data = self.data_dict(dev_in)
        self.matching(dev_in['match']['feat0'], dev_in['match']['feat1'], idx01, idx10, data, level='4c', pre_level='8c')
        return data

    def run_fine(self, data, fine_in):
        M = data['stage_4c']['mconf'].shape[0]
        M = min(M, fine_in['feat_f0'].shape[0])
        self.fine(fine_in['feat_f0'][:M], fine_in['feat_f1'][:M], data)
        st = data['stage_4c']
        return {'b_ids': st['b_ids'], 'i_ids': st['i_ids'], 'j_ids': st['j_ids'], 'mconf': st['mconf'],
                'mkpts0': data['mkpts0_f'], 'mkpts1': data['mkpts1_f'], 'expec_f': data['expec_f']}

    def run_fine_deferred(self, data, fine_in):
        """Fine stage on the capacity-sized buffers of a defer_sync extraction: the match count stays on the device (the
        fine kernel reads it there), so nothing in the step waits for the host.  Returns capacity-sized arrays + 'count';
        trim_result() slices them once the host may read the count."""
        d = data['stage_4c']['_deferred']
        cap = min(d['b_ids'].shape[0], fine_in['feat_f0'].shape[0])
        scale = data['hw0_i'][0] / data['hw0_f'][0]
        expec, mk1 = F.fine_match_forward(fine_in['feat_f0'][:cap], fine_in['feat_f1'][:cap], d['mkpts1_c'][:cap], scale, count=d['count'])
        return {'b_ids': d['b_ids'][:cap], 'i_ids': d['i_ids'][:cap], 'j_ids': d['j_ids'][:cap], 'mconf': d['mconf'][:cap],
                'mkpts0': d['mkpts0_c'][:cap], 'mkpts1': mk1, 'expec_f': expec, 'count': d['count']}

    @torch.no_grad()
    def forward(self, dev_in, keep=None):
        """dev_in: make_host_inputs() moved to the GPU.  Returns the match list dict.  `keep`, if a dict, receives
        the intermediate outputs (messages, upsampled indices, stage dict) for parity tests."""
        idx = [None, None]
        for i, call in enumerate(dev_in['qt']):
            m = self.run_qt(i, call)
            if keep is not None:
                keep.setdefault('qt_msg', []).append(m)
        for i, call in enumerate(dev_in['cas']):
            m, up = self.run_cas(i, call)
            idx[i % 2] = up
            if keep is not None:
                keep.setdefault('cas_msg', []).append(m)
                keep.setdefault('cas_idx', []).append(up)
        data = self.run_match(dev_in, idx[0], idx[1])
        out = self.run_fine(data, dev_in['fine'])
        if keep is not None:
            keep['data'] = data
        return out


def trim_result(out):
    """Capacity-sized result of a sync-free step -> the match list (reads the device-side count: the one host sync)."""
    if 'count' not in out:
        return out
    M = min(int(out['count'].item()), out['b_ids'].shape[0])
    return {k: v[:M] for k, v in out.items() if k != 'count'}


class GraphRunner:
    """CUDA-graph replay of a step with the model's own concurrency: the two directions of a layer (calls 2i, 2i+1:
    feat0->feat1 and feat1->feat0, computed from the same inputs in the reference, src/model/modules/transformer.py:300)
    run on two streams forked/joined inside the graph.  Inputs are the static device buffers given at capture time.
    whole_step=False: the graph ends before the one host sync of the path (the match count); the fine stage runs eagerly
    after it.  whole_step=True: the fine stage is captured too, reading the match count on the device -- a step is ONE graph
    launch without any host synchronisation and returns capacity-sized buffers + 'count' (see trim_result)."""

    def __init__(self, hp, dev_in, two_streams=True, whole_step=False):
        self.hp, self.dev_in, self.whole_step = hp, dev_in, whole_step
        dev = dev_in['match']['feat0'].device
        hp.matching.defer_sync = True
        try:
            for _ in range(2):                      # warm-up: function attributes, tensor-map entry point, allocator
                self._body(None)
                hp.matching.finalize(self.data)
            torch.cuda.synchronize(dev)
            self.side = torch.cuda.Stream(dev) if two_streams else None
            self.graph = torch.cuda.CUDAGraph()
            # with two concurrent streams the other direction's kernel already fills a kernel's tail; early-scheduled
            # dependent CTAs would only hold SM resources it could use (measured: 2.45 ms vs 2.39 ms per step)
            prev = F.set_pdl(not two_streams)
            # same reasoning for the library's own side-stream overlap of the transposes (casmtr_set_overlap): with the other
            # direction already co-running it only adds contention (measured: 2.186 ms with, 2.178 ms without; eager single
            # stream: 3.14 ms with, 3.18 ms without; re-checked with the r02 kernels: 2.066 ms with, 2.058 ms without)
            prev_ov = F.set_overlap(not two_streams)
            # launch geometry is fixed at capture: tell the library that two calls run side by side (the dense coarsest level
            # then keeps its 32-row CTAs: 2.083 -> 2.066 ms per step)
            prev_cc = F.set_concurrency(2 if two_streams else 1)
            try:
                with torch.cuda.graph(self.graph):
                    self._body(self.side)
            finally:
                F.set_pdl(prev)
                F.set_overlap(prev_ov)
                F.set_concurrency(prev_cc)
            self.deferred = self.data['stage_4c']['_deferred']      # static buffers the graph writes on every replay
        finally:
            hp.matching.defer_sync = False

    def _pair(self, fn, n, calls, side):
        outs = [None] * n
        main = torch.cuda.current_stream()
        for i in range(0, n, 2):
            if side is not None and i + 1 < n:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    outs[i + 1] = fn(i + 1, calls[i + 1])
                outs[i] = fn(i, calls[i])
                main.wait_stream(side)
            else:
                outs[i] = fn(i, calls[i])
                if i + 1 < n:
                    outs[i + 1] = fn(i + 1, calls[i + 1])
        return outs

    def _body(self, side):
        hp, d = self.hp, self.dev_in
        self.qt_out = self._pair(hp.run_qt, len(d['qt']), d['qt'], side)
        cas = self._pair(hp.run_cas, len(d['cas']), d['cas'], side)
        idx = [None, None]
        for i, (_, up) in enumerate(cas):
            idx[i % 2] = up
        self.data = hp.run_match(d, idx[0], idx[1])
        if self.whole_step:
            self.out = hp.run_fine_deferred(self.data, d['fine'])

    @torch.no_grad()
    def step(self):
        self.graph.replay()
        if self.whole_step:
            return self.out                         # static buffers, rewritten by every replay
        self.data['stage_4c']['_deferred'] = self.deferred
        self.hp.matching.finalize(self.data)        # host sync: the match count
        return self.hp.run_fine(self.data, self.dev_in['fine'])


class HostFedRunner:
    """End-to-end driver: the step's inputs start in (pinned) HOST memory.  A copy stream uploads each call's inputs
    while the previous calls compute (per-call ready/consumed events), the match list is read back to the host."""

    def __init__(self, hp, host, device):
        self.hp, self.device = hp, device
        self.copy_stream = torch.cuda.Stream(device)
        self.groups = [('qt', i) for i in range(len(host['qt']))] + [('cas', i) for i in range(len(host['cas']))] + [('match', None)]
        # every call's inputs are packed into ONE pinned host block and one device block (the tensors the modules see are
        # views into the device block): one large H2D copy per call instead of ~10 small ones
        self.host, self.dev = {'qt': [], 'cas': []}, {'qt': [], 'cas': []}
        self.blocks = []
        for kind, i in self.groups:
            src = host[kind] if i is None else host[kind][i]
            hv, dv, hb, db = self._pack(src)
            self.blocks.append((hb, db))
            if i is None:
                self.host[kind], self.dev[kind] = hv, dv
            else:
                self.host[kind].append(hv)
                self.dev[kind].append(dv)
        self.host['fine'] = host['fine']
        self.fine_dev = {k: torch.empty_like(v, device=device) for k, v in host['fine'].items()}
        self.ready = [torch.cuda.Event() for _ in self.groups]
        self.consumed = [torch.cuda.Event() for _ in self.groups]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _pack(self, src):
        """-> (host views, device views, pinned host block, device block) for one call's input tree."""
        leaves = []
        tree_map(lambda t: leaves.append(t) or t, {k: v for k, v in src.items() if k != 'weight'})
        offs, total = [], 0
        for t in leaves:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        hb = torch.empty(total, dtype=torch.uint8).pin_memory()
        db = torch.empty(total, dtype=torch.uint8, device=self.device)
        it = iter(offs)

        def view(block):
            def f(t):
                o = next(it)
                return block[o:o + t.numel() * t.element_size()].view(t.dtype).reshape(t.shape)
            return f
        hv = tree_map(view(hb), {k: v for k, v in src.items() if k != 'weight'})
        it = iter(offs)
        dv = tree_map(view(db), {k: v for k, v in src.items() if k != 'weight'})
        flat_h, flat_s = [], []
        tree_map(lambda t: flat_h.append(t) or t, hv)
        for h, t in zip(flat_h, leaves):
            h.copy_(t)
        if 'weight' in src:
            hv['weight'] = src['weight']
            dv['weight'] = src['weight'].to(self.device)
        return hv, dv, hb, db

    @torch.no_grad()
    def step(self):
        main = torch.cuda.current_stream(self.device)
        cs = self.copy_stream
        h2d = 0
        with torch.cuda.stream(cs):
            for gi, g in enumerate(self.groups):
                cs.wait_event(self.consumed[gi])            # the previous step's consumer of this buffer is done
                hb, db = self.blocks[gi]
                db.copy_(hb, non_blocking=True)
                h2d += hb.numel()
                self.ready[gi].record(cs)
        idx = [None, None]
        for gi, (kind, i) in enumerate(self.groups):
            main.wait_event(self.ready[gi])
            if kind == 'qt':
                self.hp.run_qt(i, self.dev['qt'][i])
            elif kind == 'cas':
                _, up = self.hp.run_cas(i, self.dev['cas'][i])
                idx[i % 2] = up
            else:
                data = self.hp.run_match(self.dev, idx[0], idx[1])      # host sync inside: the match count
            self.consumed[gi].record(main)
        M = min(data['stage_4c']['mconf'].shape[0], self.fine_dev['feat_f0'].shape[0])
        for k in ('feat_f0', 'feat_f1'):                    # only the M windows the matches select
            self.fine_dev[k][:M].copy_(self.host['fine'][k][:M], non_blocking=True)
            h2d += M * self.host['fine'][k][0].numel() * 4
        out = self.hp.run_fine(data, self.fine_dev)
        res = {k: v.cpu() for k, v in out.items()}         # device -> host read of the result (synchronises)
        self.h2d_bytes = h2d
        self.d2h_bytes = tree_bytes(res)
        return res
