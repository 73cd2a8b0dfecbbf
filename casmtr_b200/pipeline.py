"""The coarse-to-fine matching hot path of one CasMTR forward, as the sequence of module calls the
reference model makes between its backbone and its match list (SURVEY.md §3.1, §8d "units of work"):

    CasMTR-4c outdoor (reference src/model/cascade_model_stage3.py:139-178)                     config '4c'
      12 x QTAttB.forward            1/8 grid, C=256, 8 heads, topks [32,16,8]   (6 layers x 2 directions)
       4 x CascadeQTAttB.forward     1/4 grid, C=128, 4 heads, 5x5 window -> K=100 (2 cross layers x 2 directions)
       1 x CascadeMatching.forward   1/4 grid: 2 sparse correlations + softmax/argmax + NMS(5) + extraction
       1 x CascadeFineMatching.forward  [M,25,64] windows -> sub-pixel keypoints
    CasMTR-2c outdoor (src/model/cascade_model_stage4.py:139-206)                                config '2c'
      the same with threshold-only extraction at 1/4 (..._stage4.py:36), then a second cascade stage at 1/2
      (C=64, 2 heads): 4 x CascadeQTAttB + CascadeMatching with NMS(5), pre_level ['8c','4c'], windows taken from the
      1/4 stage's OWN next_idx (the real data flow of :178-195)
    CasMTR-4c indoor (configs/model_configs/indoor/loftr_ds_quadtree_cas_stage3.py:10-53)         config 'indoor'
      16 x QTAttB (8 layers, topks [32,16,16]), 4 x CascadeQTAttB with the relative position bias, threshold-only, border 1

The two directions of a layer (feat0 -> feat1 and feat1 -> feat0) are computed from the same inputs with the same weights
(src/model/modules/transformer.py:296-303, :545), so they are STACKED on the batch dimension: a batch of P pairs is one call with
B = 2P per layer, i.e. 6 (8) QTAtt launches-sets and 2 cascade launch-sets per stage instead of 12 (16) and 4.  That is the
reference's own module API with a doubled batch, not a different algorithm; `entry` selects which boundary is fed:

    entry='tokens'  QuadtreeAttention.forward's inner boundary (src/model/modules/quadtree_attention.py:81-99): token-major
                    level-0 q / k / v [2P, H*W, C] (what the q/k/v projections produce when applied as linear layers); the
                    avg-pool pyramid is built inside the library (casmtr_qtatt_tokens_fwd) -- no NCHW map, no transpose
    entry='nchw'    QTAttB.forward's boundary: the NCHW pyramid lists of the reference API (casmtr_qtatt_fwd)

Everything between those calls in the reference (1x1 convs, MLPs, self-attention blocks, the backbone, F.unfold of the
fine map) is out of scope, so the calls are fed with synthetic feature maps of the right shapes
(casmtr_b200/synth.py).  bench.py, the full-size GPU tests and smoke() all drive the path through this one class.
"""
import torch
import torch.nn as nn
import torch.nn.functional as tF

from . import functional as F
from . import synth
from .cascade_matching import CascadeMatching
from .fine_matching import CascadeFineMatching
from .modules.quadtree_attention import CascadeQTAttB, QTAttB


def _match_cfg(**kw):
    base = {'thr': 0.0101, 'test_thr': 0.2, 'pre_thr': [0.2], 'border_rm': 2, 'double_check': True,
            'train_pad_num_gt_min': 4096, 'match_type': 'softmax', 'dsmax_temperature': 1.0}
    base.update(kw)
    return base


def _cas_cfg(nms):
    return {'propagation': 'window', 'dilated': 1, 'detector_mode': None, 'grid_size': 4,
            'post_config': {'method': 'maxpool_nms' if nms else None, 'window_size': 5, 'topk': None, 'rt': None, 'rd': None}}


# plain-dict versions of the reference's yacs configs
CONFIGS = {
    # configs/model_configs/outdoor/loftr_ds_quadtree_cas_twins_large_stage3.py via get_match_config(., 0)
    '4c': {'title': 'CasMTR-4c outdoor', 'qt_layers': 6, 'topks': (32, 16, 8), 'relpe': None, 'fine_level': '4c',
           'stages': [{'level': '4c', 'down': 4, 'C': 128, 'nh': 4, 'cross': 2, 'pre_level': '8c',
                       'match': _match_cfg(), 'cas': _cas_cfg(True)}]},
    # ..._stage4.py: 1/4 stage threshold-only with border 1 (:36), 1/2 stage NMS 5 with border 2 and two previous-stage gates
    '2c': {'title': 'CasMTR-2c outdoor', 'qt_layers': 6, 'topks': (32, 16, 8), 'relpe': None, 'fine_level': '2c',
           'stages': [{'level': '4c', 'down': 4, 'C': 128, 'nh': 4, 'cross': 2, 'pre_level': '8c',
                       'match': _match_cfg(border_rm=1), 'cas': _cas_cfg(False)},
                      {'level': '2c', 'down': 2, 'C': 64, 'nh': 2, 'cross': 2, 'pre_level': ['8c', '4c'],
                       'match': _match_cfg(pre_thr=[0.2, 0.2]), 'cas': _cas_cfg(True)}]},
    # configs/model_configs/indoor/loftr_ds_quadtree_cas_stage3.py: 8 coarse layers (default.py:33), topks [32,16,16],
    # relative PE at 1/4 (LB = 10, tables [22, nhead]), no NMS (:30), test_thr 0.1 (:48), border 1 (:49)
    'indoor': {'title': 'CasMTR-4c indoor', 'qt_layers': 8, 'topks': (32, 16, 16), 'relpe': {'LB': 10, 'n_emb': 22}, 'fine_level': '4c',
               'stages': [{'level': '4c', 'down': 4, 'C': 128, 'nh': 4, 'cross': 2, 'pre_level': '8c',
                           'match': _match_cfg(thr=0.0, pre_thr=[0.2, 0.1], test_thr=0.1, border_rm=1), 'cas': _cas_cfg(False)}]},
}


class Workload:
    """Shapes of one BASELINE.json config at a given image size; `pairs` image pairs per step (per GPU)."""

    def __init__(self, height=832, width=832, pairs=1, config='4c', qt_layers=None, entry='tokens'):
        assert height % 32 == 0 and width % 32 == 0, 'image size must be a multiple of 32 (1/8 grid with a 3-level pyramid)'
        assert entry in ('tokens', 'nchw')
        cfg = CONFIGS[config]
        self.config, self.cfg, self.entry = config, cfg, entry
        self.H, self.W, self.P = height, width, pairs
        self.B = 2 * pairs                                  # batch of every attention call: both directions of every pair
        self.h8, self.w8 = height // 8, width // 8
        self.hf, self.wf = height // 2, width // 2
        self.qt_layers = cfg['qt_layers'] if qt_layers is None else qt_layers
        self.topks = list(cfg['topks'])
        self.C8, self.nh8, self.Cf = 256, 8, 64
        self.stages = []
        for st in cfg['stages']:
            s = dict(st)
            s['h'], s['w'] = height // st['down'], width // st['down']
            self.stages.append(s)
        self.window, self.fine_ww = 5, 25
        last = self.stages[-1]
        # windows pre-generated for FineMatching: a 5x5 NMS keeps at most one token in 9 (in practice ~1 in 40); threshold-only
        # extraction (indoor config) can keep every token
        nms = last['cas']['post_config']['method'] == 'maxpool_nms'
        self.fine_cap = max(64, last['h'] * last['w'] // (4 if nms else 1)) * pairs

    # call counts per pair in the reference's (un-stacked) terms
    @property
    def calls_per_pair(self):
        d = {'QTAttB': 2 * self.qt_layers}
        for s in self.stages:
            d[f"CascadeQTAttB@{s['level']}"] = 2 * s['cross']
            d[f"CascadeMatching@{s['level']}"] = 1
        d['CascadeFineMatching'] = 1
        return d

    @property
    def name(self):
        return (f"{self.cfg['title']} {self.H}x{self.W} batch={self.P} per GPU, coarse->" +
                '->'.join('1/%d' % s['down'] for s in self.stages) + ' cascade + NMS + fine')

    # ---- algorithmic (compulsory) bytes; fp32 features, int64 indices at the API edge.  Per reference CALL
    # (one direction, one pair); a stacked launch processes self.B of them.
    def bytes_qtatt_call(self):
        L = [self.h8 * self.w8 // (4 ** i) for i in range(3)]
        return 4 * self.C8 * (3 * sum(L) + L[0])

    def flops_qtatt_call(self):
        L0 = self.h8 * self.w8
        S, L1 = L0 // 16, L0 // 4
        return 4.0 * S * S * self.C8 + 4.0 * self.C8 * (L1 * 4 * self.topks[0] + L0 * 4 * self.topks[1])

    def bytes_cascade_att_call(self, s):
        L = s['h'] * s['w']
        return 3 * L * s['C'] * 4 + (L // 4) * 25 * 2 * 8 + L * s['C'] * 4 + L * 100 * 8

    def bytes_cascade_match_call(self, s):              # both directions of one pair
        L = s['h'] * s['w']
        return 2 * L * s['C'] * 4 + 2 * L * 100 * 8 + L * 100 * 4 + 2 * L * 12

    # ---- per-LAUNCH bytes of the individual kernels (DESIGN.md "kernels" table); a launch covers self.B calls (self.P pairs)
    def bytes_kernel(self, kind):
        L0, L1, L2 = [self.h8 * self.w8 // (4 ** i) for i in range(3)]
        C, nh, B = self.C8, self.nh8, self.B
        s0 = self.stages[0]
        if kind == 'qt_fine_last':       # q,k,v,out at L0 + parent message at L1 + parent top-k list (int32)
            return B * (4 * C * (4 * L0 + L1) + 4 * L1 * nh * self.topks[1])
        if kind == 'qt_fine_mid':        # same at L1/L2, plus the emitted top-k (idx int32 + score fp32)
            return B * (4 * C * (4 * L1 + L2) + 4 * L2 * nh * self.topks[0] + 8 * L1 * nh * self.topks[1])
        if kind == 'qt_coarse':
            return B * (4 * C * 4 * L2 + 8 * L2 * nh * self.topks[0])
        if kind == 'cascade_att':        # mean over the stages' launches
            return B * sum(self.bytes_cascade_att_call(s) * s['cross'] for s in self.stages) // sum(s['cross'] for s in self.stages)
        if kind == 'cascade_match':
            return self.P * sum(self.bytes_cascade_match_call(s) for s in self.stages) // len(self.stages)
        if kind == 'layout' and self.entry == 'tokens':
            # pooling pass of a QTAttB launch set: q, k, v read at L0; levels 1 and 2 written; plus the tensor-core level's operands
            # (Q_lo, K_lo fp32 at L2, V^T as an fp16 pair padded to 8 keys)
            Sp = (L2 + 7) // 8 * 8
            return B * C * 4 * (3 * L0 + 3 * L1 + 3 * L2 + 2 * L2 + Sp)
        del s0
        return None


def _tokens(x):
    """[B,C,h,w] -> token-major [B,h*w,C] contiguous."""
    return x.flatten(2).transpose(1, 2).contiguous()


def make_host_inputs(wl, seed=1234, pin=False):
    """All inputs of one step (one batch of wl.P pairs, directions stacked: rows [0,P) are image 0 -> image 1, rows [P,2P) the
    other way) as CPU tensors, grouped per call."""
    g = torch.Generator().manual_seed(seed)
    P = wl.P
    maybe_pin = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    host = {'qt': [], 'stages': []}
    for i in range(wl.qt_layers):
        qs, ks, vs, wt = synth.qtatt_inputs(wl.B, wl.C8, wl.h8, wl.w8, 3, seed=seed + 1 + i)
        if wl.entry == 'tokens':
            call = {'q': maybe_pin(_tokens(qs[0])), 'k': maybe_pin(_tokens(ks[0])), 'v': maybe_pin(_tokens(vs[0])), 'weight': wt}
        else:
            call = {'q': [maybe_pin(t) for t in qs], 'k': [maybe_pin(t) for t in ks], 'v': [maybe_pin(t) for t in vs], 'weight': wt}
        host['qt'].append(call)
    shifts = None
    for si, s in enumerate(wl.stages):
        # structured features: image 1 is image 0 rolled by a per-pair shift (+ noise), the same physical shift at every stage
        c = synth.cascade_inputs(P, s['C'], s['h'], s['w'], seed=seed + 100 + 10 * si, max_shift=8, shifts=shifts)
        shifts = c['shifts'] * 2
        layers = []
        for _ in range(s['cross']):
            f0 = c['feat0'] + 0.05 * torch.randn(P, s['C'], s['h'], s['w'], generator=g)
            f1 = c['feat1'] + 0.05 * torch.randn(P, s['C'], s['h'], s['w'], generator=g)
            q = torch.cat([f0, f1], 0)                     # direction 0: queries = image 0; direction 1: queries = image 1
            k = torch.cat([f1, f0], 0)
            v = torch.randn(2 * P, s['C'], s['h'], s['w'], generator=g)
            if wl.entry == 'tokens':
                q, k, v = _tokens(q), _tokens(k), _tokens(v)
            layers.append({'q': maybe_pin(q), 'k': maybe_pin(k), 'v': maybe_pin(v)})
        if si == 0:
            # what the 1/8 stage hands over (CoarseMatching, outside the starred path): its matches and confidences
            host['hand'] = {'next_idx': maybe_pin(torch.cat([c['next_idx01'], c['next_idx10']], 0).contiguous()),      # [2P, L/4]
                            'pre_conf': maybe_pin(c['pre_conf01'].contiguous())}
        host['stages'].append({'layers': layers, 'feat0': maybe_pin(_tokens(c['feat0'])), 'feat1': maybe_pin(_tokens(c['feat1']))})
    if wl.cfg['relpe']:
        n_emb = wl.cfg['relpe']['n_emb']
        nh = wl.stages[0]['nh']
        host['relpe'] = {'w_table': 0.1 * torch.randn(n_emb, nh, generator=g), 'h_table': 0.1 * torch.randn(n_emb, nh, generator=g)}
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

    def __init__(self, wl, qt_flags=0):
        super().__init__()
        self.wl = wl
        self.qt_flags = qt_flags
        self.qt = nn.ModuleList([QTAttB(wl.nh8, wl.C8 // wl.nh8, scale=3, topks=wl.topks) for _ in range(wl.qt_layers)])
        self.cas = nn.ModuleList([nn.ModuleList([CascadeQTAttB(s['nh'], s['C'] // s['nh'], dilated=1) for _ in range(s['cross'])])
                                  for s in wl.stages])
        self.matching = nn.ModuleList([CascadeMatching(s['match'], s['cas'], stage=i + 1) for i, s in enumerate(wl.stages)])
        self.fine = CascadeFineMatching(wl.cfg['fine_level'])
        self.eval()

    def load_level_weights(self, host):
        with torch.no_grad():
            for m, call in zip(self.qt, host['qt']):
                m.weight.copy_(call['weight'])

    def data_dict(self, dev_in):
        wl = self.wl
        d = {'bs': wl.P, 'hw0_i': (wl.H, wl.W), 'hw1_i': (wl.H, wl.W), 'hw0_8c': (wl.h8, wl.w8), 'hw1_8c': (wl.h8, wl.w8),
             'hw0_f': (wl.hf, wl.wf), 'hw1_f': (wl.hf, wl.wf)}
        for s in wl.stages:
            d[f"hw0_{s['level']}"] = d[f"hw1_{s['level']}"] = (s['h'], s['w'])
        hand = dev_in['hand']
        d['stage_8c'] = {'next_conf_c01': hand['pre_conf'], 'next_idx_c01': hand['next_idx'][:wl.P], 'next_idx_c10': hand['next_idx'][wl.P:]}
        return d

    # the individual calls, so that a host-fed runner can interleave copies with them
    def run_qt(self, i, call):
        wl = self.wl
        if wl.entry == 'tokens':
            return F.qtatt_tokens_forward(call['q'], call['k'], call['v'], (wl.h8, wl.w8), (wl.h8, wl.w8), wl.topks, wl.nh8,
                                          weight=self.qt[i].weight, flags=self.qt_flags)
        if self.qt_flags:
            return F.qtatt_forward(call['q'], call['k'], call['v'], wl.topks, wl.nh8, weight=self.qt[i].weight, flags=self.qt_flags)
        return self.qt[i](call['q'], call['k'], call['v'])

    def stage_windows(self, si, dev_in, data):
        """next_idx [2P, L/4] whose 5x5 windows the cross layers of stage si attend to (get_window_warp_idx is fused into the
        kernels): stage 0 takes the 1/8 matches it is given, later stages the previous stage's own next_idx (stage4.py:178-186)."""
        if si == 0:
            return dev_in['hand']['next_idx']
        prev = data[f"stage_{self.wl.stages[si - 1]['level']}"]
        return torch.cat([prev['next_idx_c01'], prev['next_idx_c10']], 0)

    def run_cas(self, si, li, call, next_idx, pe=None):
        s = self.wl.stages[si]
        if self.wl.entry == 'tokens':
            return F.cascade_qtatt_forward(call['q'], call['k'], call['v'], next_idx, pe, s['nh'], hw_q=(s['h'], s['w']), hw_k=(s['h'], s['w']))
        return self.cas[si][li](call['q'], call['k'], call['v'], next_idx, pe)

    def relpe(self, dev_in, data):
        """Fused relative position bias of the indoor config: both directions stacked (tgt = next_idx_c01 rows, then c10 rows)."""
        wl = self.wl
        if not wl.cfg['relpe']:
            return None
        tgt = torch.cat([data['stage_8c']['next_idx_c01'], data['stage_8c']['next_idx_c10']], 0).contiguous()
        return F.RelativePE(dev_in['relpe']['w_table'], dev_in['relpe']['h_table'], wl.cfg['relpe']['LB'], tgt, (wl.h8, wl.w8), wl.w8)

    def run_match(self, si, dev_in, data, up):
        s = self.wl.stages[si]
        P = self.wl.P
        st = dev_in['stages'][si]
        self.matching[si](st['feat0'], st['feat1'], up[:P], up[P:], data, level=s['level'], pre_level=s['pre_level'])

    def run_stage(self, si, dev_in, data, keep=None):
        nidx = self.stage_windows(si, dev_in, data)
        pe = self.relpe(dev_in, data) if si == 0 else None
        up = None
        for li, call in enumerate(dev_in['stages'][si]['layers']):
            m, up = self.run_cas(si, li, call, nidx, pe)
            if keep is not None:
                keep.setdefault('cas_msg', []).append(m)
                keep.setdefault('cas_idx', []).append(up)
                keep.setdefault('cas_nidx', []).append(nidx)
        self.run_match(si, dev_in, data, up)

    def run_fine(self, data, fine_in):
        lvl = self.wl.cfg['fine_level']
        M = data[f'stage_{lvl}']['mconf'].shape[0]
        M = min(M, fine_in['feat_f0'].shape[0])
        self.fine(fine_in['feat_f0'][:M], fine_in['feat_f1'][:M], data)
        st = data[f'stage_{lvl}']
        return {'b_ids': st['b_ids'], 'i_ids': st['i_ids'], 'j_ids': st['j_ids'], 'mconf': st['mconf'],
                'mkpts0': data['mkpts0_f'], 'mkpts1': data['mkpts1_f'], 'expec_f': data['expec_f']}

    def run_fine_deferred(self, data, fine_in):
        """Fine stage on the capacity-sized buffers of a defer_sync extraction: the match count stays on the device (the
        fine kernel reads it there), so nothing in the step waits for the host.  Returns capacity-sized arrays + 'count';
        trim_result() slices them once the host may read the count."""
        d = data[f"stage_{self.wl.cfg['fine_level']}"]['_deferred']
        cap = min(d['b_ids'].shape[0], fine_in['feat_f0'].shape[0])
        scale = data['hw0_i'][0] / data['hw0_f'][0]
        expec, mk1 = F.fine_match_forward(fine_in['feat_f0'][:cap], fine_in['feat_f1'][:cap], d['mkpts1_c'][:cap], scale, count=d['count'])
        return {'b_ids': d['b_ids'][:cap], 'i_ids': d['i_ids'][:cap], 'j_ids': d['j_ids'][:cap], 'mconf': d['mconf'][:cap],
                'mkpts0': d['mkpts0_c'][:cap], 'mkpts1': mk1, 'expec_f': expec, 'count': d['count']}

    @torch.no_grad()
    def forward(self, dev_in, keep=None):
        """dev_in: make_host_inputs() moved to the GPU.  Returns the match list dict.  `keep`, if a dict, receives
        the intermediate outputs (messages, upsampled indices, stage dict) for parity tests."""
        for i, call in enumerate(dev_in['qt']):
            m = self.run_qt(i, call)
            if keep is not None:
                keep.setdefault('qt_msg', []).append(m)
        data = self.data_dict(dev_in)
        for si in range(len(self.wl.stages)):
            self.run_stage(si, dev_in, data, keep)
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
    """CUDA-graph replay of a whole step on ONE stream: every layer's two directions are already one launch (stacked batch),
    so there is nothing to fork.  The fine stage is captured too, reading the match count on the device -- a step is ONE graph
    launch without any host synchronisation and returns capacity-sized buffers + 'count' (see trim_result).
    Inputs are the static device buffers given at capture time."""

    def __init__(self, hp, dev_in, pdl=True):
        self.hp, self.dev_in = hp, dev_in
        dev = dev_in['fine']['feat_f0'].device
        for m in hp.matching:
            m.defer_sync = True
        try:
            for _ in range(2):                      # warm-up: function attributes, tensor-map entry point, allocator
                self._body()
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            prev = F.set_pdl(pdl)
            try:
                with torch.cuda.graph(self.graph):
                    self._body()
            finally:
                F.set_pdl(prev)
        finally:
            for m in hp.matching:
                m.defer_sync = False

    def _body(self):
        hp, d = self.hp, self.dev_in
        self.qt_out = [hp.run_qt(i, call) for i, call in enumerate(d['qt'])]
        self.data = hp.data_dict(d)
        for si in range(len(hp.wl.stages)):
            hp.run_stage(si, d, self.data)
        self.out = hp.run_fine_deferred(self.data, d['fine'])

    @torch.no_grad()
    def step(self):
        self.graph.replay()
        return self.out                             # static buffers, rewritten by every replay


class HostFedRunner:
    """End-to-end driver: the step's inputs start in (pinned) HOST memory.  A copy stream uploads each call's inputs
    while the previous calls compute (per-call ready/consumed events, two device buffers per call so the copy of step n+1
    overlaps the compute of step n), the match list is read back to the host."""

    def __init__(self, hp, host, device, buffers=2):
        self.hp, self.device = hp, device
        self.copy_stream = torch.cuda.Stream(device)
        wl = hp.wl
        self.groups = [('hand', None, None)] + [('qt', i, None) for i in range(wl.qt_layers)]
        for si, s in enumerate(wl.stages):
            self.groups += [('cas', si, li) for li in range(s['cross'])] + [('match', si, None)]
        # every call's inputs are packed into ONE pinned host block and `buffers` device blocks (the tensors the modules see
        # are views into a device block): one large H2D copy per call instead of ~10 small ones
        self.nbuf = buffers
        self.blocks, self.views = [], []
        for kind, a, b in self.groups:
            src = self._src(host, kind, a, b)
            hb, dvs, dbs = self._pack(src, buffers)
            self.blocks.append((hb, dbs))
            self.views.append(dvs)
        self.host_fine = host['fine']
        self.fine_dev = {k: torch.empty_like(v, device=device) for k, v in host['fine'].items()}
        self.relpe_dev = tree_map(lambda t: t.to(device), host['relpe']) if 'relpe' in host else None
        self.ready = [[torch.cuda.Event() for _ in self.groups] for _ in range(buffers)]
        self.consumed = [[torch.cuda.Event() for _ in self.groups] for _ in range(buffers)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.n = 0
        self._prefetched = False

    @staticmethod
    def _src(host, kind, a, b):
        if kind == 'hand':
            return host['hand']
        if kind == 'qt':
            return host['qt'][a]
        st = host['stages'][a]
        if kind == 'cas':
            return st['layers'][b]
        return {k: v for k, v in st.items() if k != 'layers'}

    def _pack(self, src, buffers):
        """-> (pinned host block, [device views per buffer], [device blocks])."""
        tree = {k: v for k, v in src.items() if k != 'weight'}
        leaves = []
        tree_map(lambda t: leaves.append(t) or t, tree)
        offs, total = [], 0
        for t in leaves:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        hb = torch.empty(total, dtype=torch.uint8).pin_memory()

        def view(block):
            it = iter(offs)

            def f(t):
                o = next(it)
                return block[o:o + t.numel() * t.element_size()].view(t.dtype).reshape(t.shape)
            return tree_map(f, tree)
        flat_h = []
        tree_map(lambda t: flat_h.append(t) or t, view(hb))
        for h, t in zip(flat_h, leaves):
            h.copy_(t)
        dbs = [torch.empty(total, dtype=torch.uint8, device=self.device) for _ in range(buffers)]
        dvs = [view(db) for db in dbs]
        return hb, dvs, dbs

    def _upload(self, step):
        """Enqueue the H2D copies of step `step` on the copy stream (buffer step % nbuf)."""
        cs, bi = self.copy_stream, step % self.nbuf
        n = 0
        with torch.cuda.stream(cs):
            for gi in range(len(self.groups)):
                cs.wait_event(self.consumed[bi][gi])        # the previous user of this device buffer is done
                hb, dbs = self.blocks[gi]
                dbs[bi].copy_(hb, non_blocking=True)
                n += hb.numel()
                self.ready[bi][gi].record(cs)
        return n

    @torch.no_grad()
    def step(self):
        hp, wl = self.hp, self.hp.wl
        main = torch.cuda.current_stream(self.device)
        bi = self.n % self.nbuf
        if not self._prefetched:
            self._upload(self.n)
        h2d = sum(hb.numel() for hb, _ in self.blocks)
        if self.nbuf > 1:                                   # next step's inputs travel while this step computes
            self._upload(self.n + 1)
            self._prefetched = True
        dev_in = {'stages': [{} for _ in wl.stages]}
        data = None
        up = None
        nidx = pe = None
        for gi, (kind, a, b) in enumerate(self.groups):
            main.wait_event(self.ready[bi][gi])
            v = self.views[gi][bi]
            if kind == 'hand':
                dev_in['hand'] = v
                if self.relpe_dev is not None:
                    dev_in['relpe'] = self.relpe_dev
                data = hp.data_dict(dev_in)
                pe = hp.relpe(dev_in, data)
            elif kind == 'qt':
                hp.run_qt(a, v)
            elif kind == 'cas':
                if b == 0:
                    nidx = hp.stage_windows(a, dev_in, data)
                _, up = hp.run_cas(a, b, v, nidx, pe if a == 0 else None)
            else:
                dev_in['stages'][a] = v
                hp.run_match(a, dev_in, data, up)           # host sync inside: the match count
            self.consumed[bi][gi].record(main)
        lvl = wl.cfg['fine_level']
        M = min(data[f'stage_{lvl}']['mconf'].shape[0], self.fine_dev['feat_f0'].shape[0])
        for k in ('feat_f0', 'feat_f1'):                    # only the M windows the matches select
            self.fine_dev[k][:M].copy_(self.host_fine[k][:M], non_blocking=True)
            h2d += M * self.host_fine[k][0].numel() * 4
        out = hp.run_fine(data, self.fine_dev)
        res = {k: v.cpu() for k, v in out.items()}         # device -> host read of the result (synchronises)
        self.h2d_bytes = h2d
        self.d2h_bytes = tree_bytes(res)
        self.n += 1
        return res
