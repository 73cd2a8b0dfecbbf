"""The whole coarse-to-fine hot path through the oracle: the same call sequence casmtr_b200/pipeline.py drives on the GPU,
executed with the CPU restatements of this package on the same synthetic inputs.

TEST INFRASTRUCTURE / BASELINE ONLY -- see oracle/__init__.py.  Callers: tests/ (full-pipeline parity), and bench.py's
`--impl reference` arm and `cpu_baseline` / `gpu_*_baseline` legs, where it is the thing the CUDA path is compared WITH.

Sequence per step (reference src/model/cascade_model_stage3.py:139-178, cascade_model_stage4.py:139-206):
  per coarse layer   QuadtreeAttention.forward's pyramid (avg_pool2d, src/model/modules/quadtree_attention.py:86-89) + QTAttB.forward
  per cascade stage  get_window_warp_idx (transformer.py:416-440) [+ get_relative_pe :473-509] + CascadeQTAttB per cross layer,
                     CascadeMatching.forward + get_coarse_match
  last               CascadeFineMatching.forward
`impl='ref_kernels'` swaps the three gather ops for the reference's own CUDA extensions (oracle/ref_path.py, GPU tensors).
"""
import torch
import torch.nn.functional as tF

from . import cascade as ocas
from . import fine as ofine
from . import qtatt as oqt
from . import widen as owid


def _nchw(tok, h, w):
    """token-major [B,h*w,C] -> [B,C,h,w]."""
    B, L, C = tok.shape
    return tok.transpose(1, 2).reshape(B, C, h, w)


def _pyramid(x, levels=3):
    out = [x]
    for _ in range(levels - 1):
        out.append(tF.avg_pool2d(out[-1], kernel_size=2, stride=2))
    return out


def window_positions(idx, H, W, window=5):
    """get_window_warp_idx, 'window' propagation (transformer.py:416-433): next_idx [B,L] on an HxW grid -> [B,L,ww,2] (row, col)
    of the window around each match, shifted rigidly inside the grid."""
    r = window // 2
    cy = torch.div(idx, W, rounding_mode='trunc').clamp(r, H - 1 - r)
    cx = (idx % W).clamp(r, W - 1 - r)
    off = torch.arange(-r, r + 1, device=idx.device)
    oy, ox = torch.meshgrid(off, off, indexing='ij')
    return torch.stack([cy.unsqueeze(-1) + oy.reshape(-1), cx.unsqueeze(-1) + ox.reshape(-1)], dim=-1)


def run_step(wl, inp, impl='oracle', keep=None, sync=None, times=None):
    """inp: pipeline.make_host_inputs(wl) (CPU tensors, or the same tree on a CUDA device).  Returns the match list dict of
    the last stage + fine matching.  keep (dict) receives intermediates; times (dict) accumulates seconds per call kind."""
    import time
    qt_fn = lambda q, k, v, w: oqt.qtatt_b(q, k, v, w, wl.topks, wl.nh8)
    cas_fn = lambda q, k, v, pos, rp, nh: oqt.cascade_qtatt_b(q, k, v, pos, rp, nh)
    saved_score3d = ocas.ops.score3d
    if impl == 'ref_kernels':
        from . import ref_path
        qt_fn = lambda q, k, v, w: ref_path.qtatt_b(q, k, v, w, wl.topks, wl.nh8)
        plain = cas_fn
        cas_fn = lambda q, k, v, pos, rp, nh: ref_path.cascade_qtatt_b(q, k, v, pos, nh) if rp is None else plain(q, k, v, pos, rp, nh)
        ocas.ops.score3d = ref_path.score3d

    def clock():
        if sync is not None:
            sync()
        return time.perf_counter()

    def timed(kind, fn):
        t0 = clock()
        r = fn()
        if times is not None:
            times[kind] = times.get(kind, 0.0) + clock() - t0
        return r

    P = wl.P
    try:
        with torch.no_grad():
            for i, call in enumerate(inp['qt']):
                def one(call=call):
                    if wl.entry == 'tokens':
                        q, k, v = (_pyramid(_nchw(call[n], wl.h8, wl.w8)) for n in ('q', 'k', 'v'))
                    else:
                        q, k, v = call['q'], call['k'], call['v']
                    return qt_fn(q, k, v, call['weight'].to(q[0].device))
                m = timed('qtatt_b', one)
                if keep is not None:
                    keep.setdefault('qt_msg', []).append(m)
            hand = inp['hand']
            stage_out = {'8c': {'next_conf01': hand['pre_conf'], 'next_idx01': hand['next_idx'][:P], 'next_idx10': hand['next_idx'][P:]}}
            hw = {'8c': (wl.h8, wl.w8)}
            res = None
            for si, s in enumerate(wl.stages):
                h, w = s['h'], s['w']
                hw[s['level']] = (h, w)
                prev = stage_out['8c' if si == 0 else wl.stages[si - 1]['level']]
                nidx = torch.cat([prev['next_idx01'], prev['next_idx10']], 0)
                pos = timed('window_idx', lambda: window_positions(nidx, h // 2, w // 2, wl.window))
                rp = None
                if si == 0 and wl.cfg['relpe']:
                    tgt = torch.cat([stage_out['8c']['next_idx01'], stage_out['8c']['next_idx10']], 0)
                    rp = timed('relative_pe', lambda: owid.relative_pe(pos.cpu(), tgt.cpu(), inp['relpe']['w_table'].cpu(), inp['relpe']['h_table'].cpu(),
                                                                       wl.cfg['relpe']['LB'], (wl.h8, wl.w8), wl.w8, h).to(nidx.device))
                up = None
                for call in inp['stages'][si]['layers']:
                    def one(call=call):
                        if wl.entry == 'tokens':
                            q, k, v = (_nchw(call[n], h, w) for n in ('q', 'k', 'v'))
                        else:
                            q, k, v = call['q'], call['k'], call['v']
                        return cas_fn(q, k, v, pos, rp, s['nh'])
                    m, up = timed('cascade_qtatt_b', one)
                    if keep is not None:
                        keep.setdefault('cas_msg', []).append(m)
                        keep.setdefault('cas_idx', []).append(up)
                st = inp['stages'][si]
                pre_levels = s['pre_level'] if isinstance(s['pre_level'], list) else [s['pre_level']]

                def match():
                    o = ocas.cascade_match(st['feat0'], st['feat1'], up[:P], up[P:], None, None, s['match']['dsmax_temperature'])
                    post = s['cas']['post_config']
                    r = ocas.extract_matches(o['next_conf01'], o['next_idx01'], o['next_idx10'], (h, w), (h, w), (wl.H, wl.W),
                                             test_thr=s['match']['test_thr'], border_rm=s['match']['border_rm'],
                                             nms_window=post['window_size'] if post['method'] == 'maxpool_nms' else None,
                                             pre_confs=[(stage_out[p]['next_conf01'], hw[p][0], hw[p][1]) for p in pre_levels],
                                             pre_thrs=s['match']['pre_thr'], double_check=s['match']['double_check'])
                    return o, r
                o, res = timed('cascade_matching', match)
                stage_out[s['level']] = o
                if keep is not None:
                    keep.setdefault('stage', {})[s['level']] = dict(o, **res)
            M = min(res['mconf'].shape[0], inp['fine']['feat_f0'].shape[0])
            e, k1 = timed('fine_matching', lambda: ofine.fine_match(inp['fine']['feat_f0'][:M], inp['fine']['feat_f1'][:M],
                                                                    res['mkpts1_c'][:M].float(), wl.H / wl.hf))
    finally:
        ocas.ops.score3d = saved_score3d
    return {'b_ids': res['b_ids'], 'i_ids': res['i_ids'], 'j_ids': res['j_ids'], 'mconf': res['mconf'],
            'mkpts0': res['mkpts0_c'].float(), 'mkpts1': k1, 'expec_f': e}
