"""Generate the committed golden vectors by running the UNMODIFIED reference Python modules
(imported from /root/reference through oracle/refload.py) on seeded synthetic inputs.

    python tests/golden/make_golden.py          # only works where /root/reference exists

The reference ships no golden vectors of its own, so these files -- outputs of the
reference itself -- are what pins the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_golden.py).  Inputs are stored next to the outputs so nothing depends on RNG
reproducibility across torch versions.  int64 tensors are stored as int32 to keep the files small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from casmtr_b200 import synth  # noqa: E402
from oracle import refload  # noqa: E402


def _np(t):
    t = t.detach().cpu()
    if t.dtype == torch.int64:
        return t.to(torch.int32).numpy()
    if t.dtype == torch.bool:
        return t.to(torch.uint8).numpy()
    return t.numpy()


def save(name, **tensors):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **{k: _np(v) for k, v in tensors.items() if v is not None})
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB')


class Tap:
    """Records what the reference's process_*_level methods return, without touching its code."""

    def __init__(self, mod):
        self.idx, self.score = [], []
        for meth in ('process_coarse_level', 'process_fine_level'):
            orig = getattr(mod, meth)

            def wrapped(*a, _o=orig, **k):
                r = _o(*a, **k)
                self.score.append(r[2])
                self.idx.append(r[3])
                return r
            setattr(mod, meth, wrapped)


def gen_qtatt(ref):
    nh, C, h, w, topks = 2, 64, 16, 24, [8, 4, 4]
    qs, ks, vs, wt = synth.qtatt_inputs(1, C, h, w, 3, seed=101)
    mb = ref.QTAttB(nh, C // nh, scale=3, topks=topks)
    mb.weight.data.copy_(wt)
    tap = Tap(mb)
    with torch.no_grad():
        out_b = mb(qs, ks, vs)
    ma = ref.QTAttA(nh, C // nh, topks=topks)
    tap_a = Tap(ma)
    with torch.no_grad():
        out_a = ma(qs, ks, vs)
    t = {f'q{i}': qs[i] for i in range(3)}
    t.update({f'k{i}': ks[i] for i in range(3)})
    t.update({f'v{i}': vs[i] for i in range(3)})
    save('qtatt', weight=wt, topks=torch.tensor(topks), nhead=torch.tensor(nh), out_b=out_b, out_a=out_a,
         b_idx0=tap.idx[0], b_idx1=tap.idx[1], a_idx0=tap_a.idx[0], a_idx1=tap_a.idx[1],
         a_score0=tap_a.score[0], a_score1=tap_a.score[1], **t)


def gen_cascade_qtatt(ref):
    nh, C, h, w = 2, 64, 16, 16
    d = synth.cascade_inputs(2, C, h, w, seed=102)
    g = torch.Generator().manual_seed(7)
    v = torch.randn(2, C, h, w, generator=g)
    rp = torch.randn(2, nh, h * w, 100, generator=g)
    m = ref.CascadeQTAttB(nh, C // nh, dilated=1)
    with torch.no_grad():
        msg, up = m(d['feat0'], d['feat1'], v, d['topk_pos01'], None)
        msg_rp, _ = m(d['feat0'], d['feat1'], v, d['topk_pos01'], rp)
    save('cascade_qtatt', query=d['feat0'], key=d['feat1'], value=v, topk_pos=d['topk_pos01'], rel_pos=rp,
         nhead=torch.tensor(nh), message=msg, message_rel=msg_rp, upsampled_idx=up)


def gen_cascade_match(ref):
    B, C, h, w = 2, 32, 16, 16
    hp, wp = h // 2, w // 2
    cases = {}
    for tag, pad, nms, thr in (('nms', False, True, 0.2), ('pad', True, True, 0.2), ('thr', False, False, 0.2), ('empty', False, True, 2.0)):
        d = synth.cascade_inputs(B, C, h, w, seed=103, pad=pad)
        m = ref.CascadeQTAttB(1, C, dilated=1)   # only used for its window-index expansion (upsampled_idx)
        with torch.no_grad():
            z = torch.zeros(B, C, h, w)
            idx01 = m(z, z, z, d['topk_pos01'], None)[1]
            idx10 = m(z, z, z, d['topk_pos10'], None)[1]
        f0 = d['feat0'].flatten(2).transpose(1, 2).contiguous()
        f1 = d['feat1'].flatten(2).transpose(1, 2).contiguous()
        g = torch.Generator().manual_seed(8)
        data = {'hw0_i': (h * 4, w * 4), 'hw1_i': (h * 4, w * 4), 'hw0_4c': (h, w), 'hw1_4c': (h, w),
                'hw0_8c': (hp, wp), 'hw1_8c': (hp, wp), 'bs': B, 'stage_8c': {'next_conf_c01': d['pre_conf01']}}
        m0 = m1 = None
        if pad:
            data['mask_4c0'], data['mask_4c1'] = d['mask0'], d['mask1']
            m0, m1 = d['mask0'].flatten(1), d['mask1'].flatten(1)
            data['scale0'] = torch.rand(B, 2, generator=g) + 0.5
            data['scale1'] = torch.rand(B, 2, generator=g) + 0.5
        cfg = {'thr': 0.0101, 'test_thr': thr, 'pre_thr': [0.2], 'border_rm': 2, 'double_check': True,
               'train_pad_num_gt_min': 4096, 'match_type': 'softmax', 'dsmax_temperature': 1.0}
        cas = {'propagation': 'window', 'dilated': 1, 'detector_mode': None, 'grid_size': 4,
               'post_config': {'method': 'maxpool_nms' if nms else None, 'window_size': 5, 'topk': None, 'rt': None, 'rd': None}}
        mod = ref.CascadeMatching(cfg, cas).eval()
        with torch.no_grad():
            mod(f0, f1, idx01, idx10, data, mask_c0=m0, mask_c1=m1, level='4c', pre_level='8c')
        st = data['stage_4c']
        if tag == 'nms':
            cases.update(feat0=f0, feat1=f1, idx01=idx01, idx10=idx10, pre_conf=d['pre_conf01'],
                         conf01=st['conf_matrix'], next_conf01=st['next_conf_c01'], next_conf10=st['next_conf_c10'],
                         next_idx01=st['next_idx_c01'], next_idx10=st['next_idx_c10'])
        if tag == 'pad':
            cases.update(pad_mask0=d['mask0'], pad_mask1=d['mask1'], scale0=data['scale0'], scale1=data['scale1'],
                         pad_next_conf01=st['next_conf_c01'], pad_next_idx01=st['next_idx_c01'], pad_next_idx10=st['next_idx_c10'])
        for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0_c', 'mkpts1_c'):
            cases[f'{tag}_{k}'] = st[k]
    save('cascade_match', hw=torch.tensor([h, w]), **cases)


def gen_fine(ref):
    M = 24
    f0, f1 = synth.fine_inputs(M, 25, 64, seed=104)
    g = torch.Generator().manual_seed(9)
    mk0, mk1 = torch.rand(M, 2, generator=g) * 100, torch.rand(M, 2, generator=g) * 100
    b_ids = torch.randint(0, 2, (M,), generator=g)
    sc1 = torch.rand(2, 2, generator=g) + 0.5
    out = {}
    for tag, extra in (('plain', {}), ('scaled', {'scale0': sc1, 'scale1': sc1})):
        data = {'hw0_i': (128, 128), 'hw0_f': (64, 64), **extra,
                'stage_4c': {'mkpts0_c': mk0, 'mkpts1_c': mk1, 'mconf': torch.rand(M), 'b_ids': b_ids}}
        mod = ref.CascadeFineMatching('4c').eval()
        with torch.no_grad():
            mod(f0, f1, data)
        out[f'{tag}_expec_f'], out[f'{tag}_mkpts1_f'] = data['expec_f'], data['mkpts1_f']
    save('fine_match', feat_f0=f0, feat_f1=f1, mkpts1_c=mk1, b_ids=b_ids, scale1=sc1, **out)


def gen_windows(ref):
    """get_window_warp_idx of the reference's CascadeFeatureTransformer, called unbound on a stand-in object
    (constructing the whole transformer needs modules outside the hot path)."""
    tr = __import__('src.model.modules.transformer', fromlist=['CascadeFeatureTransformer'])
    prop = __import__('src.model.modules.propagations', fromlist=['get_propagations'])
    window, full = prop.get_propagations({'propagation': 'window', 'window_size': 5, 'dilated': 1})

    class Stand:
        pass
    s = Stand()
    s.window, s.full_window = window, full
    g = torch.Generator().manual_seed(10)
    H, W = 9, 13
    idx = torch.randint(0, H * W, (2, H * W), generator=g)
    idx[0, :4] = torch.tensor([0, W - 1, (H - 1) * W, H * W - 1])      # corners: exercise the rigid border shift
    pos, _ = tr.CascadeFeatureTransformer.get_window_warp_idx(s, idx, 2, H, W)
    save('windows', idx=idx, hw=torch.tensor([H, W]), pos=pos)


def gen_widen(ref):
    """SURVEY section 8f "next" rows: the reference's CoarseMatching, CascadeFinePreprocess, QuadtreeAttention and
    CascadeQuadtreeAttention modules, unmodified, on seeded inputs (weights stored next to the outputs)."""
    import importlib
    g = torch.Generator().manual_seed(11)
    # -- CoarseMatching (dense dual softmax at 1/8)
    cmod = importlib.import_module('src.model.functions.coarse_matching')
    B, h, w, C = 2, 12, 16, 64
    f0 = torch.randn(B, h * w, C, generator=g)
    f1 = 0.8 * f0[:, torch.randperm(h * w, generator=g)] + 0.6 * torch.randn(B, h * w, C, generator=g)
    cfg = {'thr': 0.2, 'border_rm': 2, 'train_coarse_percent': 0.3, 'train_pad_num_gt_min': 200, 'match_type': 'dual_softmax',
           'dsmax_temperature': 0.1}
    data = {'hw0_i': (h * 8, w * 8), 'hw1_i': (h * 8, w * 8), 'hw0_8c': (h, w), 'hw1_8c': (h, w), 'bs': B}
    m = cmod.CoarseMatching(cfg).eval()
    with torch.no_grad():
        m(f0, f1, data)
    st = data['stage_8c']
    save('widen_coarse_match', feat0=f0, feat1=f1, temperature=torch.tensor(0.1), next_conf01=st['next_conf_c01'], next_idx01=st['next_idx_c01'],
         next_conf10=st['next_conf_c10'], next_idx10=st['next_idx_c10'], hw=torch.tensor([h, w]), thr=torch.tensor(cfg['thr']),
         border_rm=torch.tensor(cfg['border_rm']),
         **{'m_' + k: st[k] for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0_c', 'mkpts1_c')})     # get_coarse_match (:91-153)
    # -- CascadeFinePreprocess
    fmod = importlib.import_module('src.model.functions.fine_matching')
    Bf, Cf, Cc, hc, wc, stride, M = 2, 32, 64, 10, 14, 2, 40
    ff0, ff1 = torch.randn(Bf, Cf, hc * stride, wc * stride, generator=g), torch.randn(Bf, Cf, hc * stride, wc * stride, generator=g)
    fc0, fc1 = torch.randn(Bf, hc * wc, Cc, generator=g), torch.randn(Bf, hc * wc, Cc, generator=g)
    b_ids = torch.randint(0, Bf, (M,), generator=g).sort()[0]
    i_ids, j_ids = torch.randint(0, hc * wc, (M,), generator=g), torch.randint(0, hc * wc, (M,), generator=g)
    i_ids[:4] = torch.tensor([0, wc - 1, (hc - 1) * wc, hc * wc - 1])
    out = {}
    for cat in (False, True):
        torch.manual_seed(12)
        pm = fmod.CascadeFinePreprocess({'fine_concat_coarse_feat': cat, 'fine_window_size': 5}, {'d_model': Cf}, {'d_model': Cc}, '4c').eval()
        d = {'hw0_f': (hc * stride, wc * stride), 'hw0_4c': (hc, wc), 'hw1_4c': (hc, wc), 'stage_4c': {'b_ids': b_ids, 'i_ids': i_ids, 'j_ids': j_ids}}
        with torch.no_grad():
            o0, o1 = pm(ff0, ff1, fc0, fc1, d)
        tag = 'cat' if cat else 'plain'
        out[f'{tag}_out0'], out[f'{tag}_out1'] = o0, o1
        if cat:
            out.update({'w_' + k.replace('.', '_'): v for k, v in pm.state_dict().items()})
    save('widen_fine_preprocess', feat_f0=ff0, feat_f1=ff1, feat_c0=fc0, feat_c1=fc1, b_ids=b_ids, i_ids=i_ids, j_ids=j_ids,
         hw_c=torch.tensor([hc, wc]), stride=torch.tensor(stride), **out)
    # -- QuadtreeAttention (B and A) and CascadeQuadtreeAttention
    amod = importlib.import_module('src.model.modules.quadtree_attention')
    Ba, Ca, nh, H, W, topks = 1, 64, 2, 16, 24, [8, 4, 4]
    x, t = torch.randn(Ba, H * W, Ca, generator=g), torch.randn(Ba, H * W, Ca, generator=g)
    out = {}
    for typ in ('B',):        # attn_type='A' cannot run in the reference (QTAttA.forward takes no rel_pos, quadtree_attention.py:94)
        torch.manual_seed(13)
        layer = amod.QuadtreeAttention(Ca, nh, topks, scale=3, attn_type=typ).eval()
        for p in (layer.q_proj, layer.k_proj, layer.v_proj):
            torch.nn.init.normal_(p.weight, std=0.12, generator=g)       # spread logits; the default std 0.02 gives flat softmaxes
        with torch.no_grad():
            out[f'out_{typ}'] = layer(x, t, H, W)
        out.update({f'{typ}_' + k.replace('.', '_'): v for k, v in layer.state_dict().items()})
    d = synth.cascade_inputs(1, Ca, 16, 16, seed=105)
    xc, tc = d['feat0'].flatten(2).transpose(1, 2).contiguous(), d['feat1'].flatten(2).transpose(1, 2).contiguous()
    torch.manual_seed(14)
    cl = amod.CascadeQuadtreeAttention(Ca, nh, dilated=1).eval()
    for p in (cl.q_proj, cl.k_proj, cl.v_proj):
        torch.nn.init.normal_(p.weight, std=0.12, generator=g)
    with torch.no_grad():
        co, cup = cl(xc, tc, 16, 16, idx=d['topk_pos01'])
    out.update({'C_' + k.replace('.', '_'): v for k, v in cl.state_dict().items()})
    save('widen_attention_layers', x=x, target=t, hw=torch.tensor([H, W]), topks=torch.tensor(topks), nhead=torch.tensor(nh),
         cas_x=xc, cas_target=tc, cas_topk_pos=d['topk_pos01'], cas_out=co, cas_upsampled_idx=cup, **out)


def gen_coarse_match_masked(ref):
    """The reference's CoarseMatching with padding masks (mask_c0 / mask_c1, coarse_matching.py:64-65): bottom / right bands of
    the 1/8 grids padded, differently per image and batch element."""
    import importlib
    g = torch.Generator().manual_seed(31)
    cmod = importlib.import_module('src.model.functions.coarse_matching')
    B, h, w, C = 2, 12, 16, 64
    f0 = torch.randn(B, h * w, C, generator=g)
    f1 = 0.8 * f0[:, torch.randperm(h * w, generator=g)] + 0.6 * torch.randn(B, h * w, C, generator=g)

    def band_mask(valid):
        m = torch.zeros(B, h, w, dtype=torch.bool)
        for b, (vh, vw) in enumerate(valid):
            m[b, :vh, :vw] = True
        return m.reshape(B, h * w)
    m0, m1 = band_mask([(10, 16), (12, 11)]), band_mask([(12, 13), (9, 16)])
    cfg = {'thr': 0.2, 'border_rm': 2, 'train_coarse_percent': 0.3, 'train_pad_num_gt_min': 200, 'match_type': 'dual_softmax',
           'dsmax_temperature': 0.1}
    # mask_8c0 / mask_8c1 in `data` select the padded border removal of get_coarse_match (mask_border_with_padding, :124-127)
    data = {'hw0_i': (h * 8, w * 8), 'hw1_i': (h * 8, w * 8), 'hw0_8c': (h, w), 'hw1_8c': (h, w), 'bs': B,
            'mask_8c0': m0.reshape(B, h, w), 'mask_8c1': m1.reshape(B, h, w)}
    m = cmod.CoarseMatching(cfg).eval()
    with torch.no_grad():
        m(f0, f1, data, mask_c0=m0, mask_c1=m1)
    st = data['stage_8c']
    save('widen_coarse_match_masked', feat0=f0, feat1=f1, mask0=m0, mask1=m1, temperature=torch.tensor(0.1),
         next_conf01=st['next_conf_c01'], next_idx01=st['next_idx_c01'], next_conf10=st['next_conf_c10'], next_idx10=st['next_idx_c10'],
         hw=torch.tensor([h, w]), thr=torch.tensor(cfg['thr']), border_rm=torch.tensor(cfg['border_rm']),
         **{'m_' + k: st[k] for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0_c', 'mkpts1_c')})


def gen_relative_pe(ref):
    """get_window_warp_idx + get_relative_pe of the reference's CascadeFeatureTransformer, unbound on a stand-in that carries
    what they read (window, LB, the two nn.Embedding tables), followed by the reference's CascadeQTAttB with that bias.
    Case s2: the 4c stage (window and target from the same 1/8 matches, LB = 10); case s4: the 2c stage (window from 1/4
    matches, target from the 1/8 matches, LB = 30).  Images of different sizes so the two grids cannot be confused."""
    tr = __import__('src.model.modules.transformer', fromlist=['CascadeFeatureTransformer'])
    prop = __import__('src.model.modules.propagations', fromlist=['get_propagations'])
    window, full = prop.get_propagations({'propagation': 'window', 'window_size': 5, 'dilated': 1})
    g = torch.Generator().manual_seed(21)
    B, nh = 2, 2
    hw8 = [(6, 8), (7, 9)]
    out = {}
    for tag, s, sr in (('s2', 2, 2), ('s4', 4, 4)):
        class Stand:
            pass
        st = Stand()
        st.window, st.full_window = window, full
        st.LB = 5 * 2 if sr == 2 else 5 * 6                              # transformer.py:357-360
        st.w_pos_bias = torch.nn.Embedding(st.LB * 2 + sr, nh)
        st.h_pos_bias = torch.nn.Embedding(st.LB * 2 + sr, nh)
        with torch.no_grad():
            st.w_pos_bias.weight.copy_(torch.randn(st.LB * 2 + sr, nh, generator=g))
            st.h_pos_bias.weight.copy_(torch.randn(st.LB * 2 + sr, nh, generator=g))
        out[f'{tag}_w_table'], out[f'{tag}_h_table'] = st.w_pos_bias.weight.detach(), st.h_pos_bias.weight.detach()
        data = {'hw0_8c': hw8[0], 'hw1_8c': hw8[1], 'stage_8c': {}}
        for i in (0, 1):
            (h, w), (ho, wo) = hw8[i], hw8[1 - i]
            # a coherent match field (scaled identity + shift) with 15 % random entries: tile and fallback cells both occur
            yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
            ty = (yy * ho // h + 1).clamp(0, ho - 1)
            tx = (xx * wo // w - 1).clamp(0, wo - 1)
            t8 = (ty * wo + tx).reshape(1, -1).repeat(B, 1)
            rnd = torch.rand(B, h * w, generator=g) < 0.15
            t8 = torch.where(rnd, torch.randint(0, ho * wo, (B, h * w), generator=g), t8)
            data['stage_8c']['next_idx_c01' if i == 0 else 'next_idx_c10'] = t8
        for i in (0, 1):
            (h, w), (ho, wo) = hw8[i], hw8[1 - i]
            t8 = data['stage_8c']['next_idx_c01' if i == 0 else 'next_idx_c10']
            H, W, H1, W1 = h * s, w * s, ho * s, wo * s                  # current level of the query / the other image
            if s == 2:
                prev_idx = t8                                            # the 4c stage windows around the 1/8 matches
            else:                                                        # the 2c stage windows around 1/4 matches near the 1/8 ones
                t8u = t8.reshape(B, h, 1, w, 1).expand(B, h, 2, w, 2).reshape(B, -1)
                py = (torch.div(t8u, wo, rounding_mode='trunc') * 2 + torch.randint(-3, 5, t8u.shape, generator=g)).clamp(0, ho * 2 - 1)
                px = (t8u % wo * 2 + torch.randint(-3, 5, t8u.shape, generator=g)).clamp(0, wo * 2 - 1)
                prev_idx = py * (wo * 2) + px
            pos, _ = tr.CascadeFeatureTransformer.get_window_warp_idx(st, prev_idx, B, H1 // 2, W1 // 2)
            with torch.no_grad():
                rp = tr.CascadeFeatureTransformer.get_relative_pe(st, data, H, pos, torch.device('cpu'), i=i)
            if s == 4 and i == 1:                                         # keep the fixture small: one direction at the 2c stage
                continue
            q = torch.randn(B, nh * 32, H, W, generator=g)
            k = torch.randn(B, nh * 32, H1, W1, generator=g)
            v = torch.randn(B, nh * 32, H1, W1, generator=g)
            with torch.no_grad():
                msg, up = ref.CascadeQTAttB(nh, 32, dilated=1)(q, k, v, pos, rp.to(torch.float32))
            out.update({f'{tag}_{i}_tgt_idx': t8, f'{tag}_{i}_prev_idx': prev_idx, f'{tag}_{i}_pos': pos, f'{tag}_{i}_rel_pos': rp,
                        f'{tag}_{i}_q': q, f'{tag}_{i}_k': k, f'{tag}_{i}_v': v, f'{tag}_{i}_msg': msg})
    save('widen_relative_pe', hw8=torch.tensor(hw8), nhead=torch.tensor(nh), **out)


def gen_guided(ref):
    """The reference's QTAttGuided on a single-level pyramid (the only case its merge can run, quadtree_attention.py:385)."""
    g = torch.Generator().manual_seed(41)
    B, nh, C, h, w, K = 2, 4, 128, 16, 24, 8
    q, k, v = (torch.randn(B, C, h, w, generator=g) for _ in range(3))
    pos = torch.stack([torch.randint(0, h // 2, (B, (h // 2) * (w // 2), K, nh), generator=g),
                       torch.randint(0, w // 2, (B, (h // 2) * (w // 2), K, nh), generator=g)])
    m = ref.QTAttGuided(nh, C // nh, scale=1, topks=[K]).eval()
    with torch.no_grad():
        m.weight.copy_(torch.tensor([0.7]))
        out = m([q], [k], [v], topk_pos=pos)
    m3 = ref.QTAttGuided(nh, C // nh, scale=3, topks=[K]).eval()                 # the weight vector may be longer than the pyramid
    with torch.no_grad():
        m3.weight.copy_(torch.tensor([0.7, -0.4, 1.3]))
        out3 = m3([q], [k], [v], topk_pos=pos)
    save('qtatt_guided', q=q, k=k, v=v, topk_pos=pos, nhead=torch.tensor(nh), out=out, weight3=m3.weight.detach(), out3=out3)


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = refload.load()
    gen_qtatt(ref)
    gen_cascade_qtatt(ref)
    gen_cascade_match(ref)
    gen_fine(ref)
    gen_windows(ref)
    gen_widen(ref)
    gen_coarse_match_masked(ref)
    gen_relative_pe(ref)
    gen_guided(ref)
