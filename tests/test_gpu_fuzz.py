"""GPU parity on seeded RANDOM shapes (a fixed list of draws, so the suite stays deterministic): every stage of the path
against the CPU oracle on grids, head counts, candidate counts, thresholds and paddings the hand-picked cases do not
cover -- ragged tile edges of the TMA kernels, odd numbers of tiles, one-row grids of parents, large batches of tiny maps."""
import random

import pytest
import torch

import casmtr_b200
from casmtr_b200 import functional as F
from casmtr_b200 import synth
from oracle import cascade, fine, qtatt
from oracle.compare import check_qtatt_levels

pytestmark = pytest.mark.gpu


def _draws(n, seed, fn):
    rng = random.Random(seed)
    return [fn(rng) for _ in range(n)]


def _qt_case(r):
    lv = r.choice([1, 2, 3, 3])
    unit = 1 << (lv - 1)
    h, w = unit * r.randint(2, 9), unit * r.randint(2, 9)
    coarse = (h // unit) * (w // unit)
    topks = [min(r.choice([2, 4, 8, 16, 32]), coarse)]
    for _ in range(lv - 1):
        topks.append(min(r.choice([1, 2, 4, 8, 16]), 4 * topks[-1], 32))
    return r.choice([1, 1, 2, 3]), r.choice([1, 2, 4, 8]), h, w, topks, r.choice(['A', 'B', 'B'])


@pytest.mark.parametrize('B,nh,h,w,topks,typ', _draws(14, 2024, _qt_case))
def test_fuzz_qtatt(dev, B, nh, h, w, topks, typ):
    lv = len(topks)
    qs, ks, vs, wt = synth.qtatt_inputs(B, nh * 32, h, w, lv, seed=h * 131 + w)
    c = lambda lst: [t.to(dev) for t in lst]
    if typ == 'A':
        ref, aux = qtatt.qtatt_a(qs, ks, vs, topks, nh, return_aux=True)
    else:
        ref, aux = qtatt.qtatt_b(qs, ks, vs, wt, topks, nh, return_aux=True)
    out, ti, ts = F.qtatt_forward(c(qs), c(ks), c(vs), topks, nh, weight=wt.to(dev), attn_type=typ, return_topk=True)
    if lv > 1:
        check_qtatt_levels(out, ti, ts, ref, aux, h, w, lv, f'fuzz QTAtt{typ} {B}x{nh}x{h}x{w} {topks}')
    else:
        assert (out.cpu() - ref).abs().max() < 1e-3
    # the token-major entry must agree with the pyramid entry whenever the pyramid IS the avg-pool pyramid
    if typ == 'B' and lv > 1:
        q0, k0, v0 = (t[0].flatten(2).transpose(1, 2).contiguous().to(dev) for t in (qs, ks, vs))
        pyr = lambda t: [t] + [torch.nn.functional.avg_pool2d(t, 2 ** l, 2 ** l) for l in range(1, lv)]
        a = F.qtatt_forward(pyr(qs[0].to(dev)), pyr(ks[0].to(dev)), pyr(vs[0].to(dev)), topks, nh, weight=wt.to(dev))
        b = F.qtatt_tokens_forward(q0, k0, v0, (h, w), (h, w), topks, nh, weight=wt.to(dev))
        assert ((a - b).abs().amax(dim=(2, 3)) > 1e-5).float().mean().item() < 5e-3


def _cas_case(r):
    return r.choice([1, 2, 3]), r.choice([1, 2, 4]), 2 * r.randint(5, 40), 2 * r.randint(5, 40), r.random() < 0.3, r.choice([0.0, 0.1, 0.5])


@pytest.mark.parametrize('B,nh,h,w,rel,corrupt', _draws(10, 77, _cas_case))
def test_fuzz_cascade_stage(dev, B, nh, h, w, rel, corrupt):
    """CascadeQTAttB (tile kernel + fallback list, from topk_pos and from next_idx) and CascadeMatching on its upsampled_idx."""
    C = nh * 32
    d = synth.cascade_inputs(B, C, h, w, seed=h * 977 + w, corrupt=corrupt)
    g = torch.Generator().manual_seed(h + w)
    v = torch.randn(B, C, h, w, generator=g)
    rp = torch.randn(B, nh, h * w, 100, generator=g) if rel else None
    ref_m, ref_i = qtatt.cascade_qtatt_b(d['feat0'], d['feat1'], v, d['topk_pos01'], rp, nh, 1)
    att = casmtr_b200.CascadeQTAttB(nh, 32, dilated=1)
    q, k, vv = d['feat0'].to(dev), d['feat1'].to(dev), v.to(dev)
    rpd = None if rp is None else rp.to(dev)
    m1, i1 = att(q, k, vv, d['topk_pos01'].to(dev), rpd)
    m2, i2 = att(q, k, vv, d['next_idx01'].to(dev), rpd)
    assert torch.equal(i1.cpu(), ref_i) and torch.equal(i2, i1)
    assert (m1.cpu() - ref_m).abs().max() < 1e-3 and torch.equal(m1, m2)
    _, i10 = att(k, q, vv, d['topk_pos10'].to(dev), None)
    f0, f1 = d['feat0'].flatten(2).transpose(1, 2).contiguous(), d['feat1'].flatten(2).transpose(1, 2).contiguous()
    ref = cascade.cascade_match(f0, f1, ref_i, i10.cpu(), None, None, 1.0)
    out = F.cascade_match_forward(f0.to(dev), f1.to(dev), i1, i10, w0=w, w1=w)
    for t in ('01', '10'):
        gap = ref['conf' + t].topk(2, dim=2)[0]
        clear = (gap[..., 0] - gap[..., 1]) > 1e-6
        assert torch.equal(out['next_idx' + t].cpu()[clear], ref['next_idx' + t][clear])
        assert (out['next_conf' + t].cpu() - ref['next_conf' + t]).abs().max() < 1e-5
        assert (out['conf' + t].cpu() - ref['conf' + t]).abs().max() < 1e-5


def _ext_case(r):
    return (r.choice([1, 2, 4]), 2 * r.randint(4, 24), 2 * r.randint(4, 24), r.choice([None, 3, 5]), r.random() < 0.4, r.choice([0, 1, 2, 3]),
            r.random() < 0.5, r.choice([0.05, 0.2, 0.4, 2.0]))


@pytest.mark.parametrize('B,h,w,nms,pad,border,scales,thr', _draws(12, 5, _ext_case))
def test_fuzz_match_extract(dev, B, h, w, nms, pad, border, scales, thr):
    d = synth.cascade_inputs(B, 64, h, w, seed=h * 31 + w, pad=pad)
    f0, f1 = d['feat0'].flatten(2).transpose(1, 2).contiguous(), d['feat1'].flatten(2).transpose(1, 2).contiguous()
    i01 = qtatt.quad_to_raster(qtatt.cascade_window_idx(d['topk_pos01'], h, w).reshape(B, 1, -1, 1, 100).expand(B, 1, -1, 4, 100), h // 2, w // 2).reshape(B, h * w, 100).contiguous()
    i10 = qtatt.quad_to_raster(qtatt.cascade_window_idx(d['topk_pos10'], h, w).reshape(B, 1, -1, 1, 100).expand(B, 1, -1, 4, 100), h // 2, w // 2).reshape(B, h * w, 100).contiguous()
    o = cascade.cascade_match(f0, f1, i01, i10, None, None, 1.0)
    g = torch.Generator().manual_seed(h)
    s0 = torch.rand(B, 2, generator=g) + 0.5 if scales else None
    s1 = torch.rand(B, 2, generator=g) + 0.5 if scales else None
    kw = dict(test_thr=thr, border_rm=border, nms_window=nms, pre_thrs=[0.2], double_check=True)
    ref = cascade.extract_matches(o['next_conf01'], o['next_idx01'], o['next_idx10'], (h, w), (h, w), (h * 4, w * 4),
                                  pre_confs=[(d['pre_conf01'], h // 2, w // 2)], pad_mask0=d.get('mask0'), pad_mask1=d.get('mask1'),
                                  scale0=s0, scale1=s1, **kw)
    c = lambda t: None if t is None else t.to(dev)
    out = F.match_extract(c(o['next_conf01']), c(o['next_idx01']), c(o['next_idx10']), (h, w), (h, w), (h * 4, w * 4),
                          pre_confs=[(c(d['pre_conf01']), h // 2, w // 2)], pad_mask0=c(d.get('mask0')), pad_mask1=c(d.get('mask1')),
                          scale0=c(s0), scale1=c(s1), **kw)
    for key in ('b_ids', 'i_ids', 'j_ids', 'mask'):
        assert torch.equal(out[key].cpu(), ref[key]), key
    for key in ('mconf', 'mkpts0_c', 'mkpts1_c'):
        assert torch.equal(out[key].cpu(), ref[key].to(torch.float32)), key
    M = out['b_ids'].shape[0]
    if M:
        f_0, f_1 = synth.fine_inputs(M, 25, 64, seed=M)
        e, kp = fine.fine_match(f_0, f_1, ref['mkpts1_c'].float(), 2.0)
        eo, ko = F.fine_match_forward(f_0.to(dev), f_1.to(dev), out['mkpts1_c'], 2.0)
        assert (eo.cpu() - e).abs().max() < 1e-5 and (ko.cpu() - kp).abs().max() < 1e-3
