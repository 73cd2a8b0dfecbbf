"""SURVEY section 8f "next" #3, second half: the relative position bias of the cascade cross attention
(reference CascadeFeatureTransformer.get_relative_pe, src/model/modules/transformer.py:473-509; indoor config).
The stand-alone kernel must reproduce the reference tensor bit for bit (two fp32 table rows added); the attention kernels
fed with the tables instead of the tensor must give the same message as when fed with the tensor."""
import pytest
import torch

from casmtr_b200 import functional as F
from casmtr_b200.modules.attention_layers import CascadeRelativePE
from golden_util import load
from oracle import qtatt as oqt
from oracle import widen

pytestmark = pytest.mark.gpu

CASES = [('s2', 0, 2), ('s2', 1, 2), ('s4', 0, 4)]


def _pe(g, tag, i, s, dev):
    hw8 = g['hw8'].tolist()
    return F.RelativePE(g[f'{tag}_w_table'].to(dev), g[f'{tag}_h_table'].to(dev), 10 if s == 2 else 30,
                        g[f'{tag}_{i}_tgt_idx'].to(dev), hw8[i], hw8[1 - i][1]), hw8[i][0] * s, hw8[i][1] * s, hw8[1 - i][0] * s, hw8[1 - i][1] * s


@pytest.mark.parametrize('tag,i,s', CASES)
def test_relative_pe_tensor_golden(dev, tag, i, s):
    g = load('widen_relative_pe')
    pe, H, W, H1, W1 = _pe(g, tag, i, s, dev)
    p = f'{tag}_{i}_'
    pos = F.window_warp_idx(g[p + 'prev_idx'].to(dev), H1 // 2, W1 // 2, 5)
    assert torch.equal(pos.cpu(), g[p + 'pos'])
    rp = F.relative_pe(pe, pos, (H, W))
    assert torch.equal(rp.cpu(), g[p + 'rel_pos'])                         # bit-exact


@pytest.mark.parametrize('tag,i,s', CASES)
@pytest.mark.parametrize('entry', ['topk_pos', 'next_idx', 'tokens'])
def test_fused_relative_pe_golden(dev, tag, i, s, entry):
    """CascadeQTAttB with the bias computed in the kernels vs the reference's message (1e-3 abs; observed ~1e-6) and
    vs the same kernels fed with the reference's tensor (bit-exact: the same two fp32 values are added in the same order)."""
    g = load('widen_relative_pe')
    pe, H, W, H1, W1 = _pe(g, tag, i, s, dev)
    p = f'{tag}_{i}_'
    nh = int(g['nhead'])
    q, k, v = g[p + 'q'].to(dev), g[p + 'k'].to(dev), g[p + 'v'].to(dev)
    pos, rp = g[p + 'pos'].to(dev), g[p + 'rel_pos'].to(dev)
    want_t, up_t = F.cascade_qtatt_forward(q, k, v, pos, rp, nh)
    if entry == 'topk_pos':
        got, up = F.cascade_qtatt_forward(q, k, v, pos, pe, nh)
    elif entry == 'next_idx':
        got, up = F.cascade_qtatt_forward(q, k, v, g[p + 'prev_idx'].to(dev), pe, nh)
    else:
        tok = lambda t: t.flatten(2).transpose(1, 2).contiguous()
        got, up = F.cascade_qtatt_forward(tok(q), tok(k), tok(v), g[p + 'prev_idx'].to(dev), pe, nh, hw_q=(H, W), hw_k=(H1, W1))
    assert (got.cpu() - g[p + 'msg']).abs().max() < 1e-3
    assert torch.equal(got, want_t) and torch.equal(up, up_t)


def test_module_drop_in(dev):
    """CascadeRelativePE mirrors the reference method signatures: get_window_warp_idx(idx, B, H, W), get_relative_pe(data, H, window_idx, device, i)."""
    g = load('widen_relative_pe')
    hw8 = g['hw8'].tolist()
    mod = CascadeRelativePE(int(g['nhead']), window_size=5, sr_ratio=2).to(dev)
    assert mod.LB == 10 and sorted(mod.state_dict()) == ['h_pos_bias.weight', 'w_pos_bias.weight']
    mod.load_state_dict({'w_pos_bias.weight': g['s2_w_table'], 'h_pos_bias.weight': g['s2_h_table']})
    data = {'hw0_8c': tuple(hw8[0]), 'hw1_8c': tuple(hw8[1]),
            'stage_8c': {'next_idx_c01': g['s2_0_tgt_idx'].to(dev), 'next_idx_c10': g['s2_1_tgt_idx'].to(dev)}}
    for i in (0, 1):
        H, (H1p, W1p) = hw8[i][0] * 2, hw8[1 - i]
        pos, full = mod.get_window_warp_idx(g[f's2_{i}_prev_idx'].to(dev), 2, H1p, W1p)
        assert full is None and torch.equal(pos.cpu(), g[f's2_{i}_pos'])
        rp = mod.get_relative_pe(data, H, pos, dev, i=i)
        assert torch.equal(rp.cpu(), g[f's2_{i}_rel_pos'])
    assert CascadeRelativePE(4, 5, 4).LB == 30


@pytest.mark.parametrize('B,nh,hw8_0,hw8_1,s', [(2, 4, (60, 80), (60, 80), 2),      # BASELINE configs[3]: 640x480 indoor, 1/4 level
                                                 (1, 2, (20, 26), (22, 24), 4)])     # a 1/2-level stage on unequal images
def test_fused_relative_pe_full_size(dev, B, nh, hw8_0, hw8_1, s):
    """At sizes the CPU oracle would take minutes for: tensor path == fused path bit for bit, the tensor itself against the oracle
    restatement (pinned to the reference in tests/test_oracle_golden.py), and a sample of message rows against the oracle."""
    g = torch.Generator().manual_seed(5)
    (h, w), (ho, wo) = hw8_0, hw8_1
    LB, sr = (10, 2) if s == 2 else (30, 4)
    H, W, H1, W1 = h * s, w * s, ho * s, wo * s
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    t8 = ((yy * ho // h + 2).clamp(0, ho - 1) * wo + (xx * wo // w + 1).clamp(0, wo - 1)).reshape(1, -1).repeat(B, 1)
    t8 = torch.where(torch.rand(B, h * w, generator=g) < 0.1, torch.randint(0, ho * wo, (B, h * w), generator=g), t8)
    if s == 2:
        prev = t8
    else:
        t8u = t8.reshape(B, h, 1, w, 1).expand(B, h, 2, w, 2).reshape(B, -1)
        py = (torch.div(t8u, wo, rounding_mode='trunc') * 2 + torch.randint(-3, 5, t8u.shape, generator=g)).clamp(0, ho * 2 - 1)
        px = (t8u % wo * 2 + torch.randint(-3, 5, t8u.shape, generator=g)).clamp(0, wo * 2 - 1)
        prev = py * (wo * 2) + px
    wt, ht = torch.randn(2 * LB + sr, nh, generator=g), torch.randn(2 * LB + sr, nh, generator=g)
    q, k, v = torch.randn(B, nh * 32, H, W, generator=g), torch.randn(B, nh * 32, H1, W1, generator=g), torch.randn(B, nh * 32, H1, W1, generator=g)
    pe = F.RelativePE(wt.to(dev), ht.to(dev), LB, t8.to(dev), (h, w), wo)
    pos = F.window_warp_idx(prev.to(dev), H1 // 2, W1 // 2, 5)
    rp = F.relative_pe(pe, pos, (H, W))
    assert torch.equal(rp.cpu(), widen.relative_pe(pos.cpu(), t8, wt, ht, LB, (h, w), wo, H))
    qd, kd, vd = q.to(dev), k.to(dev), v.to(dev)
    want, up_w = F.cascade_qtatt_forward(qd, kd, vd, pos, rp, nh)
    got, up = F.cascade_qtatt_forward(qd, kd, vd, prev.to(dev), pe, nh)
    assert torch.equal(got, want) and torch.equal(up, up_w)
    # oracle on the first batch element only (seconds)
    ref, _ = oqt.cascade_qtatt_b(q[:1], k[:1], v[:1], pos[:1].cpu(), rp[:1].cpu(), nh, 1)
    assert (got[:1].cpu() - ref).abs().max() < 1e-3


@pytest.mark.parametrize('window', [1, 3])
def test_fused_relative_pe_other_windows(dev, window):
    """Windows other than 5x5 take the generic gather kernel (runtime candidate count): tensor path == fused path, both == oracle."""
    g = torch.Generator().manual_seed(40 + window)
    B, nh, (h, w), (ho, wo), s, LB = 2, 2, (8, 10), (9, 7), 2, 10
    H, W, H1, W1 = h * s, w * s, ho * s, wo * s
    t8 = torch.randint(0, ho * wo, (B, h * w), generator=g)
    wt, ht = torch.randn(2 * LB + 2, nh, generator=g), torch.randn(2 * LB + 2, nh, generator=g)
    q, k, v = torch.randn(B, nh * 32, H, W, generator=g), torch.randn(B, nh * 32, H1, W1, generator=g), torch.randn(B, nh * 32, H1, W1, generator=g)
    pe = F.RelativePE(wt.to(dev), ht.to(dev), LB, t8.to(dev), (h, w), wo)
    pos = F.window_warp_idx(t8.to(dev), H1 // 2, W1 // 2, window)
    rp = F.relative_pe(pe, pos, (H, W))
    assert torch.equal(rp.cpu(), widen.relative_pe(pos.cpu(), t8, wt, ht, LB, (h, w), wo, H))
    want, up_w = F.cascade_qtatt_forward(q.to(dev), k.to(dev), v.to(dev), pos, rp, nh)
    for entry in (pos, t8.to(dev)):
        got, up = F.cascade_qtatt_forward(q.to(dev), k.to(dev), v.to(dev), entry, pe, nh, window=window)
        assert torch.equal(got, want) and torch.equal(up, up_w)
    ref, _ = oqt.cascade_qtatt_b(q, k, v, pos.cpu(), rp.cpu(), nh, 1)
    assert (got.cpu() - ref).abs().max() < 1e-3


def test_relative_pe_rejects_mismatched_grids(dev):
    g = torch.Generator().manual_seed(3)
    pe = F.RelativePE(torch.randn(22, 2).to(dev), torch.randn(22, 2).to(dev), 10, torch.zeros(1, 48, dtype=torch.int64, device=dev), (6, 8), 9)
    pos = torch.zeros(1, 48, 25, 2, dtype=torch.int64, device=dev)
    with pytest.raises(RuntimeError):                       # 12 x 18 is not a multiple of the 6 x 8 grid
        F.relative_pe(pe, torch.zeros(1, 54, 25, 2, dtype=torch.int64, device=dev), (12, 18))
    with pytest.raises(RuntimeError):                       # three heads, two-head tables
        F.cascade_qtatt_forward(torch.zeros(1, 96, 12, 16, device=dev), torch.zeros(1, 96, 14, 18, device=dev), torch.zeros(1, 96, 14, 18, device=dev),
                                pos, pe, 3)
