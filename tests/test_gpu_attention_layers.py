"""SURVEY section 8f "next" #2: the token-major entries (pyramid built inside, no NCHW maps) and the QuadtreeAttention /
CascadeQuadtreeAttention drop-ins (reference src/model/modules/quadtree_attention.py:9-171) against the reference
formulation: permute to NCHW -> 1x1 conv -> avg_pool2d pyramid -> QTAttB / CascadeQTAttB -> proj."""
import pytest
import torch
import torch.nn.functional as tF

import casmtr_b200
from casmtr_b200 import functional as F
from casmtr_b200 import synth
from oracle import qtatt as oqt

pytestmark = pytest.mark.gpu


def _nchw(t, H, W):
    B, N, C = t.shape
    return t.permute(0, 2, 1).reshape(B, C, H, W).contiguous()


@pytest.mark.parametrize('B,nh,H,W,H1,W1,topks,typ', [(2, 4, 32, 48, 32, 48, [8, 8, 8], 'B'), (1, 8, 104, 104, 104, 104, [32, 16, 8], 'B'),
                                                      (1, 2, 16, 24, 32, 16, [6, 5], 'B'), (2, 4, 32, 32, 32, 32, [8, 8, 8], 'A')])
def test_tokens_entry_equals_pyramid_entry(dev, B, nh, H, W, H1, W1, topks, typ):
    """Same kernels, different front end: the results must agree to fp32 rounding of the pooling (the four taps are
    summed in avg_pool2d's order, so in practice bit for bit) -- top-k sets identical, messages within 1e-5."""
    C, n = nh * 32, len(topks)
    g = torch.Generator().manual_seed(21)
    q, k, v = (torch.randn(B, H * W, C, generator=g).to(dev), torch.randn(B, H1 * W1, C, generator=g).to(dev),
               torch.randn(B, H1 * W1, C, generator=g).to(dev))
    w = torch.randn(n, generator=g).to(dev)
    got, gi, gs = F.qtatt_tokens_forward(q, k, v, (H, W), (H1, W1), topks, nh, weight=w, attn_type=typ, return_topk=True)
    def pyramid(t, h, ww):
        x, out = _nchw(t, h, ww), []
        for l in range(n):
            out.append(x)
            if l != n - 1:
                x = tF.avg_pool2d(x, kernel_size=2, stride=2)
        return out
    want, wi, ws = F.qtatt_forward(pyramid(q, H, W), pyramid(k, H1, W1), pyramid(v, H1, W1), topks, nh, weight=w, attn_type=typ, return_topk=True)
    for a, b in zip(gi, wi):
        same = (a.sort(dim=2)[0] == b.sort(dim=2)[0]).all(dim=2).float().mean().item()
        assert same > 0.999, same                                   # a 1-ulp pooling difference may flip an exact near-tie
    bad = ((got - want).abs().amax(dim=(2, 3)) > 1e-5).float().mean().item()
    assert bad < 1e-3, bad


def test_pool_matches_avg_pool2d(dev):
    B, H, W, C = 2, 12, 20, 64
    x = torch.randn(B, H * W, C, device=dev)
    # a 2-level QTAttB whose coarse level sees the pooled maps: compare against the oracle fed with avg_pool2d maps
    got = F.qtatt_tokens_forward(x, x, x, (H, W), (H, W), [4, 4], 2, weight=torch.zeros(2, device=dev), attn_type='B')
    pyr = [_nchw(x, H, W).cpu(), tF.avg_pool2d(_nchw(x, H, W), 2, 2).cpu()]
    want = oqt.qtatt_b(pyr, pyr, pyr, torch.zeros(2), [4, 4], 2)
    assert (got.cpu() - want).abs().max() < 1e-3


@pytest.mark.parametrize('typ', ['B', 'A'])
def test_quadtree_attention_layer(dev, typ):
    """Drop-in layer vs the reference's formulation of the same layer around the (already verified) NCHW QTAtt modules."""
    B, C, nh, H, W = 2, 128, 4, 32, 32
    topks = [8, 8, 8]
    torch.manual_seed(5)
    layer = casmtr_b200.QuadtreeAttention(C, nh, topks, scale=3, attn_type=typ).to(dev).eval()
    for p in (layer.q_proj, layer.k_proj, layer.v_proj):
        torch.nn.init.normal_(p.weight, std=0.09)                   # logits with some spread (std 0.02 gives flat softmaxes)
    x, t = torch.randn(B, H * W, C, device=dev), torch.randn(B, H * W, C, device=dev)
    with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):     # cuDNN convolutions default to TF32: keep the check fp32
        got = layer(x, t, H, W)
        q, k, v = layer.q_proj(_nchw(x, H, W)), layer.k_proj(_nchw(t, H, W)), layer.v_proj(_nchw(t, H, W))
        qs, ks, vs = [], [], []
        for i in range(3):
            qs.append(q.float()), ks.append(k.float()), vs.append(v.float())
            if i != 2:
                q, k, v = (tF.avg_pool2d(a, kernel_size=2, stride=2) for a in (q, k, v))
        want = layer.proj(layer.py_att(qs, ks, vs).view(B, -1, C))
    assert got.shape == want.shape == (B, H * W, C)
    bad = ((got - want).abs().amax(dim=2) > 1e-4).float().mean().item()       # conv vs linear rounding can flip a near-tie
    assert bad < 5e-3, bad
    assert set(layer.state_dict()) == {'q_proj.weight', 'k_proj.weight', 'v_proj.weight', 'proj.weight', 'proj.bias'} | ({'py_att.weight'} if typ == 'B' else set())


def test_cascade_quadtree_attention_layer(dev):
    B, nh, h, w = 1, 4, 48, 64
    C = nh * 32
    d = synth.cascade_inputs(B, C, h, w, seed=31)
    torch.manual_seed(6)
    layer = casmtr_b200.CascadeQuadtreeAttention(C, nh, dilated=1).to(dev).eval()
    for p in (layer.q_proj, layer.k_proj, layer.v_proj):
        torch.nn.init.normal_(p.weight, std=0.09)
    x = d['feat0'].flatten(2).transpose(1, 2).contiguous().to(dev)
    t = d['feat1'].flatten(2).transpose(1, 2).contiguous().to(dev)
    tp = d['topk_pos01'].to(dev)
    with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):
        got, up = layer(x, t, h, w, idx=tp)
        q, k, v = layer.q_proj(_nchw(x, h, w)), layer.k_proj(_nchw(t, h, w)), layer.v_proj(_nchw(t, h, w))
        msg, up_want = layer.cross_attn(q.float(), k.float(), v.float(), tp, None)
        want = layer.proj(msg.view(B, -1, C))
    assert torch.equal(up, up_want)                                  # integer work: bit-exact
    assert (got - want).abs().max() < 1e-3


# ---- SURVEY section 8f "next" #3: window-index plumbing
def test_window_idx_vs_reference(dev):
    """casmtr_window_idx_fwd against the output of the reference's get_window_warp_idx (tests/golden/windows.npz)."""
    from golden_util import load
    g = load('windows')
    H, W = g['hw'].tolist()
    pos = F.window_warp_idx(g['idx'].to(dev), H, W, 5)
    assert torch.equal(pos.cpu(), g['pos'])


@pytest.mark.parametrize('B,nh,h,w,tokens', [(1, 4, 64, 64, False), (2, 2, 48, 80, True), (1, 4, 208, 208, False)])
def test_cascade_from_next_idx(dev, B, nh, h, w, tokens):
    """Fused window expansion: CascadeQTAttB fed with next_idx == fed with the expanded topk_pos, bit for bit."""
    C = nh * 32
    d = synth.cascade_inputs(B, C, h, w, seed=41)
    v = torch.randn(B, C, h, w, generator=torch.Generator().manual_seed(3))
    q, k, v = d['feat0'].to(dev), d['feat1'].to(dev), v.to(dev)
    tp, ni = d['topk_pos01'].to(dev), d['next_idx01'].to(dev)
    assert torch.equal(F.window_warp_idx(ni, h // 2, w // 2, 5), tp)
    att = casmtr_b200.CascadeQTAttB(nh, 32, dilated=1)
    want, up_want = att(q, k, v, tp, None)
    if tokens:
        tk = lambda x: x.flatten(2).transpose(1, 2).contiguous()
        got, up = F.cascade_qtatt_forward(tk(q), tk(k), tk(v), ni, None, nh, hw_q=(h, w), hw_k=(h, w))
    else:
        got, up = att(q, k, v, ni, None)
    assert torch.equal(up, up_want) and torch.equal(got, want)
