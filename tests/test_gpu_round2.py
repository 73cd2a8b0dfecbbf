"""GPU tests added in round 2 for the gaps VERDICT r1 / ADVICE r1 list:
  * the import-name shims, through the path the reference's own Python takes (`import score_computation_cuda` ...);
  * the legacy FineMatching head;
  * CascadeQTAttB windows that cross the key grid's border (the TMA tile path must hand them to the gather kernel);
  * QTAttB.weight longer than the pyramid (soft-max over the whole parameter, reference :264);
  * NaN features in CascadeMatching (no out-of-bounds gather);
  * two devices in one process (per-device kernel attributes) when the box has them."""
import importlib
import sys

import pytest
import torch

from casmtr_b200 import functional as F
from oracle import cascade as ocas, fine as ofine, ops, qtatt as oqt

pytestmark = pytest.mark.gpu


def test_shims_resolve_the_reference_import_names_forward_and_backward(dev):
    """cuda_imp/QuadTreeAttention/QuadtreeAttention/functions/quadtree_attention.py:1-2 does `import score_computation_cuda`,
    `import value_aggregation_cuda`; src/model/functions/cascade_functions.py:1 does `import fast_score_computation`.  After
    shims.install() those names resolve to modules backed by libcasmtr_b200.so; the reference's autograd Functions are restated
    here around them exactly as the reference calls them (forward returns a list, value aggregation writes into a caller-
    allocated, zero-filled output, backward accumulates into zero-filled gradients)."""
    from casmtr_b200 import shims
    shims.install()
    for name in ('score_computation_cuda', 'value_aggregation_cuda', 'fast_score_computation'):
        sys.modules.pop(name, None)
    sc, va, fs = (importlib.import_module(n) for n in ('score_computation_cuda', 'value_aggregation_cuda', 'fast_score_computation'))
    assert 'casmtr_b200/shims' in sc.__file__.replace('\\', '/')

    class RefScore(torch.autograd.Function):            # functions/quadtree_attention.py:7-19
        @staticmethod
        def forward(ctx, query, key, index):
            x = sc.score_forward(query, key, index)
            ctx.save_for_backward(query, key, index)
            return x[0]

        @staticmethod
        def backward(ctx, grad_output):
            query, key, index = ctx.saved_tensors
            x = sc.score_backward(grad_output.contiguous(), query, key, index)
            return x[0], x[1], None

    class RefValue(torch.autograd.Function):            # :25-51
        @staticmethod
        def forward(ctx, score, value, index):
            ctx.save_for_backward(score, value, index)
            f = score.shape[2]
            s = score.flatten(1, 2).contiguous()
            i = index.flatten(1, 2).contiguous()
            b, N, _, H = s.shape
            out = s.new_zeros([b, N, H, value.shape[-1]]).contiguous()
            va.value_aggregation_forward(s, value, i, out)
            return out.reshape(b, N // f, f, H, -1)

        @staticmethod
        def backward(ctx, grad_output):
            score, value, index = ctx.saved_tensors
            f = score.shape[2]
            s, i = score.flatten(1, 2).contiguous(), index.flatten(1, 2).contiguous()
            go = grad_output.flatten(1, 2).contiguous()
            gs, gv = s.new_zeros(s.shape).contiguous(), value.new_zeros(value.shape).contiguous()
            va.value_aggregation_backward(go, s, value, i, gs, gv)
            return gs.reshape(score.shape), gv, None

    g = torch.Generator().manual_seed(5)
    B, N1, N2, H, D, K = 2, 24, 96, 4, 32, 12
    q = torch.randn(B, N1, 4, H, D, generator=g, dtype=torch.float64)
    k = torch.randn(B, N2, H, D, generator=g, dtype=torch.float64)
    v = torch.randn(B, N2, H, D, generator=g, dtype=torch.float64)
    idx = torch.randint(0, N2, (B, N1, K, H), generator=g)
    w = torch.randn(B, N1, 4, H, D, generator=g, dtype=torch.float64)
    # oracle: fp64 autograd through the op restatements
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    s_ref = ops.score5d(qr, kr, idx)                                     # [B,N1,4,K,H]
    a_ref = torch.softmax(s_ref, dim=3)
    idx5 = idx.unsqueeze(2).expand(B, N1, 4, K, H)
    o_ref = ops.value_agg(a_ref.flatten(1, 2), vr, idx5.flatten(1, 2)).reshape(B, N1, 4, H, D)
    (o_ref * w).sum().backward()
    qd, kd, vd = (t.float().to(dev).requires_grad_(True) for t in (q, k, v))
    s = RefScore.apply(qd, kd, idx.to(dev))
    a = torch.softmax(s, dim=3)
    o = RefValue.apply(a, vd, idx5.contiguous().to(dev))
    (o * w.float().to(dev)).sum().backward()
    assert (o.detach().cpu() - o_ref.detach().float()).abs().max() < 1e-4
    for got, ref in ((qd.grad, qr.grad), (kd.grad, kr.grad), (vd.grad, vr.grad)):
        assert (got.cpu() - ref.float()).abs().max() < 1e-4 * max(1.0, ref.abs().max().item())
    # the single-head correlation of the cascade stages (cascade_functions.py:8-22)
    q3, k3 = torch.randn(B, 40, 64, generator=g), torch.randn(B, 50, 64, generator=g)
    i3 = torch.randint(0, 50, (B, 40, 9), generator=g)
    out = fs.score_forward(q3.to(dev), k3.to(dev), i3.to(dev))
    assert isinstance(out, list) and (out[0].cpu() - ops.score3d(q3, k3, i3)).abs().max() < 1e-4
    gq, gk = fs.score_backward(torch.ones(B, 40, 9, device=dev), q3.to(dev), k3.to(dev), i3.to(dev))
    q3r, k3r = q3.double().requires_grad_(True), k3.double().requires_grad_(True)
    ops.score3d(q3r, k3r, i3).sum().backward()
    assert (gq.cpu() - q3r.grad.float()).abs().max() < 1e-4 and (gk.cpu() - k3r.grad.float()).abs().max() < 1e-3


def test_legacy_fine_matching_head(dev):
    """FineMatching (reference src/model/functions/fine_matching.py:195-261): same arithmetic as the cascade head, coarse
    matches read from the top level of `data`; incl. the per-sample scale1 branch (:257-258) and the M == 0 branch (:222-229)."""
    from casmtr_b200 import FineMatching
    g = torch.Generator().manual_seed(11)
    M, WW, C = 37, 25, 64
    f0, f1 = torch.randn(M, WW, C, generator=g), torch.randn(M, WW, C, generator=g)
    mk0, mk1 = torch.rand(M, 2, generator=g) * 400, torch.rand(M, 2, generator=g) * 400
    b_ids = torch.randint(0, 2, (M,), generator=g).sort()[0]
    scale1 = torch.tensor([[1.0, 1.0], [1.5, 0.75]])
    for with_scale in (False, True):
        data = {'hw0_i': (416, 416), 'hw0_f': (208, 208), 'mkpts0_c': mk0.to(dev), 'mkpts1_c': mk1.to(dev), 'b_ids': b_ids.to(dev),
                'mconf': torch.rand(M, generator=g).to(dev)}
        if with_scale:
            data['scale0'], data['scale1'] = scale1.to(dev), scale1.to(dev)
        FineMatching().eval()(f0.to(dev), f1.to(dev), data)
        e, k = ofine.fine_match(f0, f1, mk1, 2.0, scale1[b_ids] if with_scale else None)
        assert (data['expec_f'].cpu() - e).abs().max() < 1e-5 and (data['mkpts1_f'].cpu() - k).abs().max() < 1e-3
        assert torch.equal(data['mkpts0_f'].cpu(), mk0)
    data = {'hw0_i': (416, 416), 'hw0_f': (208, 208), 'mkpts0_c': mk0[:0].to(dev), 'mkpts1_c': mk1[:0].to(dev)}
    FineMatching().eval()(f0[:0].to(dev), f1[:0].to(dev), data)
    assert data['expec_f'].shape == (0, 3) and data['mkpts1_f'].shape == (0, 2)


def test_cascade_windows_crossing_the_key_border_match_the_clamped_reference(dev):
    """a row-major 5x5 window that sticks out of the key grid (possible when the caller does not border-shift, or with
    keys smaller than queries) must read the reference's clamped / row-wrapped tokens (torch.clamp of the flat index,
    quadtree_attention.py:428), not zero-filled TMA rows: the tile producer hands such cells to the gather kernel."""
    g = torch.Generator().manual_seed(3)
    B, C, nh, h, w = 1, 128, 4, 48, 64
    q, k, v = (torch.randn(B, C, h, w, generator=g) for _ in range(3))
    hp_, wp_ = h // 2, w // 2
    # window origins: a coherent field shifted so that whole blocks hang over the right / bottom / left / top borders
    py, px = torch.meshgrid(torch.arange(hp_), torch.arange(wp_), indexing='ij')
    r0 = (py - 2 + 3).reshape(-1)               # bottom rows: r0 + 5 > hp
    c0 = (px - 2 + 4).reshape(-1)               # right columns: c0 + 5 > wp
    r0[: 3 * wp_] -= 6                          # top rows: negative origins
    c0[::wp_] -= 7                              # left column: negative origins
    off = torch.arange(5)
    oy, ox = torch.meshgrid(off, off, indexing='ij')
    pos = torch.stack([r0[:, None] + oy.reshape(-1), c0[:, None] + ox.reshape(-1)], -1).unsqueeze(0).contiguous()      # [1,Np,25,2]
    ref_m, ref_up = oqt.cascade_qtatt_b(q, k, v, pos, None, nh)
    m, up = F.cascade_qtatt_forward(q.to(dev), k.to(dev), v.to(dev), pos.to(dev), None, nh)
    assert torch.equal(up.cpu(), ref_up)
    assert (m.cpu() - ref_m).abs().max() < 1e-4


def test_qtatt_b_weight_longer_than_the_pyramid(dev):
    """QTAttB(scale=4) called with a 3-level pyramid: the reference soft-maxes the whole 4-entry parameter and uses entries 0..2
    (quadtree_attention.py:264-282); ADVICE r1: the kernel must not renormalise over the first 3."""
    from casmtr_b200 import QTAttB, synth
    qs, ks, vs, _ = synth.qtatt_inputs(1, 128, 32, 32, 3, seed=77)
    m = QTAttB(4, 32, scale=4, topks=[16, 8, 8]).to(dev).eval()
    with torch.no_grad():
        m.weight.copy_(torch.tensor([0.3, -1.2, 0.8, 2.0]))
    out = m([t.to(dev) for t in qs], [t.to(dev) for t in ks], [t.to(dev) for t in vs])
    ref = oqt.qtatt_b(qs, ks, vs, m.weight.detach().cpu(), [16, 8, 8], 4)
    assert (out.cpu() - ref).abs().max() < 1e-4
    short = oqt.qtatt_b(qs, ks, vs, m.weight.detach().cpu()[:3], [16, 8, 8], 4)
    assert (ref - short).abs().max() > 1e-2      # the two normalisations really differ on this input


def test_cascade_matching_nan_features_stay_in_bounds(dev):
    """with NaN features no score compares equal to the row maximum; the arg-max sentinel must not be used as a
    gather index.  The reference returns a NaN confidence with a valid index."""
    g = torch.Generator().manual_seed(9)
    B, L, C, K = 1, 64, 128, 100
    f0, f1 = torch.randn(B, L, C, generator=g), torch.randn(B, L, C, generator=g)
    f0[0, 5] = float('nan')
    idx = torch.randint(0, L, (B, L, K), generator=g)
    o = F.cascade_match_forward(f0.to(dev), f1.to(dev), idx.to(dev), idx.to(dev))
    torch.cuda.synchronize()                       # an out-of-bounds read would surface here as a sticky error
    assert torch.isnan(o['next_conf01'][0, 5]) and 0 <= int(o['next_idx01'][0, 5]) < L
    ref = ocas.cascade_match(f0, f1, idx, idx)
    ok = torch.ones(L, dtype=torch.bool)
    ok[5] = False
    rows10 = ~(idx[0] == 5).any(-1)                # rows of the other direction that do not touch the NaN token
    assert torch.equal(o['next_idx01'].cpu()[0, ok], ref['next_idx01'][0, ok])
    assert torch.equal(o['next_idx10'].cpu()[0, rows10], ref['next_idx10'][0, rows10])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two devices in one process')
def test_second_device_in_the_same_process():
    """cudaFuncSetAttribute is per device; a process that uses two GPUs must get the > 48 KB dynamic shared memory
    opt-in on both.  Same inputs on cuda:0 and cuda:1 -> identical results."""
    from casmtr_b200 import synth
    qs, ks, vs, wt = synth.qtatt_inputs(1, 256, 64, 64, 3, seed=5)
    c = synth.cascade_inputs(1, 128, 64, 64, seed=6)
    outs = []
    for d in ('cuda:0', 'cuda:1'):
        dv = torch.device(d)
        a = F.qtatt_forward([t.to(dv) for t in qs], [t.to(dv) for t in ks], [t.to(dv) for t in vs], [32, 16, 8], 8, weight=wt.to(dv))
        m, up = F.cascade_qtatt_forward(c['feat0'].to(dv), c['feat1'].to(dv), c['feat1'].to(dv), c['topk_pos01'].to(dv), None, 4)
        f0, f1 = (c[n].flatten(2).transpose(1, 2).contiguous().to(dv) for n in ('feat0', 'feat1'))
        o = F.cascade_match_forward(f0, f1, up, up, w0=64, w1=64)
        torch.cuda.synchronize(dv)
        outs.append((a.cpu(), m.cpu(), o['next_idx01'].cpu()))
    for x, y in zip(*outs):
        assert torch.equal(x, y)


def test_qtatt_guided_single_level(dev):
    """QTAttGuided: drop-in for the one case the reference can run (single-level pyramid), against the reference's own
    output (tests/golden/qtatt_guided.npz) and, in raster order, against the oracle; the layer one level up (QuadtreeAttention with
    attn_type='Guided', scale=1) against the oracle's layer restatement."""
    from golden_util import load
    from casmtr_b200 import QTAttGuided, QuadtreeAttention
    g = load('qtatt_guided')
    nh = int(g['nhead'])
    pos = g['topk_pos'].long()
    q, k, v = g['q'].to(dev), g['k'].to(dev), g['v'].to(dev)
    m = QTAttGuided(nh, q.shape[1] // nh, scale=1, topks=[pos.shape[3]]).to(dev).eval()
    with torch.no_grad():
        m.weight.copy_(torch.tensor([0.7]))
        out = m([q], [k], [v], topk_pos=pos.to(dev))
    assert (out.cpu() - g['out']).abs().max() < 1e-4
    m3 = QTAttGuided(nh, q.shape[1] // nh, scale=3, topks=[pos.shape[3]]).to(dev).eval()
    with torch.no_grad():
        m3.weight.copy_(g['weight3'])
        assert (m3([q], [k], [v], topk_pos=pos.to(dev)).cpu() - g['out3']).abs().max() < 1e-4
        m3.reference_order = False
        ras = m3([q], [k], [v], topk_pos=pos.to(dev))
    assert (ras.cpu() - oqt.qtatt_guided(g['q'], g['k'], g['v'], pos, g['weight3'], nh, reference_order=False)).abs().max() < 1e-4
    with pytest.raises(NotImplementedError):
        m([q, q], [k, k], [v, v], topk_pos=pos.to(dev))
    # other K / head counts, rectangular, keys on a different grid than the queries
    gen = torch.Generator().manual_seed(3)
    for (B, nh2, h0, w0, h1, w1, K) in ((1, 8, 32, 40, 32, 40, 32), (2, 2, 8, 16, 12, 20, 5), (1, 4, 16, 16, 24, 8, 16)):
        C = nh2 * 32
        q2, k2, v2 = torch.randn(B, C, h0, w0, generator=gen), torch.randn(B, C, h1, w1, generator=gen), torch.randn(B, C, h1, w1, generator=gen)
        Np = (h0 // 2) * (w0 // 2)
        p2 = torch.stack([torch.randint(0, h1 // 2, (B, Np, K, nh2), generator=gen), torch.randint(0, w1 // 2, (B, Np, K, nh2), generator=gen)])
        wt = torch.randn(2, generator=gen)
        ref = oqt.qtatt_guided(q2, k2, v2, p2, wt, nh2, reference_order=False)
        got = F.qtatt_guided_forward(q2.to(dev), k2.to(dev), v2.to(dev), p2.to(dev), wt.to(dev), nh2)
        assert (got.cpu() - ref).abs().max() < 1e-4, (B, nh2, h0, w0, K)
    # the attention layer around it
    layer = QuadtreeAttention(128, nh, [pos.shape[3]], scale=1, attn_type='Guided').to(dev).eval()
    x = torch.randn(2, 16 * 24, 128, generator=gen).to(dev)
    with torch.no_grad():
        y = layer(x, x, 16, 24, topk_pos=pos.to(dev))
        sd = {n: t.cpu() for n, t in layer.state_dict().items()}
        xq = torch.nn.functional.conv2d(x.cpu().transpose(1, 2).reshape(2, 128, 16, 24), sd['q_proj.weight'])
        xk = torch.nn.functional.conv2d(x.cpu().transpose(1, 2).reshape(2, 128, 16, 24), sd['k_proj.weight'])
        xv = torch.nn.functional.conv2d(x.cpu().transpose(1, 2).reshape(2, 128, 16, 24), sd['v_proj.weight'])
        msg = oqt.qtatt_guided(xq, xk, xv, pos, sd['py_att.weight'], nh).reshape(2, -1, 128)
        want = torch.nn.functional.linear(msg, sd['proj.weight'], sd['proj.bias'])
    assert (y.cpu() - want).abs().max() < 1e-3
