"""GPU A/B against the reference's OWN CUDA kernels: oracle/_ref/ holds the three reference extensions built unmodified
from /root/reference (oracle/build_ref.py, sm_100a).  Skipped when they have not been built.
  * op level: libcasmtr_b200's drop-ins (C ABI) vs score_computation_cuda / value_aggregation_cuda /
    fast_score_computation on identical tensors (fp32, only the summation order differs: 1e-4 abs);
  * module level: the fused CascadeQTAttB / CascadeMatching correlation results vs the same quantities recomputed with
    the reference kernels from the indices the fused path reports (scores within 1e-3 abs, BASELINE.json tolerance)."""
import pytest
import torch

import casmtr_b200
from casmtr_b200 import functional as F
from casmtr_b200 import synth
from oracle import build_ref

pytestmark = pytest.mark.gpu
NAMES = ('score_computation_cuda', 'value_aggregation_cuda', 'fast_score_computation')


@pytest.fixture(scope='module')
def ref():
    if not all(build_ref.built(n) for n in NAMES):
        pytest.skip('oracle/_ref not built (python -m oracle.build_ref in the build container)')
    return {n: build_ref.load(n) for n in NAMES}


@pytest.mark.parametrize('B,N1,N2,H,D,K', [(2, 12, 48, 4, 32, 20), (1, 676, 2704, 8, 32, 128), (1, 7, 30, 3, 16, 5)])
def test_score5d_vs_reference_kernel(dev, ref, B, N1, N2, H, D, K):
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B, N1, 4, H, D, generator=g).to(dev)
    k = torch.randn(B, N2, H, D, generator=g).to(dev)
    idx = torch.randint(0, N2, (B, N1, K, H), generator=g).to(dev)
    want = ref['score_computation_cuda'].score_forward(q, k, idx)[0]
    torch.cuda.synchronize()
    got = F.score5d(q, k, idx)
    assert got.shape == want.shape and (got - want).abs().max() < 1e-4


@pytest.mark.parametrize('B,N,K,H,M,D', [(2, 40, 16, 4, 50, 32), (1, 2704, 64, 8, 2704, 32)])
def test_value_agg_vs_reference_kernel(dev, ref, B, N, K, H, M, D):
    g = torch.Generator().manual_seed(2)
    s = torch.rand(B, N, K, H, generator=g).to(dev)
    v = torch.randn(B, M, H, D, generator=g).to(dev)
    idx = torch.randint(0, M, (B, N, K, H), generator=g).to(dev)
    want = torch.zeros(B, N, H, D, device=dev)
    ref['value_aggregation_cuda'].value_aggregation_forward(s, v, idx, want)
    got = F.value_agg(s, v, idx)
    assert (got - want).abs().max() < 1e-4


@pytest.mark.parametrize('B,N1,N2,C,K', [(2, 48, 50, 64, 10), (1, 4096, 4096, 128, 100)])
def test_score3d_vs_reference_kernel(dev, ref, B, N1, N2, C, K):
    g = torch.Generator().manual_seed(3)
    q = torch.randn(B, N1, C, generator=g).to(dev)
    k = torch.randn(B, N2, C, generator=g).to(dev)
    idx = torch.randint(0, N2, (B, N1, K), generator=g).to(dev)
    want = ref['fast_score_computation'].score_forward(q, k, idx)[0]
    torch.cuda.synchronize()
    got = F.score3d(q, k, idx)
    assert (got - want).abs().max() < 2e-4


def test_cascade_stage_vs_reference_kernels(dev, ref):
    """Fused CascadeQTAttB + CascadeMatching (TMA-tiled kernels) against the reference kernels fed with the
    upsampled_idx the fused attention returns: attention message and confidence volume within 1e-3, argmax bit-exact."""
    B, nh, h, w = 1, 4, 64, 64
    C = nh * 32
    d = synth.cascade_inputs(B, C, h, w, seed=77)
    v = torch.randn(B, C, h, w, generator=torch.Generator().manual_seed(5))
    q, k, v, tp = d['feat0'].to(dev), d['feat1'].to(dev), v.to(dev), d['topk_pos01'].to(dev)
    msg, up = casmtr_b200.CascadeQTAttB(nh, 32, dilated=1)(q, k, v, tp, None)
    # reference formulation of the same attention from its own kernels (modules/quadtree_attention.py:431-447)
    tok = lambda x: x.flatten(2).transpose(1, 2).reshape(B, h * w, nh, 32).contiguous()
    qc = q.reshape(B, nh, 32, h // 2, 2, w // 2, 2).permute(0, 3, 5, 4, 6, 1, 2).reshape(B, (h // 2) * (w // 2), 4, nh, 32).contiguous()
    cell_rows = up.reshape(B, h // 2, 2, w // 2, 2, 100)[:, :, 0, :, 0].reshape(B, (h // 2) * (w // 2), 100)   # one list per parent cell
    idx = cell_rows.unsqueeze(-1).expand(-1, -1, -1, nh).contiguous()
    qk = ref['score_computation_cuda'].score_forward(qc, tok(k), idx)[0] / 32 ** 0.5          # [B, Np, 4, 100, nh]
    torch.cuda.synchronize()
    A = torch.softmax(qk, dim=-2)
    out = torch.zeros(B, (h // 2) * (w // 2) * 4, nh, 32, device=dev)
    ref['value_aggregation_cuda'].value_aggregation_forward(A.reshape(B, -1, 100, nh).contiguous(), tok(v),
                                                            idx.unsqueeze(2).expand(-1, -1, 4, -1, -1).reshape(B, -1, 100, nh).contiguous(), out)
    want = out.reshape(B, h // 2, w // 2, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B, h * w, C)
    assert (msg - want).abs().max() < 1e-3
    # correlation volume of the matching stage
    f0, f1 = q.flatten(2).transpose(1, 2).contiguous(), k.flatten(2).transpose(1, 2).contiguous()
    _, up10 = casmtr_b200.CascadeQTAttB(nh, 32, dilated=1)(k, q, v, d['topk_pos10'].to(dev), None)
    o = F.cascade_match_forward(f0, f1, up, up10, w0=w, w1=w)
    sim = ref['fast_score_computation'].score_forward(f0 / C ** 0.5, f1 / C ** 0.5, up)[0]
    torch.cuda.synchronize()
    conf = torch.softmax(sim, dim=2)
    assert (o['conf01'] - conf).abs().max() < 1e-3
    gap = conf.topk(2, dim=2)[0]
    clear = (gap[..., 0] - gap[..., 1]) > 1e-6                                              # rows without an fp32 near-tie
    want_idx = torch.gather(up, 2, conf.argmax(dim=2, keepdim=True)).squeeze(-1)
    assert torch.equal(o['next_idx01'][clear], want_idx[clear])


def test_reference_data_flow_restated(dev, ref):
    """oracle/ref_path.py (the reference's QTAttB / CascadeQTAttB flow around its own kernels, used as the GPU baseline in
    bench.py) computes what the oracle and the fused CUDA path compute."""
    from oracle import qtatt as oqt, ref_path
    B, nh, h, w, topks = 1, 4, 32, 48, [8, 8, 4]
    qs, ks, vs, wt = synth.qtatt_inputs(B, nh * 32, h, w, 3, seed=61)
    want = oqt.qtatt_b(qs, ks, vs, wt, topks, nh)
    got = ref_path.qtatt_b([x.to(dev) for x in qs], [x.to(dev) for x in ks], [x.to(dev) for x in vs], wt.to(dev), topks, nh)
    ours = F.qtatt_forward([x.to(dev) for x in qs], [x.to(dev) for x in ks], [x.to(dev) for x in vs], topks, nh, weight=wt.to(dev))
    assert (got.cpu() - want).abs().max() < 1e-4 and (got - ours).abs().max() < 1e-4
    d = synth.cascade_inputs(1, nh * 32, 32, 32, seed=62)
    v = torch.randn(1, nh * 32, 32, 32, generator=torch.Generator().manual_seed(4))
    wm, wu = oqt.cascade_qtatt_b(d['feat0'], d['feat1'], v, d['topk_pos01'], None, nh)
    gm, gu = ref_path.cascade_qtatt_b(d['feat0'].to(dev), d['feat1'].to(dev), v.to(dev), d['topk_pos01'].to(dev), nh)
    assert torch.equal(gu.cpu(), wu) and (gm.cpu() - wm).abs().max() < 1e-4


# ---- backward halves against the reference's own backward kernels (score_backward / value_aggregation_backward)
def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize('B,N1,N2,H,D,K', [(2, 12, 48, 4, 32, 20), (1, 169, 676, 8, 32, 128)])
def test_score5d_backward_vs_reference_kernel(dev, ref, B, N1, N2, H, D, K):
    g = torch.Generator().manual_seed(11)
    q = torch.randn(B, N1, 4, H, D, generator=g).to(dev)
    k = torch.randn(B, N2, H, D, generator=g).to(dev)
    idx = torch.randint(0, N2, (B, N1, K, H), generator=g).to(dev)
    go = torch.randn(B, N1, 4, K, H, generator=g).to(dev)
    wq, wk = ref['score_computation_cuda'].score_backward(go, q, k, idx)
    torch.cuda.synchronize()
    gq, gk = F.score5d_backward(go, q, k, idx)
    assert _rel(gq, wq) < 1e-5 and _rel(gk, wk) < 1e-5


@pytest.mark.parametrize('B,N,K,H,M,D', [(2, 40, 16, 4, 50, 32), (1, 676, 64, 8, 676, 32)])
def test_value_agg_backward_vs_reference_kernel(dev, ref, B, N, K, H, M, D):
    g = torch.Generator().manual_seed(12)
    s = torch.rand(B, N, K, H, generator=g).to(dev)
    v = torch.randn(B, M, H, D, generator=g).to(dev)
    idx = torch.randint(0, M, (B, N, K, H), generator=g).to(dev)
    go = torch.randn(B, N, H, D, generator=g).to(dev)
    ws, wv = torch.zeros_like(s), torch.zeros_like(v)
    ref['value_aggregation_cuda'].value_aggregation_backward(go, s, v, idx, ws, wv)
    gs, gv = F.value_agg_backward(go, s, v, idx)
    assert _rel(gs, ws) < 1e-5 and _rel(gv, wv) < 1e-5


@pytest.mark.parametrize('B,N1,N2,C,K', [(2, 48, 50, 64, 10), (1, 1024, 1024, 128, 100)])
def test_score3d_backward_vs_reference_kernel(dev, ref, B, N1, N2, C, K):
    g = torch.Generator().manual_seed(13)
    q = torch.randn(B, N1, C, generator=g).to(dev)
    k = torch.randn(B, N2, C, generator=g).to(dev)
    idx = torch.randint(0, N2, (B, N1, K), generator=g).to(dev)
    go = torch.randn(B, N1, K, generator=g).to(dev)
    wq, wk = ref['fast_score_computation'].score_backward(go, q, k, idx)
    torch.cuda.synchronize()
    gq, gk = F.score3d_backward(go, q, k, idx)
    assert _rel(gq, wq) < 1e-5 and _rel(gk, wk) < 1e-5
