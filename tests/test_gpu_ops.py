"""GPU parity: op-level drop-ins (C ABI via casmtr_b200.functional) vs the CPU oracle."""
import pytest
import torch

from casmtr_b200 import functional as F
from oracle import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,N1,N2,H,D,K', [(2, 12, 48, 4, 32, 20), (1, 169, 676, 8, 32, 128), (1, 7, 30, 3, 16, 5), (1, 5, 9, 12, 8, 33)])
def test_score5d(dev, B, N1, N2, H, D, K):
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B, N1, 4, H, D, generator=g)
    k = torch.randn(B, N2, H, D, generator=g)
    idx = torch.randint(0, N2, (B, N1, K, H), generator=g)
    ref = ops.score5d(q, k, idx)
    out = F.score5d(q.to(dev), k.to(dev), idx.to(dev)).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max() < 1e-4          # fp32, only the summation order differs


@pytest.mark.parametrize('B,N,K,H,M,D', [(2, 40, 16, 4, 50, 32), (1, 300, 64, 8, 333, 32), (1, 9, 5, 2, 11, 8)])
def test_value_agg(dev, B, N, K, H, M, D):
    g = torch.Generator().manual_seed(2)
    s = torch.rand(B, N, K, H, generator=g)
    v = torch.randn(B, M, H, D, generator=g)
    idx = torch.randint(0, M, (B, N, K, H), generator=g)
    ref = ops.value_agg(s, v, idx)
    out = torch.full((B, N, H, D), 7.0, device=dev)      # not pre-zeroed on purpose: must be overwritten
    F.value_agg(s.to(dev), v.to(dev), idx.to(dev), out)
    assert (out.cpu() - ref).abs().max() < 1e-4


@pytest.mark.parametrize('B,N1,N2,C,K', [(2, 40, 50, 64, 10), (1, 256, 256, 128, 100), (1, 33, 70, 256, 7), (1, 16, 16, 12, 3)])
def test_score3d(dev, B, N1, N2, C, K):
    g = torch.Generator().manual_seed(3)
    q = torch.randn(B, N1, C, generator=g)
    k = torch.randn(B, N2, C, generator=g)
    idx = torch.randint(0, N2, (B, N1, K), generator=g)
    ref = ops.score3d(q, k, idx)
    out = F.score3d(q.to(dev), k.to(dev), idx.to(dev)).cpu()
    assert (out - ref).abs().max() < 2e-4


def test_nchw_to_tokens(dev):
    x = torch.randn(2, 96, 13, 17)
    out = F.nchw_to_tokens(x.to(dev)).cpu()
    assert torch.equal(out, x.flatten(2).transpose(1, 2).contiguous())


def test_cpu_tensor_rejected():
    with pytest.raises(RuntimeError):
        F.score3d(torch.zeros(1, 4, 8), torch.zeros(1, 4, 8), torch.zeros(1, 4, 2, dtype=torch.long))


# ---- backward halves (SURVEY 8f "next" #4): against torch autograd through the oracle's forward restatements (fp64 on the CPU).
# The scattered gradients (key / value) sum thousands of terms in an order the hardware picks: tolerance relative to the
# largest gradient, 1e-5 (observed ~1e-6); the gathered ones (query / score) 1e-4 abs like the forward.
def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize('B,N1,N2,H,D,K', [(2, 12, 48, 4, 32, 20), (1, 169, 676, 8, 32, 128), (1, 7, 30, 3, 16, 5), (1, 40, 9, 8, 32, 33),
                                           (1, 5, 9, 12, 8, 33)])
def test_score5d_backward(dev, B, N1, N2, H, D, K):
    g = torch.Generator().manual_seed(4)
    q = torch.randn(B, N1, 4, H, D, generator=g)
    k = torch.randn(B, N2, H, D, generator=g)
    idx = torch.randint(0, N2, (B, N1, K, H), generator=g)            # N2 < N1 * K: many query rows hit the same key row
    go = torch.randn(B, N1, 4, K, H, generator=g)
    qd, kd = q.double().requires_grad_(), k.double().requires_grad_()
    ops.score5d(qd, kd, idx).backward(go.double())
    gq, gk = F.score5d_backward(go.to(dev), q.to(dev), k.to(dev), idx.to(dev))
    assert (gq.cpu() - qd.grad.float()).abs().max() < 1e-4 and _rel(gk.cpu(), kd.grad.float()) < 1e-5


@pytest.mark.parametrize('B,N,K,H,M,D', [(2, 40, 16, 4, 50, 32), (1, 300, 64, 8, 333, 32), (1, 9, 5, 2, 11, 8), (1, 64, 12, 3, 7, 64)])
def test_value_agg_backward(dev, B, N, K, H, M, D):
    g = torch.Generator().manual_seed(5)
    s = torch.rand(B, N, K, H, generator=g)
    v = torch.randn(B, M, H, D, generator=g)
    idx = torch.randint(0, M, (B, N, K, H), generator=g)
    go = torch.randn(B, N, H, D, generator=g)
    sd, vd = s.double().requires_grad_(), v.double().requires_grad_()
    ops.value_agg(sd, vd, idx).backward(go.double())
    gs, gv = F.value_agg_backward(go.to(dev), s.to(dev), v.to(dev), idx.to(dev))
    assert (gs.cpu() - sd.grad.float()).abs().max() < 1e-4 and _rel(gv.cpu(), vd.grad.float()) < 1e-5


@pytest.mark.parametrize('B,N1,N2,C,K', [(2, 40, 50, 64, 10), (1, 256, 256, 128, 100), (1, 33, 70, 256, 7), (1, 16, 16, 12, 3), (1, 8, 8, 512, 4)])
def test_score3d_backward(dev, B, N1, N2, C, K):
    g = torch.Generator().manual_seed(6)
    q = torch.randn(B, N1, C, generator=g)
    k = torch.randn(B, N2, C, generator=g)
    idx = torch.randint(0, N2, (B, N1, K), generator=g)
    go = torch.randn(B, N1, K, generator=g)
    qd, kd = q.double().requires_grad_(), k.double().requires_grad_()
    ops.score3d(qd, kd, idx).backward(go.double())
    gq, gk = F.score3d_backward(go.to(dev), q.to(dev), k.to(dev), idx.to(dev))
    assert (gq.cpu() - qd.grad.float()).abs().max() < 2e-4 and _rel(gk.cpu(), kd.grad.float()) < 1e-5


def test_autograd_functions_train_the_reference_formulation(dev):
    """The reference's op-level autograd API (functions/quadtree_attention.py, cascade_functions.py) end to end: gradients of a
    fine-level attention step written with score_computation_op / value_aggregation_op / ScoreComputation match the same step
    written with torch gathers (autograd of the oracle ops) on the same device."""
    import casmtr_b200
    from casmtr_b200.functions.quadtree_attention import score_computation_op, value_aggregation_op
    g = torch.Generator().manual_seed(7)
    B, N1, N2, H, D, K = 1, 36, 144, 4, 32, 16
    q0 = torch.randn(B, N1, 4, H, D, generator=g).to(dev)
    k0 = torch.randn(B, N2, H, D, generator=g).to(dev)
    v0 = torch.randn(B, N2, H, D, generator=g).to(dev)
    idx = torch.randint(0, N2, (B, N1, K, H), generator=g).to(dev)
    idx5 = idx.view(B, N1, 1, K, H).repeat(1, 1, 4, 1, 1)

    def step(score_op, agg_op, q, k, v):
        a = torch.softmax(score_op(q, k, idx) / D ** 0.5, dim=-2)             # [B,N1,4,K,H]
        return agg_op(a, v, idx5)                                             # [B,N1,4,H,D]
    w = torch.randn(B, N1, 4, H, D, generator=g).to(dev)
    grads = []
    for ops_ in ((score_computation_op, value_aggregation_op), (ops.score5d, ops.value_agg5)):
        q, k, v = q0.clone().requires_grad_(), k0.clone().requires_grad_(), v0.clone().requires_grad_()
        (step(*ops_, q, k, v) * w).sum().backward()
        grads.append((q.grad, k.grad, v.grad))
    for a, b in zip(*grads):
        assert _rel(a, b) < 1e-5
    qf, kf = torch.randn(B, 64, 128, generator=g).to(dev).requires_grad_(), torch.randn(B, 50, 128, generator=g).to(dev).requires_grad_()
    i3 = torch.randint(0, 50, (B, 64, 10), generator=g).to(dev)
    casmtr_b200.ScoreComputation.apply(qf, kf, i3).square().sum().backward()
    q2, k2 = qf.detach().clone().requires_grad_(), kf.detach().clone().requires_grad_()
    ops.score3d(q2, k2, i3).square().sum().backward()
    assert _rel(qf.grad, q2.grad) < 1e-5 and _rel(kf.grad, k2.grad) < 1e-5
