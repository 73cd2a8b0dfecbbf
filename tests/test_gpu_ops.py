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
