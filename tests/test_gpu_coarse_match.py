"""GPU parity of the tcgen05 dense coarse-matching statistics (SURVEY section 8f "next" #1) against the reference
formulation (src/model/functions/coarse_matching.py:60-75) evaluated with torch on the CPU in fp64/fp32."""
import pytest
import torch

from casmtr_b200 import functional as F

pytestmark = pytest.mark.gpu


def _ref(f0, f1, T):
    C = f0.shape[-1]
    sim = torch.einsum('nlc,nsc->nls', f0.double() / C ** 0.5, f1.double() / C ** 0.5) / T
    p01, p10 = torch.softmax(sim, 2), torch.softmax(sim, 1)
    c01, i01 = p01.max(dim=2)
    c10, i10 = p10.max(dim=1)
    # gap between the best and the second best logit of every row / column: arg-max is only specified where it is clear
    t2 = sim.topk(2, dim=2)[0]
    s2 = sim.topk(2, dim=1)[0]
    return c01.float(), i01, c10.float(), i10, (t2[..., 0] - t2[..., 1]), (s2[:, 0] - s2[:, 1])


@pytest.mark.parametrize('B,L0,L1,C', [(1, 256, 256, 256), (2, 300, 520, 256), (1, 1024, 768, 64), (1, 10816, 10816, 256)])
def test_coarse_match_stats(dev, B, L0, L1, C):
    g = torch.Generator().manual_seed(L0 + L1)
    f0 = torch.randn(B, L0, C, generator=g)
    f1 = torch.randn(B, L1, C, generator=g)
    if L0 == L1:                                    # plant real correspondences so that the soft-max is peaked like in the model
        perm = torch.randperm(L1, generator=g)
        f1 = 0.8 * f0[:, perm] + 0.6 * f1
    T = 0.1
    c01, i01, c10, i10, g01, g10 = _ref(f0, f1, T)
    o = F.coarse_match_forward(f0.to(dev), f1.to(dev), T)
    assert (o['next_conf01'].cpu() - c01).abs().max() < 1e-3 and (o['next_conf10'].cpu() - c10).abs().max() < 1e-3
    clear01, clear10 = g01 > 1e-4, g10 > 1e-4
    assert torch.equal(o['next_idx01'].cpu()[clear01], i01[clear01])
    assert torch.equal(o['next_idx10'].cpu()[clear10], i10[clear10])
    assert clear01.float().mean() > 0.99


def test_coarse_matching_module(dev):
    import casmtr_b200
    cfg = {'thr': 0.2, 'border_rm': 0, 'match_type': 'dual_softmax', 'dsmax_temperature': 0.1, 'train_coarse_percent': 0.3,
           'train_pad_num_gt_min': 200}
    g = torch.Generator().manual_seed(5)
    f0 = torch.randn(1, 400, 256, generator=g)
    f1 = 0.8 * f0[:, torch.randperm(400, generator=g)] + 0.6 * torch.randn(1, 400, 256, generator=g)
    data = {}
    casmtr_b200.CoarseMatching(cfg).eval()(f0.to(dev), f1.to(dev), data, level='8c')
    c01, i01, c10, i10, g01, g10 = _ref(f0, f1, 0.1)
    st = data['stage_8c']
    assert torch.equal(st['next_idx_c01'].cpu()[g01 > 1e-4], i01[g01 > 1e-4]) and (st['next_conf_c10'].cpu() - c10).abs().max() < 1e-3
    assert st['conf_matrix'] is None
