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
    data = {'hw0_i': (160, 160), 'hw1_i': (160, 160), 'hw0_8c': (20, 20), 'hw1_8c': (20, 20)}
    casmtr_b200.CoarseMatching(cfg).eval()(f0.to(dev), f1.to(dev), data, level='8c')
    c01, i01, c10, i10, g01, g10 = _ref(f0, f1, 0.1)
    st = data['stage_8c']
    assert torch.equal(st['next_idx_c01'].cpu()[g01 > 1e-4], i01[g01 > 1e-4]) and (st['next_conf_c10'].cpu() - c10).abs().max() < 1e-3
    assert st['conf_matrix'] is None


# ---- padding masks (reference coarse_matching.py:64-65: masked_fill_ with -INF = -1e9)
def test_coarse_match_masked_golden(dev):
    """Against the reference's own CoarseMatching run with mask_c0 / mask_c1 (tests/golden/make_golden.py): indices bit-exact
    (padded rows: 0), confidences within 1e-3 (padded rows: the uniform 1 / columns)."""
    from golden_util import load
    g = load('widen_coarse_match_masked')
    m0, m1 = g['mask0'].bool(), g['mask1'].bool()
    o = F.coarse_match_forward(g['feat0'].to(dev), g['feat1'].to(dev), float(g['temperature']), m0.to(dev), m1.to(dev))
    assert torch.equal(o['next_idx01'].cpu(), g['next_idx01']) and torch.equal(o['next_idx10'].cpu(), g['next_idx10'])
    assert (o['next_conf01'].cpu() - g['next_conf01']).abs().max() < 1e-3 and (o['next_conf10'].cpu() - g['next_conf10']).abs().max() < 1e-3
    assert torch.equal(o['next_conf01'].cpu()[~m0], g['next_conf01'][~m0])          # 1 / 192 exactly


@pytest.mark.parametrize('B,hw0,valid0,hw1,valid1,C', [(1, (104, 104), (104, 78), (104, 104), (69, 104), 256),     # 832^2, MegaDepth-style bands
                                                       (2, (20, 30), (13, 30), (24, 18), (24, 18), 64)])
def test_coarse_match_masked_equals_the_valid_sub_problem(dev, B, hw0, valid0, hw1, valid1, C):
    """Size-independent property: with band masks the statistics of the valid rows must equal those of the dense problem on
    the gathered valid tokens (same kernel, no masks), indices mapped back; padded rows are (1 / columns, 0)."""
    g = torch.Generator().manual_seed(hw0[0] + hw1[1])
    L0, L1 = hw0[0] * hw0[1], hw1[0] * hw1[1]
    f0, f1 = torch.randn(B, L0, C, generator=g).to(dev), torch.randn(B, L1, C, generator=g).to(dev)

    def band(hw, valid):
        m = torch.zeros(hw, dtype=torch.bool)
        m[:valid[0], :valid[1]] = True
        return m.reshape(-1)
    v0, v1 = band(hw0, valid0).to(dev), band(hw1, valid1).to(dev)
    o = F.coarse_match_forward(f0, f1, 0.1, v0[None].expand(B, -1), v1[None].expand(B, -1))
    sub = F.coarse_match_forward(f0[:, v0].contiguous(), f1[:, v1].contiguous(), 0.1)
    id0, id1 = torch.nonzero(v0).flatten(), torch.nonzero(v1).flatten()
    assert torch.equal(o['next_idx01'][:, v0], id1[sub['next_idx01']]) and torch.equal(o['next_idx10'][:, v1], id0[sub['next_idx10']])
    assert (o['next_conf01'][:, v0] - sub['next_conf01']).abs().max() < 1e-6 and (o['next_conf10'][:, v1] - sub['next_conf10']).abs().max() < 1e-6
    if (~v0).any():
        assert (o['next_idx01'][:, ~v0] == 0).all() and (o['next_conf01'][:, ~v0] == 1.0 / L1).all()
    if (~v1).any():
        assert (o['next_idx10'][:, ~v1] == 0).all() and (o['next_conf10'][:, ~v1] == 1.0 / L0).all()


def test_coarse_matching_module_with_masks(dev):
    import casmtr_b200
    from golden_util import load
    g = load('widen_coarse_match_masked')
    cfg = {'thr': 0.2, 'border_rm': 2, 'match_type': 'dual_softmax', 'dsmax_temperature': float(g['temperature']), 'train_coarse_percent': 0.3,
           'train_pad_num_gt_min': 200}
    data = {'hw0_i': (96, 128), 'hw1_i': (96, 128), 'hw0_8c': (12, 16), 'hw1_8c': (12, 16)}
    mod = casmtr_b200.CoarseMatching(cfg).eval()
    mod(g['feat0'].to(dev), g['feat1'].to(dev), data, mask_c0=g['mask0'].bool().to(dev), mask_c1=g['mask1'].bool().to(dev))
    st = data['stage_8c']
    assert torch.equal(st['next_idx_c01'].cpu(), g['next_idx01']) and torch.equal(st['next_idx_c10'].cpu(), g['next_idx10'])
    with pytest.raises(RuntimeError):
        mod(g['feat0'].to(dev), g['feat1'].to(dev), data, mask_c0=g['mask0'].bool().to(dev))


# ---- mutual-nearest-neighbour match list of the 1/8 stage (reference coarse_matching.py:91-153), second tensor-core pass
@pytest.mark.parametrize('name,padded', [('widen_coarse_match', False), ('widen_coarse_match_masked', True)])
def test_coarse_match_list_golden(dev, name, padded):
    """Against the list the reference's own CoarseMatching.get_coarse_match produced (tests/golden/make_golden.py): ids bit-exact,
    in torch.where order; mconf within 1e-5; keypoints exact.  The padded fixture carries mask_8c0 / mask_8c1 in `data`, i.e. the
    padded border removal (mask_border_with_padding)."""
    import casmtr_b200
    from golden_util import load
    g = load(name)
    h, w = g['hw'].tolist()
    cfg = {'thr': float(g['thr']), 'border_rm': int(g['border_rm']), 'match_type': 'dual_softmax', 'dsmax_temperature': float(g['temperature']),
           'train_coarse_percent': 0.3, 'train_pad_num_gt_min': 200}
    data = {'hw0_i': (h * 8, w * 8), 'hw1_i': (h * 8, w * 8), 'hw0_8c': (h, w), 'hw1_8c': (h, w)}
    kw = {}
    if padded:
        m0, m1 = g['mask0'].bool().to(dev), g['mask1'].bool().to(dev)
        data['mask_8c0'], data['mask_8c1'] = m0.reshape(-1, h, w), m1.reshape(-1, h, w)
        kw = {'mask_c0': m0, 'mask_c1': m1}
    casmtr_b200.CoarseMatching(cfg).eval()(g['feat0'].to(dev), g['feat1'].to(dev), data, **kw)
    st = data['stage_8c']
    assert st['b_ids'].numel() == g['m_b_ids'].numel() > 20
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(st[k].cpu(), g['m_' + k]), k
    assert (st['mconf'].cpu() - g['m_mconf']).abs().max() < 1e-5
    assert torch.equal(st['mkpts0_c'].cpu(), g['m_mkpts0_c'].float()) and torch.equal(st['mkpts1_c'].cpu(), g['m_mkpts1_c'].float())
    assert torch.equal(st['m_bids'].cpu(), g['m_b_ids']) and not st['gt_mask'].any()


@pytest.mark.parametrize('B,hw0,hw1,C,border', [(2, (40, 52), (40, 52), 256, 2), (1, (104, 104), (104, 104), 256, 0), (1, (30, 40), (24, 36), 64, 1)])
def test_coarse_match_list_vs_oracle(dev, B, hw0, hw1, C, border):
    """Larger random problems against the dense-matrix oracle (oracle/widen.py:coarse_matches); matches whose confidence sits within
    1e-4 of the threshold, or whose row / column maximum of conf is not clear by 1e-4 relative, are excluded from the comparison."""
    import casmtr_b200
    from oracle import widen
    L0, L1 = hw0[0] * hw0[1], hw1[0] * hw1[1]
    g = torch.Generator().manual_seed(L0 + L1 + C)
    f0 = torch.randn(B, L0, C, generator=g)
    n = min(L0, L1)
    f1 = torch.randn(B, L1, C, generator=g)
    perm = torch.randperm(L1, generator=g)[:n]
    f1[:, perm] = 0.8 * f0[:, :n] + 0.6 * f1[:, perm]          # planted correspondences
    thr = 0.2
    ref = widen.coarse_matches(f0, f1, 0.1, thr, border, hw0, hw1, (hw0[0] * 8, hw0[1] * 8))
    cfg = {'thr': thr, 'border_rm': border, 'match_type': 'dual_softmax', 'dsmax_temperature': 0.1, 'train_coarse_percent': 0.3,
           'train_pad_num_gt_min': 200}
    data = {'hw0_i': (hw0[0] * 8, hw0[1] * 8), 'hw1_i': (hw1[0] * 8, hw1[1] * 8), 'hw0_8c': hw0, 'hw1_8c': hw1}
    casmtr_b200.CoarseMatching(cfg).eval()(f0.to(dev), f1.to(dev), data)
    st = data['stage_8c']
    got = {(int(b), int(i)): (int(j), float(c)) for b, i, j, c in zip(st['b_ids'].cpu(), st['i_ids'].cpu(), st['j_ids'].cpu(), st['mconf'].cpu())}
    want = {(int(b), int(i)): (int(j), float(c)) for b, i, j, c in zip(ref['b_ids'], ref['i_ids'], ref['j_ids'], ref['mconf'])}
    assert len(want) > 100                                       # (at 10816 tokens the dual soft-max is thin: ~600 confident matches)
    for key in set(got) | set(want):
        c = (got.get(key) or want.get(key))[1]
        if abs(c - thr) < 1e-4:
            continue                                            # at the threshold: fp32 rounding decides
        assert key in got and key in want, key
        assert got[key][0] == want[key][0] and abs(got[key][1] - want[key][1]) < 1e-4, key
    order = st['b_ids'].cpu() * L0 + st['i_ids'].cpu()
    assert (order[1:] > order[:-1]).all()                       # torch.where order
