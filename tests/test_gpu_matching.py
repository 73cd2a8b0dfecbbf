"""GPU parity: fused cascade matching, NMS + match extraction and fine matching vs the CPU oracle.
Integer outputs (next_idx, masks, match lists) bit-exact; fp within 1e-3 abs (observed ~1e-6)."""
import pytest
import torch

import casmtr_b200
from casmtr_b200 import functional as F
from casmtr_b200 import synth
from oracle import cascade, fine, qtatt

pytestmark = pytest.mark.gpu


def _stage(B, C, h, w, seed, pad=False):
    d = synth.cascade_inputs(B, C, h, w, seed=seed, pad=pad)
    idx01 = qtatt.cascade_window_idx(d['topk_pos01'], h, w)
    idx10 = qtatt.cascade_window_idx(d['topk_pos10'], h, w)
    up = lambda c: qtatt.quad_to_raster(c.reshape(B, 1, -1, 1, 100).expand(B, 1, -1, 4, 100), h // 2, w // 2).reshape(B, h * w, 100).contiguous()
    d['idx01'], d['idx10'] = up(idx01), up(idx10)
    d['f0'] = d['feat0'].flatten(2).transpose(1, 2).contiguous()
    d['f1'] = d['feat1'].flatten(2).transpose(1, 2).contiguous()
    return d


@pytest.mark.parametrize('grid', [False, True, 'mixed'])
@pytest.mark.parametrize('B,C,h,w,pad', [(2, 128, 32, 32, False), (1, 64, 32, 48, False), (2, 128, 32, 32, True), (1, 256, 16, 16, False)])
def test_cascade_match(dev, B, C, h, w, pad, grid):
    """grid=False: one warp per query row; True: the sibling-sharing kernel (2x2 queries of a parent cell share their
    candidate rows); 'mixed': some cells get different lists per sibling, which the kernel must detect and handle."""
    d = _stage(B, C, h, w, 31, pad)
    if grid == 'mixed':
        g = torch.Generator().manual_seed(77)
        hit = torch.rand(B, h * w, generator=g) < 0.3
        d['idx01'][hit] = torch.randint(0, h * w, (int(hit.sum()), 100), generator=g)
        d['idx10'][:, 5] = torch.randint(0, h * w, (B, 100), generator=g)
    m0 = d['mask0'].flatten(1) if pad else None
    m1 = d['mask1'].flatten(1) if pad else None
    ref = cascade.cascade_match(d['f0'], d['f1'], d['idx01'], d['idx10'], m0, m1, 1.0)
    out = F.cascade_match_forward(d['f0'].to(dev), d['f1'].to(dev), d['idx01'].to(dev), d['idx10'].to(dev),
                                  None if m0 is None else m0.to(dev), None if m1 is None else m1.to(dev), 1.0,
                                  w0=w if grid else 0, w1=w if grid else 0)
    for t in ('01', '10'):
        assert torch.equal(out['next_idx' + t].cpu(), ref['next_idx' + t])
        assert (out['next_conf' + t].cpu() - ref['next_conf' + t]).abs().max() < 1e-5
        assert (out['conf' + t].cpu() - ref['conf' + t]).abs().max() < 1e-5


@pytest.mark.parametrize('B,h,w,nms,pad,border,scales,thr', [
    (2, 32, 32, 5, False, 2, False, 0.2),
    (2, 32, 48, None, False, 1, False, 0.2),
    (2, 32, 32, 5, True, 2, True, 0.2),
    (1, 32, 32, 3, False, 0, False, 0.1),
    (3, 16, 16, 5, False, 2, False, 2.0),       # nothing survives -> fallback keeps element 0 of every sample
])
def test_match_extract(dev, B, h, w, nms, pad, border, scales, thr):
    d = _stage(B, 128, h, w, 41, pad)
    o = cascade.cascade_match(d['f0'], d['f1'], d['idx01'], d['idx10'], None, None, 1.0)
    g = torch.Generator().manual_seed(9)
    s0 = torch.rand(B, 2, generator=g) + 0.5 if scales else None
    s1 = torch.rand(B, 2, generator=g) + 0.5 if scales else None
    kw = dict(test_thr=thr, border_rm=border, nms_window=nms, pre_thrs=[0.2], double_check=True)
    ref = cascade.extract_matches(o['next_conf01'], o['next_idx01'], o['next_idx10'], (h, w), (h, w), (h * 4, w * 4),
                                  pre_confs=[(d['pre_conf01'], h // 2, w // 2)],
                                  pad_mask0=d.get('mask0'), pad_mask1=d.get('mask1'), scale0=s0, scale1=s1, **kw)
    c = lambda t: None if t is None else t.to(dev)
    out = F.match_extract(c(o['next_conf01']), c(o['next_idx01']), c(o['next_idx10']), (h, w), (h, w), (h * 4, w * 4),
                          pre_confs=[(c(d['pre_conf01']), h // 2, w // 2)],
                          pad_mask0=c(d.get('mask0')), pad_mask1=c(d.get('mask1')), scale0=c(s0), scale1=c(s1), **kw)
    for k in ('b_ids', 'i_ids', 'j_ids', 'mask'):
        assert torch.equal(out[k].cpu(), ref[k]), k
    for k in ('mconf', 'mkpts0_c', 'mkpts1_c'):
        assert torch.equal(out[k].cpu(), ref[k].to(torch.float32)), k
    assert len(ref['b_ids']) > 0


def test_nms_ties(dev):
    """maxpool NMS on a heavily tied map: first maximum in row-major window order wins."""
    conf = (torch.randint(0, 4, (3, 20 * 30)).float() / 4).contiguous()
    ref = cascade.nms_mask(conf, 20, 30, 5, 0.1)
    zeros = torch.zeros(3, 600, dtype=torch.int64, device=dev)
    out = F.match_extract(conf.to(dev), zeros, zeros, (20, 30), (20, 30), (80, 120), test_thr=0.1, border_rm=0,
                          nms_window=5, double_check=False)
    assert torch.equal(out['mask'].cpu(), ref)


def test_cascade_matching_module(dev):
    """Module API: data dict in, data['stage_4c'] out, same keys as the reference."""
    B, h, w = 2, 32, 32
    d = _stage(B, 128, h, w, 51)
    cfg = {'thr': 0.0101, 'test_thr': 0.2, 'pre_thr': [0.2], 'border_rm': 2, 'double_check': True,
           'train_pad_num_gt_min': 4096, 'match_type': 'softmax', 'dsmax_temperature': 1.0}
    cas = {'propagation': 'window', 'dilated': 1, 'detector_mode': None, 'grid_size': 4,
           'post_config': {'method': 'maxpool_nms', 'window_size': 5, 'topk': None, 'rt': None, 'rd': None}}
    mod = casmtr_b200.CascadeMatching(cfg, cas).eval()
    data = {'hw0_i': (h * 4, w * 4), 'hw1_i': (h * 4, w * 4), 'hw0_4c': (h, w), 'hw1_4c': (h, w),
            'hw0_8c': (h // 2, w // 2), 'hw1_8c': (h // 2, w // 2), 'bs': B,
            'stage_8c': {'next_conf_c01': d['pre_conf01'].to(dev)}}
    mod(d['f0'].to(dev), d['f1'].to(dev), d['idx01'].to(dev), d['idx10'].to(dev), data, level='4c', pre_level='8c')
    o = cascade.cascade_match(d['f0'], d['f1'], d['idx01'], d['idx10'], None, None, 1.0)
    ref = cascade.extract_matches(o['next_conf01'], o['next_idx01'], o['next_idx10'], (h, w), (h, w), (h * 4, w * 4),
                                  test_thr=0.2, border_rm=2, nms_window=5, pre_confs=[(d['pre_conf01'], h // 2, w // 2)],
                                  pre_thrs=[0.2], double_check=True)
    st = data['stage_4c']
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(st[k].cpu(), ref[k])
    assert torch.equal(st['mkpts0_c'].cpu(), ref['mkpts0_c'].float())
    assert (st['conf_matrix'].cpu() - o['conf01']).abs().max() < 1e-5
    assert torch.equal(data['m_bids'].cpu(), ref['b_ids'])


@pytest.mark.parametrize('M,WW,C', [(1, 25, 64), (37, 25, 64), (1000, 25, 128), (5, 9, 32)])
def test_fine_match(dev, M, WW, C):
    f0, f1 = synth.fine_inputs(M, WW, C, seed=61)
    mk = torch.rand(M, 2) * 100
    ref_e, ref_k = fine.fine_match(f0, f1, mk, 2.0)
    e, k = F.fine_match_forward(f0.to(dev), f1.to(dev), mk.to(dev), 2.0)
    assert (e.cpu() - ref_e).abs().max() < 1e-5
    assert (k.cpu() - ref_k).abs().max() < 1e-4


def test_fine_matching_module_empty(dev):
    mod = casmtr_b200.CascadeFineMatching('4c').eval()
    data = {'hw0_i': (128, 128), 'hw0_f': (64, 64),
            'stage_4c': {'mkpts0_c': torch.zeros(0, 2, device=dev), 'mkpts1_c': torch.zeros(0, 2, device=dev),
                         'mconf': torch.zeros(0, device=dev), 'b_ids': torch.zeros(0, dtype=torch.long, device=dev)}}
    mod(torch.zeros(0, 25, 64, device=dev), torch.zeros(0, 25, 64, device=dev), data)
    assert data['expec_f'].shape == (0, 3)


def test_pack_matches_kernel(dev):
    """The packed all-gather block written by the library == the torch reference packing (casmtr_b200.dist.pack_matches)."""
    from casmtr_b200 import dist as cdist
    g = torch.Generator().manual_seed(3)
    M, cap = 37, 64
    m = {'b_ids': torch.randint(0, 3, (M,), generator=g).sort()[0], 'i_ids': torch.randint(0, 9999, (M,), generator=g),
         'j_ids': torch.randint(0, 9999, (M,), generator=g), 'mconf': torch.rand(M, generator=g),
         'mkpts0': torch.rand(M, 2, generator=g) * 800, 'mkpts1': torch.rand(M, 2, generator=g) * 800}
    block = F.pack_matches({k: v.to(dev) for k, v in m.items()}, pair_offset=5, cap=cap).cpu()
    assert block.shape == (cap + 1, 44)
    assert int(block[0, :8].clone().view(torch.int64)) == M
    assert torch.equal(block[1:M + 1], cdist.pack_matches(m, pair_offset=5))
    back = cdist.unpack_gathered(block.unsqueeze(0))
    assert torch.equal(back['b_ids'], m['b_ids'] + 5) and torch.equal(back['mkpts1'], m['mkpts1'])
    empty = F.pack_matches({k: v[:0].to(dev) for k, v in m.items()}, pair_offset=0, cap=4).cpu()
    assert int(empty[0, :8].clone().view(torch.int64)) == 0


@pytest.mark.parametrize('cat', [False, True])
def test_fine_preprocess_vs_unfold(dev, cat):
    """CascadeFinePreprocess: gathered windows == the reference formulation F.unfold(...)[b_ids, ids] (fine_matching.py:47-55)."""
    import torch.nn.functional as tF
    B, C, Hc, Wc, stride, W = 2, 64, 24, 40, 2, 5
    g = torch.Generator().manual_seed(8)
    ff0, ff1 = torch.randn(B, C, Hc * stride, Wc * stride, generator=g), torch.randn(B, C, Hc * stride, Wc * stride, generator=g)
    fc0, fc1 = torch.randn(B, Hc * Wc, 128, generator=g), torch.randn(B, Hc * Wc, 128, generator=g)
    M = 300
    b = torch.randint(0, B, (M,), generator=g).sort()[0]
    i, j = torch.randint(0, Hc * Wc, (M,), generator=g), torch.randint(0, Hc * Wc, (M,), generator=g)
    i[:4] = torch.tensor([0, Wc - 1, (Hc - 1) * Wc, Hc * Wc - 1])                  # corners: zero padding
    mod = casmtr_b200.CascadeFinePreprocess({'fine_concat_coarse_feat': cat, 'fine_window_size': W}, {'d_model': C}, {'d_model': 128}, '4c').eval()
    data = {'hw0_f': (Hc * stride, Wc * stride), 'hw0_4c': (Hc, Wc), 'hw1_4c': (Hc, Wc), 'stage_4c': {'b_ids': b.to(dev), 'i_ids': i.to(dev), 'j_ids': j.to(dev)}}
    with torch.no_grad():
        o0, o1 = mod.to(dev)(ff0.to(dev), ff1.to(dev), fc0.to(dev), fc1.to(dev), data)
        u0 = tF.unfold(ff0, (W, W), stride=stride, padding=W // 2).reshape(B, C, W * W, -1).permute(0, 3, 2, 1)[b, i]
        u1 = tF.unfold(ff1, (W, W), stride=stride, padding=W // 2).reshape(B, C, W * W, -1).permute(0, 3, 2, 1)[b, j]
        if cat:
            m = mod.cpu()
            cw = m.down_proj(torch.cat([fc0[b, i], fc1[b, j]], 0))
            cf = m.merge_feat(torch.cat([torch.cat([u0, u1], 0), cw.unsqueeze(1).expand(-1, W * W, -1)], -1))
            u0, u1 = torch.chunk(cf, 2, dim=0)
    tol = 1e-3 if cat else 0.0                      # the gather is exact; the Linear layers run in cuBLAS (TF32 off) vs CPU
    assert (o0.cpu() - u0).abs().max() <= tol and (o1.cpu() - u1).abs().max() <= tol
