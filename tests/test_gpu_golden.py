"""GPU parity against the golden vectors produced by the REFERENCE itself (tests/golden/*.npz, generated in
the build container by tests/golden/make_golden.py from the unmodified reference Python modules).  The CUDA path is
driven through the module API -> C ABI.  Integer outputs bit-exact, fp32 within 1e-3 abs (observed ~1e-6)."""
import pytest
import torch

import casmtr_b200
from casmtr_b200 import functional as F

from golden_util import load, same_sets

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _pyr(g, p, dev):
    return [g[f'{p}{i}'].to(dev) for i in range(3)]


def test_qtatt_b_vs_reference(dev):
    g = load('qtatt')
    topks, nh = g['topks'].tolist(), int(g['nhead'])
    out, idx, _ = F.qtatt_forward(_pyr(g, 'q', dev), _pyr(g, 'k', dev), _pyr(g, 'v', dev), topks, nh,
                                  weight=g['weight'].to(dev), attn_type='B', return_topk=True)
    assert (out.cpu() - g['out_b']).abs().max() < TOL
    assert same_sets(idx[0].cpu(), g['b_idx0'], 2) and same_sets(idx[1].cpu(), g['b_idx1'], 2)
    assert torch.equal(idx[0].cpu(), g['b_idx0'])          # no ties in the fixture: the descending order matches too


def test_qtatt_a_vs_reference(dev):
    g = load('qtatt')
    topks, nh = g['topks'].tolist(), int(g['nhead'])
    out, idx, sc = F.qtatt_forward(_pyr(g, 'q', dev), _pyr(g, 'k', dev), _pyr(g, 'v', dev), topks, nh,
                                   attn_type='A', return_topk=True)
    assert (out.cpu() - g['out_a']).abs().max() < TOL
    assert same_sets(idx[0].cpu(), g['a_idx0'], 2) and same_sets(idx[1].cpu(), g['a_idx1'], 2)
    assert (torch.sort(sc[1].cpu(), dim=2)[0] - torch.sort(g['a_score1'], dim=2)[0]).abs().max() < 1e-5


def test_cascade_qtatt_vs_reference(dev):
    g = load('cascade_qtatt')
    nh = int(g['nhead'])
    m = casmtr_b200.CascadeQTAttB(nh, g['query'].shape[1] // nh, dilated=1)
    msg, up = m(g['query'].to(dev), g['key'].to(dev), g['value'].to(dev), g['topk_pos'].to(dev), None)
    assert torch.equal(up.cpu(), g['upsampled_idx'])
    assert (msg.cpu() - g['message']).abs().max() < TOL
    msg2, _ = m(g['query'].to(dev), g['key'].to(dev), g['value'].to(dev), g['topk_pos'].to(dev), g['rel_pos'].to(dev))
    assert (msg2.cpu() - g['message_rel']).abs().max() < TOL


def _module(nms, thr):
    cfg = {'thr': 0.0101, 'test_thr': thr, 'pre_thr': [0.2], 'border_rm': 2, 'double_check': True,
           'train_pad_num_gt_min': 4096, 'match_type': 'softmax', 'dsmax_temperature': 1.0}
    cas = {'propagation': 'window', 'dilated': 1, 'detector_mode': None, 'grid_size': 4,
           'post_config': {'method': 'maxpool_nms' if nms else None, 'window_size': 5, 'topk': None, 'rt': None, 'rd': None}}
    return casmtr_b200.CascadeMatching(cfg, cas).eval()


@pytest.mark.parametrize('tag,nms,thr,pad', [('nms', True, 0.2, False), ('thr', False, 0.2, False), ('empty', True, 2.0, False),
                                             ('pad', True, 0.2, True)])
def test_cascade_matching_vs_reference(dev, tag, nms, thr, pad):
    g = load('cascade_match')
    h, w = g['hw'].tolist()
    B = g['feat0'].shape[0]
    data = {'hw0_i': (h * 4, w * 4), 'hw1_i': (h * 4, w * 4), 'hw0_4c': (h, w), 'hw1_4c': (h, w),
            'hw0_8c': (h // 2, w // 2), 'hw1_8c': (h // 2, w // 2), 'bs': B,
            'stage_8c': {'next_conf_c01': g['pre_conf'].to(dev)}}
    m0 = m1 = None
    if pad:
        data['mask_4c0'], data['mask_4c1'] = g['pad_mask0'].bool().to(dev), g['pad_mask1'].bool().to(dev)
        m0, m1 = data['mask_4c0'].flatten(1), data['mask_4c1'].flatten(1)
        data['scale0'], data['scale1'] = g['scale0'].to(dev), g['scale1'].to(dev)
    _module(nms, thr)(g['feat0'].to(dev), g['feat1'].to(dev), g['idx01'].to(dev), g['idx10'].to(dev), data,
                      mask_c0=m0, mask_c1=m1, level='4c', pre_level='8c')
    st = data['stage_4c']
    if tag == 'nms':
        assert torch.equal(st['next_idx_c01'].cpu(), g['next_idx01']) and torch.equal(st['next_idx_c10'].cpu(), g['next_idx10'])
        assert (st['conf_matrix'].cpu() - g['conf01']).abs().max() < 1e-5
        assert (st['next_conf_c10'].cpu() - g['next_conf10']).abs().max() < 1e-5
    if tag == 'pad':
        assert torch.equal(st['next_idx_c01'].cpu(), g['pad_next_idx01']) and torch.equal(st['next_idx_c10'].cpu(), g['pad_next_idx10'])
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(st[k].cpu(), g[f'{tag}_{k}']), (tag, k)
    assert (st['mconf'].cpu() - g[f'{tag}_mconf']).abs().max() < 1e-5
    for k in ('mkpts0_c', 'mkpts1_c'):
        assert torch.equal(st[k].cpu(), g[f'{tag}_{k}']), (tag, k)


def test_fine_matching_vs_reference(dev):
    g = load('fine_match')
    M = g['feat_f0'].shape[0]
    for tag, scaled in (('plain', False), ('scaled', True)):
        data = {'hw0_i': (128, 128), 'hw0_f': (64, 64),
                'stage_4c': {'mkpts0_c': torch.zeros(M, 2, device=dev), 'mkpts1_c': g['mkpts1_c'].to(dev),
                             'mconf': torch.zeros(M, device=dev), 'b_ids': g['b_ids'].to(dev)}}
        if scaled:
            data['scale0'] = data['scale1'] = g['scale1'].to(dev)
        casmtr_b200.CascadeFineMatching('4c').eval()(g['feat_f0'].to(dev), g['feat_f1'].to(dev), data)
        assert (data['expec_f'].cpu() - g['plain_expec_f']).abs().max() < 1e-5
        assert (data['mkpts1_f'].cpu() - g[f'{tag}_mkpts1_f']).abs().max() < 1e-4


# ---- SURVEY section 8f "next" rows against the reference's own modules (tests/golden/widen_*.npz)
def _load_sd(mod, g, prefix):
    sd = {}
    for k in mod.state_dict():
        sd[k] = g[prefix + k.replace('.', '_')]
    mod.load_state_dict(sd)                                    # the reference's state-dict keys, unchanged
    return mod


def test_coarse_matching_vs_reference(dev):
    g = load('widen_coarse_match')
    cfg = {'thr': 0.2, 'border_rm': 2, 'train_coarse_percent': 0.3, 'train_pad_num_gt_min': 200, 'match_type': 'dual_softmax',
           'dsmax_temperature': float(g['temperature'])}
    data = {'hw0_i': (96, 128), 'hw1_i': (96, 128), 'hw0_8c': (12, 16), 'hw1_8c': (12, 16), 'bs': 2}
    casmtr_b200.CoarseMatching(cfg).eval()(g['feat0'].to(dev), g['feat1'].to(dev), data)
    st = data['stage_8c']
    assert (st['next_conf_c01'].cpu() - g['next_conf01']).abs().max() < TOL and (st['next_conf_c10'].cpu() - g['next_conf10']).abs().max() < TOL
    # arg-max is specified where the best two logits are apart (3xTF32 products carry ~1e-6 relative error)
    C = g['feat0'].shape[-1]
    sim = torch.einsum('nlc,nsc->nls', g['feat0'].double(), g['feat1'].double()) / C / float(g['temperature'])
    t01, t10 = sim.topk(2, dim=2)[0], sim.topk(2, dim=1)[0]
    clear01, clear10 = (t01[..., 0] - t01[..., 1]) > 1e-4, (t10[:, 0] - t10[:, 1]) > 1e-4
    assert clear01.float().mean() > 0.99
    assert torch.equal(st['next_idx_c01'].cpu()[clear01], g['next_idx01'][clear01])
    assert torch.equal(st['next_idx_c10'].cpu()[clear10], g['next_idx10'][clear10])


@pytest.mark.parametrize('cat', [False, True])
def test_fine_preprocess_vs_reference(dev, cat):
    g = load('widen_fine_preprocess')
    hc, wc = g['hw_c'].tolist()
    s = int(g['stride'])
    mod = casmtr_b200.CascadeFinePreprocess({'fine_concat_coarse_feat': cat, 'fine_window_size': 5}, {'d_model': 32}, {'d_model': 64}, '4c').eval()
    if cat:
        _load_sd(mod, g, 'w_')
    data = {'hw0_f': (hc * s, wc * s), 'hw0_4c': (hc, wc), 'hw1_4c': (hc, wc),
            'stage_4c': {'b_ids': g['b_ids'].to(dev), 'i_ids': g['i_ids'].to(dev), 'j_ids': g['j_ids'].to(dev)}}
    with torch.no_grad():
        o0, o1 = mod.to(dev)(g['feat_f0'].to(dev), g['feat_f1'].to(dev), g['feat_c0'].to(dev), g['feat_c1'].to(dev), data)
    tag = 'cat' if cat else 'plain'
    if cat:
        assert (o0.cpu() - g['cat_out0']).abs().max() < TOL and (o1.cpu() - g['cat_out1']).abs().max() < TOL
    else:
        assert torch.equal(o0.cpu(), g[f'{tag}_out0']) and torch.equal(o1.cpu(), g[f'{tag}_out1'])       # a gather: bit-exact


def test_attention_layers_vs_reference(dev):
    g = load('widen_attention_layers')
    H, W = g['hw'].tolist()
    nh, topks = int(g['nhead']), g['topks'].tolist()
    layer = _load_sd(casmtr_b200.QuadtreeAttention(64, nh, topks, scale=3, attn_type='B'), g, 'B_').to(dev).eval()
    with torch.no_grad():
        out = layer(g['x'].to(dev), g['target'].to(dev), H, W)
    bad = ((out.cpu() - g['out_B']).abs().amax(dim=2) > 1e-4).float().mean().item()   # GEMM rounding may flip an exact near-tie of the top-k
    assert bad < 5e-3, bad
    cl = _load_sd(casmtr_b200.CascadeQuadtreeAttention(64, nh, dilated=1), g, 'C_').to(dev).eval()
    with torch.no_grad():
        co, up = cl(g['cas_x'].to(dev), g['cas_target'].to(dev), 16, 16, idx=g['cas_topk_pos'].to(dev))
    assert torch.equal(up.cpu(), g['cas_upsampled_idx'])
    assert (co.cpu() - g['cas_out']).abs().max() < TOL
