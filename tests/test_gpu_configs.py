"""GPU parity at the shapes of the other BASELINE.json configs (parity-test cases, not bench lines):
  cfg 3  CasMTR-2c outdoor: the extra cascade stage at 1/2 resolution (416x416 tokens, C=64, 2 heads, K=100)
  cfg 4  CasMTR-4c indoor 640x480: 60x80 / 120x160 grids, topks [32,16,16], rel_pos bias in CascadeQTAttB,
         threshold-only detection (no NMS), border 1
  cfg 5  size sweep: QTAttB at the 1/8 grids of 512 / 1024 / 1152 images (coarsest level 256 / 1024 / 1296 keys)
Each is compared with the CPU oracle (seconds at these sizes) through the module API."""
import pytest
import torch

import casmtr_b200
from casmtr_b200 import functional as F
from casmtr_b200 import synth
from oracle import cascade as ocas, qtatt as oqt
from oracle.compare import check_qtatt_levels

pytestmark = pytest.mark.gpu


def _idx(tp, B, h, w):
    c = oqt.cascade_window_idx(tp, h, w)
    return oqt.quad_to_raster(c.reshape(B, 1, -1, 1, 100).expand(B, 1, -1, 4, 100), h // 2, w // 2).reshape(B, h * w, 100).contiguous()


def test_cfg3_half_resolution_stage(dev):
    B, C, nh, h, w = 1, 64, 2, 416, 416
    d = synth.cascade_inputs(B, C, h, w, seed=303, max_shift=8)
    v = torch.randn(B, C, h, w, generator=torch.Generator().manual_seed(1))
    att = casmtr_b200.CascadeQTAttB(nh, 32, dilated=1)
    msg, up01 = att(d['feat0'].to(dev), d['feat1'].to(dev), v.to(dev), d['topk_pos01'].to(dev), None)
    _, up10 = att(d['feat1'].to(dev), d['feat0'].to(dev), v.to(dev), d['topk_pos10'].to(dev), None)
    idx01, idx10 = _idx(d['topk_pos01'], B, h, w), _idx(d['topk_pos10'], B, h, w)
    assert torch.equal(up01.cpu(), idx01) and torch.equal(up10.cpu(), idx10)
    ref_m, _ = oqt.cascade_qtatt_b(d['feat0'], d['feat1'], v, d['topk_pos01'], None, nh)
    assert (msg.cpu() - ref_m).abs().max() < 1e-3
    f0 = d['feat0'].flatten(2).transpose(1, 2).contiguous()
    f1 = d['feat1'].flatten(2).transpose(1, 2).contiguous()
    o = F.cascade_match_forward(f0.to(dev), f1.to(dev), up01, up10, need_conf=False, w0=w, w1=w)      # C = 64: two 32-channel slices
    ref = ocas.cascade_match(f0, f1, idx01, idx10)
    for t in ('01', '10'):
        assert torch.equal(o['next_idx' + t].cpu(), ref['next_idx' + t])
        assert (o['next_conf' + t].cpu() - ref['next_conf' + t]).abs().max() < 1e-5
    # last stage of the 2c model: NMS 5, pre-stage gates of both previous stages (cascade_matching.py:199-206)
    g = torch.Generator().manual_seed(2)
    pre = [(torch.rand(B, (h // 2) * (w // 2), generator=g), h // 2, w // 2), (torch.rand(B, (h // 4) * (w // 4), generator=g), h // 4, w // 4)]
    kw = dict(test_thr=0.2, border_rm=2, nms_window=5, pre_thrs=[0.2, 0.2], double_check=True)
    want = ocas.extract_matches(ref['next_conf01'], ref['next_idx01'], ref['next_idx10'], (h, w), (h, w), (2 * h, 2 * w), pre_confs=pre, **kw)
    got = F.match_extract(o['next_conf01'], o['next_idx01'], o['next_idx10'], (h, w), (h, w), (2 * h, 2 * w),
                          pre_confs=[(p.to(dev), a, b) for p, a, b in pre], **kw)
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(got[k].cpu(), want[k]), k
    assert len(want['b_ids']) > 100


def test_cfg4_indoor_stage(dev):
    B, C, nh, h, w = 2, 128, 4, 120, 160
    d = synth.cascade_inputs(B, C, h, w, seed=404, max_shift=6)
    g = torch.Generator().manual_seed(3)
    v = torch.randn(B, C, h, w, generator=g)
    rp = 0.5 * torch.randn(B, nh, h * w, 100, generator=g)          # relative position bias (indoor config: relative_pe True)
    att = casmtr_b200.CascadeQTAttB(nh, 32, dilated=1)
    msg, up01 = att(d['feat0'].to(dev), d['feat1'].to(dev), v.to(dev), d['topk_pos01'].to(dev), rp.to(dev))
    ref_m, ref_i = oqt.cascade_qtatt_b(d['feat0'], d['feat1'], v, d['topk_pos01'], rp, nh)
    assert torch.equal(up01.cpu(), ref_i) and (msg.cpu() - ref_m).abs().max() < 1e-3
    _, up10 = att(d['feat1'].to(dev), d['feat0'].to(dev), v.to(dev), d['topk_pos10'].to(dev), None)
    f0 = d['feat0'].flatten(2).transpose(1, 2).contiguous()
    f1 = d['feat1'].flatten(2).transpose(1, 2).contiguous()
    cfg = {'thr': 0.0, 'test_thr': 0.1, 'pre_thr': [0.2, 0.1], 'border_rm': 1, 'double_check': True,
           'train_pad_num_gt_min': 4096, 'match_type': 'softmax', 'dsmax_temperature': 1.0}
    cas = {'propagation': 'window', 'dilated': 1, 'detector_mode': None, 'grid_size': 4,
           'post_config': {'method': None, 'window_size': 5, 'topk': None, 'rt': None, 'rd': None}}
    data = {'hw0_i': (4 * h, 4 * w), 'hw1_i': (4 * h, 4 * w), 'hw0_4c': (h, w), 'hw1_4c': (h, w), 'hw0_8c': (h // 2, w // 2),
            'hw1_8c': (h // 2, w // 2), 'bs': B, 'stage_8c': {'next_conf_c01': d['pre_conf01'].to(dev)}}
    casmtr_b200.CascadeMatching(cfg, cas).eval()(f0.to(dev), f1.to(dev), up01, up10, data, level='4c', pre_level='8c')
    ref = ocas.cascade_match(f0, f1, up01.cpu(), up10.cpu())
    want = ocas.extract_matches(ref['next_conf01'], ref['next_idx01'], ref['next_idx10'], (h, w), (h, w), (4 * h, 4 * w), test_thr=0.1,
                                border_rm=1, nms_window=None, pre_confs=[(d['pre_conf01'], h // 2, w // 2)], pre_thrs=[0.2], double_check=True)
    st = data['stage_4c']
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(st[k].cpu(), want[k]), k
    assert (st['conf_matrix'].cpu() - ref['conf01']).abs().max() < 1e-5 and len(want['b_ids']) > 1000


# coarsest-level keys: 64 -> 256 (8 values per lane), 80 -> 300 (16), 128 -> 1024 (32, one row at a time), 144 -> 1296 (42),
# 152 -> 1444 (beyond the register path: looped soft-max / top-k)
@pytest.mark.parametrize('grid,topks', [(64, [32, 16, 8]), (128, [32, 16, 8]), (144, [32, 16, 8]), (152, [32, 16, 8]), (80, [32, 16, 16])])
def test_cfg5_size_sweep_qtatt(dev, grid, topks):
    nh = 8
    h, w = (60, 80) if grid == 80 else (grid, grid)
    qs, ks, vs, wt = synth.qtatt_inputs(1, nh * 32, h, w, 3, seed=500 + grid)
    ref, aux = oqt.qtatt_b(qs, ks, vs, wt, topks, nh, return_aux=True)
    out, idx, sc = F.qtatt_forward([t.to(dev) for t in qs], [t.to(dev) for t in ks], [t.to(dev) for t in vs], topks, nh,
                                   weight=wt.to(dev), attn_type='B', return_topk=True)
    check_qtatt_levels(out, idx, sc, ref, aux, h, w, 3, f'QTAttB {h}x{w}')
