"""GPU parity at BASELINE.json's full sizes, driven through the same pipeline object bench.py times
(casmtr_b200/pipeline.py: the two directions of every layer stacked on the batch dimension, token-major entry), against the
whole-path oracle (oracle/hotpath.py) on the same synthetic inputs:
  configs[1]  CasMTR-4c outdoor 832x832                      -> `run` fixture (one coarse layer instead of six: same kernels)
  configs[2]  CasMTR-2c outdoor 832x832 (1/8 -> 1/4 -> 1/2)  -> test_2c_full_size
  configs[3]  CasMTR-4c indoor 640x480 with relative PE      -> test_indoor_full_size
plus size-independent properties of the domain:
  * CascadeQTAttB: upsampled_idx is exactly the 10x10 window arithmetic of the matches; message rows are convex
    combinations of value rows (inside their min/max);
  * QTAttB: linear in the values (top-k selection does not depend on V): f(q, k, 2v) == 2 f(q, k, v);
  * CascadeMatching: confidence rows sum to 1, next_conf is the row maximum, next_idx is the arg-max candidate;
  * match list: row-major (b, i) order, mutual nearest neighbours, above threshold, a strict 5x5 local maximum (first in
    scan order on ties), inside the border; keypoints are the grid coordinates times the scale."""
import pytest
import torch

from casmtr_b200 import functional as F
from casmtr_b200 import pipeline
from oracle import fine as ofine, hotpath, qtatt as oqt
from oracle.compare import check_qtatt_levels

pytestmark = pytest.mark.gpu


def _run(dev, **kw):
    wl = pipeline.Workload(**kw)
    host = pipeline.make_host_inputs(wl, seed=4321)
    hp = pipeline.HotPath(wl).to(dev)
    hp.load_level_weights(host)
    dev_in = pipeline.tree_map(lambda t: t.to(dev), host)
    keep = {}
    out = hp(dev_in, keep=keep)
    torch.cuda.synchronize()
    okeep = {}
    ref = hotpath.run_step(wl, host, keep=okeep)
    return wl, host, hp, dev_in, keep, out, okeep, ref


@pytest.fixture(scope='module')
def run(dev):
    return _run(dev, height=832, width=832, pairs=1, config='4c', qt_layers=2)


def _check_stage(wl, si, keep, okeep, tol_msg=1e-3):
    """Cascade attention + matching + extraction of stage si against the oracle: integer work bit-exact."""
    s = wl.stages[si]
    lvl = s['level']
    first = sum(x['cross'] for x in wl.stages[:si])
    for li in range(s['cross']):
        assert torch.equal(keep['cas_idx'][first + li].cpu(), okeep['cas_idx'][first + li]), f'{lvl} upsampled_idx layer {li}'
        err = (keep['cas_msg'][first + li].cpu() - okeep['cas_msg'][first + li]).abs().max().item()
        assert err < tol_msg, f'{lvl} cascade message layer {li}: {err}'
    st, ost = keep['data'][f'stage_{lvl}'], okeep['stage'][lvl]
    assert torch.equal(st['next_idx_c01'].cpu(), ost['next_idx01']) and torch.equal(st['next_idx_c10'].cpu(), ost['next_idx10'])
    assert (st['next_conf_c01'].cpu() - ost['next_conf01']).abs().max() < 1e-5
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(st[k].cpu(), ost[k]), f'{lvl} {k}'
    assert torch.equal(st['mkpts0_c'].cpu(), ost['mkpts0_c'].float()) and torch.equal(st['mkpts1_c'].cpu(), ost['mkpts1_c'].float())
    return st


def _check_final(out, ref):
    M = ref['mconf'].shape[0]
    assert out['mconf'].shape[0] == M
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(out[k].cpu(), ref[k]), k
    assert (out['expec_f'].cpu() - ref['expec_f']).abs().max() < 1e-5 and (out['mkpts1'].cpu() - ref['mkpts1']).abs().max() < 1e-3


def test_qtatt_b_full_size_vs_oracle(run, dev):
    wl, host, hp, dev_in, keep, _, okeep, _ = run
    c, d = host['qt'][0], dev_in['qt'][0]
    pyr = [hotpath._pyramid(hotpath._nchw(c[n], wl.h8, wl.w8)) for n in ('q', 'k', 'v')]
    ref, aux = oqt.qtatt_b(pyr[0], pyr[1], pyr[2], c['weight'], wl.topks, wl.nh8, return_aux=True)
    out, idx, sc = F.qtatt_tokens_forward(d['q'], d['k'], d['v'], (wl.h8, wl.w8), (wl.h8, wl.w8), wl.topks, wl.nh8,
                                          weight=hp.qt[0].weight, return_topk=True)
    # 2 x 8 x 14196 rows of random logits: an fp32 near-tie at a k-th place (relative gap < 1e-5) is expected once in a few
    # thousand rows for ANY summation order; those rows are checked to be genuine ties, everything else is exact
    err, ties = check_qtatt_levels(out, idx, sc, ref, aux, wl.h8, wl.w8, 3, 'QTAttB 832^2', max_tie_frac=2e-3)
    assert torch.equal(out, keep['qt_msg'][0])                  # the pipeline call is the same computation, bit for bit
    diff = (keep['qt_msg'][1].cpu() - okeep['qt_msg'][1]).abs().amax(-1)      # [2P, L, nh]: the second layer, through the pipeline
    assert (diff > 1e-3).float().mean() < 2e-3                                 # only descendants of near-tie rows may differ


def test_qtatt_b_linear_in_values(run, dev):
    wl, host, hp, dev_in, keep, *_ = run
    d = dev_in['qt'][1]
    twice = hp.run_qt(1, {'q': d['q'], 'k': d['k'], 'v': 2.0 * d['v']})
    assert (twice - 2.0 * keep['qt_msg'][1]).abs().max() < 1e-4


def test_cascade_qtatt_full_size(run, dev):
    wl, host, hp, dev_in, keep, _, okeep, _ = run
    s = wl.stages[0]
    for i in range(s['cross']):
        pos = hotpath.window_positions(host['hand']['next_idx'], s['h'] // 2, s['w'] // 2)
        want = oqt.cascade_window_idx(pos, s['h'], s['w'])                                        # [2P, Np, 100]
        want = oqt.quad_to_raster(want.reshape(wl.B, 1, -1, 1, 100).expand(wl.B, 1, -1, 4, 100), s['h'] // 2, s['w'] // 2)
        assert torch.equal(keep['cas_idx'][i].cpu(), want.reshape(wl.B, s['h'] * s['w'], 100))    # integer work: bit-exact
        v = host['stages'][0]['layers'][i]['v']                                                  # token-major [2P, L, C]
        m = keep['cas_msg'][i].cpu()
        assert (m <= v.amax(dim=1, keepdim=True) + 1e-4).all() and (m >= v.amin(dim=1, keepdim=True) - 1e-4).all()


def test_cascade_stage_and_match_list_vs_oracle(run, dev):
    wl, host, hp, dev_in, keep, out, okeep, ref = run
    st = _check_stage(wl, 0, keep, okeep)
    conf = st['conf_matrix'].cpu()
    assert (conf.sum(-1) - 1).abs().max() < 1e-5
    nconf, arg = conf.max(dim=-1)
    assert torch.equal(nconf, st['next_conf_c01'].cpu())
    picked = torch.gather(keep['cas_idx'][-1][:wl.P].cpu(), 2, arg.unsqueeze(-1)).squeeze(-1)
    assert torch.equal(picked, st['next_idx_c01'].cpu())
    _check_final(out, ref)


def test_match_list_properties(run, dev):
    wl, host, hp, dev_in, keep, out, *_ = run
    s = wl.stages[0]
    h4, w4 = s['h'], s['w']
    st = keep['data']['stage_4c']
    b, i, j = st['b_ids'].cpu(), st['i_ids'].cpu(), st['j_ids'].cpu()
    M = b.numel()
    assert M > 100
    order = b * (h4 * w4) + i
    assert (order[1:] > order[:-1]).all()                                       # torch.where order, no duplicates
    nconf = st['next_conf_c01'].cpu()
    assert (nconf[b, i] > 0.2).all() and torch.equal(nconf[b, i], st['mconf'].cpu())
    assert torch.equal(st['next_idx_c10'].cpu()[b, j], i)                       # mutual nearest neighbours
    y, x = i // w4, i % w4
    assert (y >= 2).all() and (x >= 2).all() and (y < h4 - 2).all() and (x < w4 - 2).all()
    grid = torch.nn.functional.pad(nconf.reshape(wl.P, h4, w4), (2, 2, 2, 2), value=float('-inf'))
    win = grid.unfold(1, 5, 1).unfold(2, 5, 1).reshape(wl.P, h4, w4, 25)        # 5x5 neighbourhoods, row-major
    centre = win[..., 12]
    assert (centre[b, y, x] >= win[b, y, x].amax(-1)).all()                     # local maxima
    assert (centre[b, y, x].unsqueeze(-1) > win[b, y, x][:, :12]).all()         # strictly above everything scanned earlier
    scale = wl.H / h4
    assert torch.equal(st['mkpts0_c'].cpu(), torch.stack([x, y], 1).float() * scale)
    # fine stage ran on exactly these matches
    assert out['mkpts1'].shape == (M, 2) and out['expec_f'].shape == (M, 3)
    e, k = ofine.fine_match(host['fine']['feat_f0'][:M], host['fine']['feat_f1'][:M], st['mkpts1_c'].cpu(), wl.H / wl.hf)
    assert (out['expec_f'].cpu() - e).abs().max() < 1e-5 and (out['mkpts1'].cpu() - k).abs().max() < 1e-3


def test_whole_step_graph_equals_eager(run, dev):
    """The sync-free CUDA-graph step (match count consumed on the device by the fine stage) returns exactly the eager
    result once trimmed; replaying it is idempotent."""
    wl, host, hp, dev_in, keep, out, *_ = run
    gr = pipeline.GraphRunner(hp, dev_in)
    for _ in range(2):
        got = pipeline.trim_result(gr.step())
        torch.cuda.synchronize()
        assert got['mconf'].shape[0] == out['mconf'].shape[0] > 100
        for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0', 'mkpts1', 'expec_f'):
            assert torch.equal(got[k], out[k]), k
    # the packed all-gather block built from the device-side count equals the one built from the trimmed list
    blk_dev = F.pack_matches(gr.step(), 7, wl.fine_cap)
    blk = F.pack_matches({k: out[k] for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0', 'mkpts1')}, 7, wl.fine_cap)
    M = out['mconf'].shape[0]
    assert torch.equal(blk_dev[:M + 1], blk[:M + 1])


def test_nchw_entry_equals_token_entry(run, dev):
    """The reference-API entry (NCHW pyramid lists through QTAttB / CascadeQTAttB modules) and the token-major entry are the
    same computation: identical match lists, messages equal to fp32 rounding of the pooled pyramid."""
    wl, host, hp, dev_in, keep, out, *_ = run
    wl2 = pipeline.Workload(832, 832, pairs=1, config='4c', qt_layers=1, entry='nchw')
    host2 = pipeline.make_host_inputs(wl2, seed=4321)
    hp2 = pipeline.HotPath(wl2).to(dev)
    hp2.load_level_weights(host2)
    keep2 = {}
    out2 = hp2(pipeline.tree_map(lambda t: t.to(dev), host2), keep=keep2)
    assert (keep2['qt_msg'][0].flatten(2) - keep['qt_msg'][0].flatten(2)).abs().max() < 1e-5
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(out2[k], out[k]), k


def test_2c_full_size(dev):
    """BASELINE.json configs[2]: CasMTR-2c at 832^2 -- a second cascade stage at 1/2 resolution (173 056 tokens, C = 64, 2 heads)
    whose windows come from the 1/4 stage's own matches, two previous-stage gates, NMS only at the last stage
    (reference src/model/cascade_model_stage4.py:178-195)."""
    wl, host, hp, dev_in, keep, out, okeep, ref = _run(dev, height=832, width=832, pairs=1, config='2c', qt_layers=1)
    assert [s['level'] for s in wl.stages] == ['4c', '2c'] and wl.stages[1]['h'] == 416
    _check_stage(wl, 0, keep, okeep)
    st = _check_stage(wl, 1, keep, okeep)
    assert st['mconf'].shape[0] > 100
    # the 1/2 stage attended to the windows of the 1/4 stage's matches
    prev = keep['data']['stage_4c']
    assert torch.equal(keep['cas_nidx'][2], torch.cat([prev['next_idx_c01'], prev['next_idx_c10']], 0))
    _check_final(out, ref)
    gr = pipeline.GraphRunner(hp, dev_in)
    got = pipeline.trim_result(gr.step())
    for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts1'):
        assert torch.equal(got[k], out[k]), k


def test_indoor_full_size(dev):
    """BASELINE.json configs[3] per GPU: CasMTR-4c indoor 640x480, 4 pairs, topks [32,16,16], relative position bias computed
    inside the cascade attention kernels, threshold-only extraction with border 1."""
    wl, host, hp, dev_in, keep, out, okeep, ref = _run(dev, height=480, width=640, pairs=4, config='indoor', qt_layers=1)
    assert wl.topks == [32, 16, 16] and wl.B == 8
    _check_stage(wl, 0, keep, okeep)
    _check_final(out, ref)
