"""GPU parity at BASELINE.json's full size (configs[1]: CasMTR-4c outdoor, 832x832, batch 1), driven through the same
pipeline object bench.py times.  One call of each kind is compared with the CPU oracle directly (it finishes in seconds
at this size); every call is checked through size-independent properties of the domain:
  * CascadeQTAttB: upsampled_idx is exactly the 10x10 window arithmetic of topk_pos; message rows are convex combinations
    of value rows (inside their min/max);
  * QTAttB: linear in the values (top-k selection does not depend on V): f(q, k, 2v) == 2 f(q, k, v);
  * CascadeMatching: confidence rows sum to 1, next_conf is the row maximum, next_idx is the arg-max candidate;
  * match list: row-major (b, i) order, mutual nearest neighbours, above threshold, a strict 5x5 local maximum (first in
    scan order on ties), inside the border; keypoints are the grid coordinates times the scale."""
import pytest
import torch

from casmtr_b200 import pipeline
from oracle import cascade as ocas, fine as ofine, qtatt as oqt
from oracle.compare import check_qtatt_levels

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def run(dev):
    wl = pipeline.Workload(832, 832, pairs=1, qt_calls=2, cas_calls=4)
    host = pipeline.make_host_inputs(wl, seed=4321)
    hp = pipeline.HotPath(wl).to(dev)
    hp.load_level_weights(host)
    dev_in = pipeline.tree_map(lambda t: t.to(dev), host)
    keep = {}
    out = hp(dev_in, keep=keep)
    torch.cuda.synchronize()
    return wl, host, hp, dev_in, keep, out


def test_qtatt_b_full_size_vs_oracle(run, dev):
    wl, host, hp, dev_in, keep, _ = run
    from casmtr_b200 import functional as F
    c = host['qt'][0]
    ref, aux = oqt.qtatt_b(c['q'], c['k'], c['v'], c['weight'], wl.topks, wl.nh8, return_aux=True)
    d = dev_in['qt'][0]
    out, idx, sc = F.qtatt_forward(d['q'], d['k'], d['v'], wl.topks, wl.nh8, weight=d['weight'], attn_type='B', return_topk=True)
    err, ties = check_qtatt_levels(out, idx, sc, ref, aux, wl.h8, wl.w8, 3, 'QTAttB 832^2')
    assert torch.equal(out, keep['qt_msg'][0])                  # the pipeline call is the same computation, bit for bit


def test_qtatt_b_linear_in_values(run, dev):
    wl, host, hp, dev_in, keep, _ = run
    d = dev_in['qt'][1]
    twice = hp.qt[1](d['q'], d['k'], [2.0 * v for v in d['v']])
    assert (twice - 2.0 * keep['qt_msg'][1]).abs().max() < 1e-4


def test_cascade_qtatt_full_size(run, dev):
    wl, host, hp, dev_in, keep, _ = run
    for i in range(wl.cas_calls):
        c = host['cas'][i]
        want = oqt.cascade_window_idx(c['topk_pos'], wl.h4, wl.w4)                                  # [B, Np, 100]
        want = oqt.quad_to_raster(want.reshape(wl.B, 1, -1, 1, 100).expand(wl.B, 1, -1, 4, 100), wl.h4 // 2, wl.w4 // 2)
        assert torch.equal(keep['cas_idx'][i].cpu(), want.reshape(wl.B, wl.h4 * wl.w4, 100))          # integer work: bit-exact
        v = c['v'].flatten(2).transpose(1, 2)                                                        # [B, L, C]
        m = keep['cas_msg'][i].cpu()
        assert (m <= v.amax(dim=1, keepdim=True) + 1e-4).all() and (m >= v.amin(dim=1, keepdim=True) - 1e-4).all()
    c = host['cas'][0]
    ref_m, _ = oqt.cascade_qtatt_b(c['q'], c['k'], c['v'], c['topk_pos'], None, wl.nh4)
    assert (keep['cas_msg'][0].cpu() - ref_m).abs().max() < 1e-3


def test_cascade_matching_full_size(run, dev):
    wl, host, hp, dev_in, keep, out = run
    st = keep['data']['stage_4c']
    idx01, idx10 = keep['cas_idx'][2].cpu(), keep['cas_idx'][3].cpu()
    conf = st['conf_matrix'].cpu()
    assert (conf.sum(-1) - 1).abs().max() < 1e-5
    nconf, arg = conf.max(dim=-1)
    assert torch.equal(nconf, st['next_conf_c01'].cpu())
    picked = torch.gather(idx01, 2, arg.unsqueeze(-1)).squeeze(-1)
    assert torch.equal(picked, st['next_idx_c01'].cpu())
    # the oracle on the same inputs (a few seconds on the host)
    m = host['match']
    o = ocas.cascade_match(m['feat0'], m['feat1'], idx01, idx10, None, None, 1.0)
    assert torch.equal(o['next_idx01'], st['next_idx_c01'].cpu()) and torch.equal(o['next_idx10'], st['next_idx_c10'].cpu())
    assert (o['next_conf01'] - st['next_conf_c01'].cpu()).abs().max() < 1e-5
    r = ocas.extract_matches(o['next_conf01'], o['next_idx01'], o['next_idx10'], (wl.h4, wl.w4), (wl.h4, wl.w4), (wl.H, wl.W),
                             test_thr=0.2, border_rm=2, nms_window=5, pre_confs=[(m['pre_conf'], wl.h8, wl.w8)],
                             pre_thrs=[0.2], double_check=True)
    for k in ('b_ids', 'i_ids', 'j_ids'):
        assert torch.equal(st[k].cpu(), r[k]), k
    assert torch.equal(st['mkpts0_c'].cpu(), r['mkpts0_c'].float()) and torch.equal(st['mkpts1_c'].cpu(), r['mkpts1_c'].float())


def test_match_list_properties(run, dev):
    wl, host, hp, dev_in, keep, out = run
    st = keep['data']['stage_4c']
    b, i, j = st['b_ids'].cpu(), st['i_ids'].cpu(), st['j_ids'].cpu()
    M = b.numel()
    assert M > 100
    order = b * (wl.h4 * wl.w4) + i
    assert (order[1:] > order[:-1]).all()                                       # torch.where order, no duplicates
    nconf = st['next_conf_c01'].cpu()
    assert (nconf[b, i] > 0.2).all() and torch.equal(nconf[b, i], st['mconf'].cpu())
    assert torch.equal(st['next_idx_c10'].cpu()[b, j], i)                       # mutual nearest neighbours
    y, x = i // wl.w4, i % wl.w4
    assert (y >= 2).all() and (x >= 2).all() and (y < wl.h4 - 2).all() and (x < wl.w4 - 2).all()
    grid = torch.nn.functional.pad(nconf.reshape(wl.B, wl.h4, wl.w4), (2, 2, 2, 2), value=float('-inf'))
    win = grid.unfold(1, 5, 1).unfold(2, 5, 1).reshape(wl.B, wl.h4, wl.w4, 25)   # 5x5 neighbourhoods, row-major
    centre = win[..., 12]
    assert (centre[b, y, x] >= win[b, y, x].amax(-1)).all()                     # local maxima
    assert (centre[b, y, x].unsqueeze(-1) > win[b, y, x][:, :12]).all()         # strictly above everything scanned earlier
    scale = wl.H / wl.h4
    assert torch.equal(st['mkpts0_c'].cpu(), torch.stack([x, y], 1).float() * scale)
    # fine stage ran on exactly these matches
    assert out['mkpts1'].shape == (M, 2) and out['expec_f'].shape == (M, 3)
    e, k = ofine.fine_match(host['fine']['feat_f0'][:M], host['fine']['feat_f1'][:M], st['mkpts1_c'].cpu(), wl.H / wl.hf)
    assert (out['expec_f'].cpu() - e).abs().max() < 1e-5 and (out['mkpts1'].cpu() - k).abs().max() < 1e-3


@pytest.mark.parametrize('two_streams', [False, True])
def test_whole_step_graph_equals_eager(run, dev, two_streams):
    """The sync-free CUDA-graph step (match count consumed on the device by the fine stage) returns exactly the eager
    result once trimmed; replaying it is idempotent."""
    wl, host, hp, dev_in, keep, out = run
    gr = pipeline.GraphRunner(hp, dev_in, two_streams=two_streams, whole_step=True)
    for _ in range(2):
        got = pipeline.trim_result(gr.step())
        torch.cuda.synchronize()
        assert got['mconf'].shape[0] == out['mconf'].shape[0] > 100
        for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0', 'mkpts1', 'expec_f'):
            assert torch.equal(got[k], out[k]), k
    # the packed all-gather block built from the device-side count equals the one built from the trimmed list
    from casmtr_b200 import functional as F
    blk_dev = F.pack_matches(gr.step(), 7, wl.fine_cap)
    blk = F.pack_matches({k: out[k] for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0', 'mkpts1')}, 7, wl.fine_cap)
    M = out['mconf'].shape[0]
    assert torch.equal(blk_dev[:M + 1], blk[:M + 1])
