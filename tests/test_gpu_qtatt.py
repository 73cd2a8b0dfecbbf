"""GPU parity: fused QTAttA / QTAttB / CascadeQTAttB (module API -> C ABI) vs the CPU oracle.
Indices bit-exact (compared as sets per row: torch.topk tie/sort order is unspecified, the synthetic
inputs have no near-ties), messages within 1e-3 abs (BASELINE.json tolerance; observed ~1e-6)."""
import pytest
import torch

import casmtr_b200
from casmtr_b200 import functional as F
from casmtr_b200 import synth
from oracle import qtatt

from oracle.compare import check_qtatt_levels

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _cuda(lst, dev):
    return [t.to(dev) for t in lst]


def _check_levels(out, tk_idx, tk_sc, ref, aux, h, w, lv, what):
    check_qtatt_levels(out, tk_idx, tk_sc, ref, aux, h, w, lv, what, tol=TOL)


@pytest.mark.parametrize('B,nh,h,w,topks', [
    (1, 8, 32, 32, [32, 16, 8]),      # BASELINE cfg 1 (256x256)
    (2, 4, 24, 40, [16, 8, 8]),       # rectangular, batch 2
    (1, 8, 60, 80, [32, 16, 16]),     # BASELINE cfg 4 grid (640x480)
    (1, 2, 16, 16, [8, 4, 2]),
    (1, 8, 16, 16, [4, 4]),           # two levels
    (1, 8, 8, 8, [16]),               # single (dense) level
])
def test_qtatt_b(dev, B, nh, h, w, topks):
    lv = len(topks)
    qs, ks, vs, wt = synth.qtatt_inputs(B, nh * 32, h, w, lv, seed=11)
    ref, aux = qtatt.qtatt_b(qs, ks, vs, wt, topks, nh, return_aux=True)
    out, tk_idx, tk_sc = F.qtatt_forward(_cuda(qs, dev), _cuda(ks, dev), _cuda(vs, dev), topks, nh,
                                         weight=wt.to(dev), attn_type='B', return_topk=True)
    _check_levels(out, tk_idx, tk_sc, ref, aux, h, w, lv, 'QTAttB')


@pytest.mark.parametrize('B,nh,h,w,topks', [(1, 8, 32, 32, [32, 16, 8]), (2, 4, 16, 24, [8, 8, 4]), (1, 8, 16, 16, [8, 8])])
def test_qtatt_a(dev, B, nh, h, w, topks):
    lv = len(topks)
    qs, ks, vs, _ = synth.qtatt_inputs(B, nh * 32, h, w, lv, seed=12)
    ref, aux = qtatt.qtatt_a(qs, ks, vs, topks, nh, return_aux=True)
    out, tk_idx, tk_sc = F.qtatt_forward(_cuda(qs, dev), _cuda(ks, dev), _cuda(vs, dev), topks, nh,
                                         attn_type='A', return_topk=True)
    _check_levels(out, tk_idx, tk_sc, ref, aux, h, w, lv, 'QTAttA')


def test_qtatt_b_module_state_dict(dev):
    m = casmtr_b200.QTAttB(8, 32, scale=3, topks=[32, 16, 8]).to(dev)
    assert list(m.state_dict().keys()) == ['weight']
    qs, ks, vs, wt = synth.qtatt_inputs(1, 256, 32, 32, 3, seed=13)
    m.load_state_dict({'weight': wt})
    with torch.no_grad():
        out = m(_cuda(qs, dev), _cuda(ks, dev), _cuda(vs, dev))
    assert out.shape == (1, 1024, 8, 32)
    assert (out.cpu() - qtatt.qtatt_b(qs, ks, vs, wt, [32, 16, 8], 8)).abs().max() < TOL


@pytest.mark.parametrize('B,nh,h,w,rel,dil', [(2, 4, 16, 16, False, 1), (1, 4, 24, 32, True, 1), (1, 2, 16, 16, False, 2), (1, 4, 52, 52, False, 1)])
def test_cascade_qtatt_b(dev, B, nh, h, w, rel, dil):
    C = nh * 32
    d = synth.cascade_inputs(B, C, h, w, seed=21)
    g = torch.Generator().manual_seed(5)
    v = torch.randn(B, C, h, w, generator=g)
    rp = torch.randn(B, nh, h * w, 100, generator=g) if rel else None
    ref_m, ref_i = qtatt.cascade_qtatt_b(d['feat0'], d['feat1'], v, d['topk_pos01'], rp, nh, dil)
    m = casmtr_b200.CascadeQTAttB(nh, 32, dilated=dil)
    out_m, out_i = m(d['feat0'].to(dev), d['feat1'].to(dev), v.to(dev), d['topk_pos01'].to(dev), None if rp is None else rp.to(dev))
    assert torch.equal(out_i.cpu(), ref_i)                      # integer work: bit-exact
    assert (out_m.cpu() - ref_m).abs().max() < TOL


def test_side_stream_overlap_is_equivalent_and_capturable(dev):
    """casmtr_set_overlap: the finer levels' transposes on the library's side stream must give bit-identical results, stay
    ordered with the caller's stream (inputs produced just before the call, outputs consumed just after it) and be legal
    inside a CUDA-graph capture of the caller's stream."""
    topks, nh, h, w = [16, 8, 8], 4, 32, 48
    qs, ks, vs, wt = synth.qtatt_inputs(2, nh * 32, h, w, 3, seed=17)
    qs, ks, vs, wt = _cuda(qs, dev), _cuda(ks, dev), _cuda(vs, dev), wt.to(dev)
    prev = F.set_overlap(False)
    try:
        want = F.qtatt_forward(qs, ks, vs, topks, nh, weight=wt)
        F.set_overlap(True)
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(3):                                  # back-to-back calls rotate through the side lanes
                q2 = [t * 1.0 for t in qs]                      # produced on the caller's stream right before the call
                got = F.qtatt_forward(q2, ks, vs, topks, nh, weight=wt)
                total = got.sum()                               # consumed right after it
        s.synchronize()
        assert torch.equal(got, want) and torch.isfinite(total)
        g = torch.cuda.CUDAGraph()
        static_q = [t.clone() for t in qs]
        with torch.cuda.graph(g):
            cap = F.qtatt_forward(static_q, ks, vs, topks, nh, weight=wt)
        cap.zero_()
        g.replay()
        torch.cuda.synchronize(dev)
        assert torch.equal(cap, want)
        for t in static_q:
            t.mul_(0.5)
        g.replay()
        torch.cuda.synchronize(dev)
        F.set_overlap(False)
        assert torch.equal(cap, F.qtatt_forward(static_q, ks, vs, topks, nh, weight=wt))
    finally:
        F.set_overlap(prev)
