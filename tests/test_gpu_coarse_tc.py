"""GPU parity of the tensor-core coarsest quadtree level (csrc/qtatt_coarse_tc.cu: tcgen05 S = Q K^T and O = P V with
error-compensated TF32 splits, exact soft-max / top-k in between) against the CPU oracle and against the fp32 SIMT kernel
(csrc/qtatt_coarse.cu, selected with CASMTR_QT_SIMT_COARSE): identical top-k key sets (fp32 near-ties excepted, oracle/compare.py),
top-k scores within 1e-5, messages within 1e-4 of the oracle.
Reference: QTAttB / QTAttA.process_coarse_level, cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:161-178, :25-44."""
import pytest
import torch

from casmtr_b200 import _lib, synth
from casmtr_b200 import functional as F
from oracle import qtatt
from oracle.compare import check_qtatt_levels, topk_bad_rows

pytestmark = pytest.mark.gpu


def _cuda(lst, dev):
    return [t.to(dev) for t in lst]


# (B, nhead, h, w, topks): coarsest level = (h/4) x (w/4) keys -> kernel variant <rows per CTA, values per lane>
SHAPES = [
    (2, 8, 32, 32, [32, 16, 8]),        # 64 keys    <64, 8>   (256x256 image, BASELINE cfg 1)
    (3, 4, 48, 64, [16, 8, 8]),         # 192 keys   <64, 8>, three batch elements, rectangular
    (1, 8, 60, 80, [32, 16, 16]),       # 300 keys   <64, 16>  (640x480, cfg 4); 300 = 4 full chunks + 44
    (2, 8, 104, 104, [32, 16, 8]),      # 676 keys   <64, 22>  (832x832, cfg 2): 11 row tiles, the last one holds 36 rows
    (1, 4, 128, 128, [32, 16, 8]),      # 1024 keys  <32, 32>  (1024x1024, cfg 5)
    (1, 2, 144, 144, [32, 16, 8]),      # 1296 keys  <32, 42>  (1152x1152, cfg 5)
]


@pytest.mark.parametrize('B,nh,h,w,topks', SHAPES)
@pytest.mark.parametrize('typ', ['B', 'A'])
def test_tc_coarse_level_vs_oracle_and_simt(dev, B, nh, h, w, topks, typ):
    if typ == 'A' and h * w > 64 * 80:
        pytest.skip('type A is checked on the smaller shapes')
    qs, ks, vs, wt = synth.qtatt_inputs(B, nh * 32, h, w, 3, seed=31 + h)
    if typ == 'B':
        ref, aux = qtatt.qtatt_b(qs, ks, vs, wt, topks, nh, return_aux=True)
    else:
        ref, aux = qtatt.qtatt_a(qs, ks, vs, topks, nh, return_aux=True)
    kw = dict(weight=wt.to(dev) if typ == 'B' else None, attn_type=typ, return_topk=True)
    out, idx, sc = F.qtatt_forward(_cuda(qs, dev), _cuda(ks, dev), _cuda(vs, dev), topks, nh, **kw)
    out_s, idx_s, sc_s = F.qtatt_forward(_cuda(qs, dev), _cuda(ks, dev), _cuda(vs, dev), topks, nh, flags=_lib.QT_SIMT_COARSE, **kw)
    torch.cuda.synchronize()
    check_qtatt_levels(out, idx, sc, ref, aux, h, w, 3, f'tc QTAtt{typ} {h}x{w}', tol=1e-4)
    # the two kernels against each other at the coarsest level: same sets, scores to fp32 rounding
    bad = topk_bad_rows(idx[0].cpu(), sc[0].cpu(), idx_s[0].cpu(), sc_s[0].cpu(), 'tc vs simt')
    assert bad.float().mean() <= 1e-3
    d = (torch.sort(sc[0], dim=2)[0] - torch.sort(sc_s[0], dim=2)[0]).abs().amax(dim=2)
    assert d[~bad.to(dev)].max() < 1e-6


def test_tc_coarse_single_level_message(dev):
    """One-level call (dense attention only): the message IS the second GEMM's output / row sum, no finer level hides an error."""
    B, nh, h, w = 2, 8, 16, 24                # 384 keys
    qs, ks, vs, wt = synth.qtatt_inputs(B, nh * 32, h, w, 1, seed=5)
    ref, aux = qtatt.qtatt_b(qs, ks, vs, wt, [32], nh, return_aux=True)
    out, idx, sc = F.qtatt_forward(_cuda(qs, dev), _cuda(ks, dev), _cuda(vs, dev), [32], nh, weight=wt.to(dev), return_topk=True)
    assert (out.cpu() - ref).abs().max() < 2e-5        # three-term TF32 splits in both GEMMs: ~2^-21 relative per product (observed 8e-6 on values of magnitude 3)
    assert torch.equal(torch.sort(idx[0].cpu(), dim=2)[0], torch.sort(aux['topk_idx'][0], dim=2)[0])
    # exported lists are in torch.topk's descending order wherever two neighbouring scores are not within fp32 rounding of each other
    gi, gs, ri = idx[0].cpu(), sc[0].cpu(), aux['topk_idx'][0]
    assert (gs[:, :, :-1] >= gs[:, :, 1:]).all()
    clear = (aux['topk_score'][0][:, :, :-1] - aux['topk_score'][0][:, :, 1:]) > 1e-6 * aux['topk_score'][0][:, :, :-1]
    row_clear = clear.all(dim=2, keepdim=True).expand_as(gi)
    assert torch.equal(gi[row_clear], ri[row_clear])


def test_tc_coarse_is_what_runs_by_default(dev):
    """The default path for a 676-key level is the tensor-core kernel (launch accounting shows the extra prep launch)."""
    qs, ks, vs, wt = synth.qtatt_inputs(1, 256, 104, 104, 3, seed=3)
    args = (_cuda(qs, dev), _cuda(ks, dev), _cuda(vs, dev), [32, 16, 8], 8)
    n0 = F.launch_count()
    F.qtatt_forward(*args, weight=wt.to(dev))
    n1 = F.launch_count()
    F.qtatt_forward(*args, weight=wt.to(dev), flags=_lib.QT_SIMT_COARSE)
    n2 = F.launch_count()
    assert (n1 - n0) == (n2 - n1) + 1
