"""The reference's GPU data flow around ITS OWN extension kernels (oracle/_ref, built unmodified from /root/reference):
what `cuda_imp` does on a GPU, restated so it can run on the GPU box where /root/reference does not exist.
Test infrastructure / baseline only -- see oracle/__init__.py; used by tests/test_gpu_vs_reference_ext.py and by
bench.py's `gpu_reference_kernels_baseline` (the "reference build on the same B200" number of SURVEY.md section 8d).

Every tensor layout is the one the reference extensions require:
  score_computation_cuda.score_forward(query [B,N,4,H,D], key [B,S,H,D], index [B,N,K,H]) -> [B,N,4,K,H]
  value_aggregation_cuda.value_aggregation_forward(score [B,N',K,H], value [B,S,H,D], index [B,N',K,H], out [B,N',H,D])
  fast_score_computation.score_forward(query [B,N,C], key [B,S,C], index [B,N,K]) -> [B,N,K]
Follows cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:161-286 (QTAttB), :431-452
(CascadeQTAttB) and functions/quadtree_attention.py:6-38 (the two autograd wrappers, forward only).
"""
import torch

from . import build_ref

_EXT = {}


def ext(name):
    if name not in _EXT:
        _EXT[name] = build_ref.load(name)
    return _EXT[name]


def available():
    return all(build_ref.built(n) for n in ('score_computation_cuda', 'value_aggregation_cuda', 'fast_score_computation'))


def _tokens(x, nhead):                       # [B,C,H,W] -> [B,HW,nhead,D]   (:165-167, :185-186)
    B, C, H, W = x.shape
    return x.flatten(2).transpose(1, 2).reshape(B, H * W, nhead, C // nhead).contiguous()


def _children(x, nhead):                     # [B,C,H,W] -> [B,(H/2 W/2),4,nhead,D], child f = 2*t1 + t2   (:188-189)
    B, C, H, W = x.shape
    x = x.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 3, 5, 1)
    return x.reshape(B, (H // 2) * (W // 2), 4, nhead, C // nhead).contiguous()


def _raster(x, hp):                          # [B,(hp wp),4,...] -> [B,(hp 2 wp 2),...]   (:226-227, :284)
    B, Np = x.shape[:2]
    wp = Np // hp
    rest = x.shape[3:]
    x = x.reshape(B, hp, wp, 2, 2, *rest).permute(0, 1, 3, 2, 4, *range(5, 5 + len(rest)))
    return x.reshape(B, hp * 2 * wp * 2, *rest).contiguous()


def _child_index(pos_rc, w1):                # (row, col) [2,B,N,k,H] on the parent grid -> key indices of the 4 children [B,N,4k,H]   (:193-199)
    r2, c2 = pos_rc[0] * 2, pos_rc[1] * 2
    idx = torch.stack([(r2 + x) * w1 + c2 + y for x in (0, 1) for y in (0, 1)], dim=3)           # [B,N,k,4,H]
    B, N, k, _, H = idx.shape
    return idx.reshape(B, N, 4 * k, H).contiguous()


def _value_agg(A, value, idx5):              # A, idx5 [B,N,4,K,H]; value [B,S,H,D] -> [B,N,4,H,D]   (functions/quadtree_attention.py:25-38)
    B, N, F, K, H = A.shape
    out = torch.zeros(B, N * F, H, value.shape[-1], device=A.device, dtype=A.dtype)
    ext('value_aggregation_cuda').value_aggregation_forward(A.reshape(B, N * F, K, H).contiguous(), value,
                                                            idx5.reshape(B, N * F, K, H).contiguous(), out)
    return out.reshape(B, N, F, H, value.shape[-1])


def qtatt_b(queries, keys, values, weight, topks, nhead):
    """QTAttB.forward with the reference kernels.  Pyramids finest first, [B,C,H,W] CUDA fp32 -> message [B,L,nhead,D]."""
    n = len(queries)
    msgs, pos, score = [], None, None
    k_prev = topks[0]
    for i in range(n):
        q, k, v = queries[n - 1 - i], keys[n - 1 - i], values[n - 1 - i]
        B, C, h0, w0 = q.shape
        h1, w1 = k.shape[2:]
        D = C // nhead
        kt, vt = _tokens(k, nhead), _tokens(v, nhead)
        if i == 0:                                                               # dense level (:161-178)
            qk = torch.einsum('nlhd,nshd->nlsh', _tokens(q, nhead), kt) * (1.0 / D ** 0.5)
            A = torch.softmax(qk, dim=-2)
            score, idx = torch.topk(A, dim=-2, k=topks[0], largest=True)
            msgs.append(torch.einsum('nlsh,nshd->nlhd', A, vt))
        else:                                                                    # sparse level (:180-229)
            cand = _child_index(pos, w1)                                         # [B,Np,4k,H]
            qk = ext('score_computation_cuda').score_forward(_children(q, nhead), kt, cand)[0] * (1.0 / D ** 0.5)
            idx5 = cand.unsqueeze(2).repeat(1, 1, 4, 1, 1)                       # [B,Np,4,4k,H]
            A = torch.softmax(qk, dim=-2)
            score, sel = torch.topk(A, dim=-2, k=topks[i], largest=True)
            msgs.append(_value_agg(A, vt, idx5))                                 # [B,Np,4,H,D]
            idx = _raster(torch.gather(idx5, 3, sel), h0 // 2)                   # [B,L,k,H]
            score = _raster(score, h0 // 2)
            k_prev = topks[i]
        pos = torch.stack([torch.div(idx, w1, rounding_mode='trunc'), idx % w1])  # (:253)
    w = torch.softmax(weight, dim=0)                                             # merge (:262-284)
    out = msgs[0] * w[0]
    for i in range(1, n):
        out = _raster(out.unsqueeze(2) + msgs[i] * w[i], queries[n - i].shape[2])
    return out


def cascade_qtatt_b(query, key, value, topk_pos, nhead):
    """CascadeQTAttB.forward (dilated 1, no rel_pos) with the reference kernels -> (message [B,L,C], upsampled_idx [B,L,4k])."""
    B, C, h0, w0 = query.shape
    h1, w1 = key.shape[2:]
    D = C // nhead
    kt, vt = _tokens(key, nhead), _tokens(value, nhead)
    pos = topk_pos.permute(3, 0, 1, 2).unsqueeze(-1).expand(-1, -1, -1, -1, nhead)               # [2,B,Np,k,H]  (:418)
    cand = torch.clamp(_child_index(pos, w1), 0, h1 * w1 - 1)                                    # (:428-429)
    qk = ext('score_computation_cuda').score_forward(_children(query, nhead), kt, cand)[0] * (1.0 / D ** 0.5)
    idx5 = cand.unsqueeze(2).repeat(1, 1, 4, 1, 1)
    A = torch.softmax(qk, dim=-2)
    msg = _raster(_value_agg(A, vt, idx5), h0 // 2).reshape(B, h0 * w0, C)
    up = _raster(idx5[..., 0], h0 // 2)                                                          # [B,L,4k]  (:450)
    return msg, up


def score3d(query, key, index):
    """ScoreComputation.apply (src/model/functions/cascade_functions.py:8-22) -> fast_score_computation.score_forward."""
    return ext('fast_score_computation').score_forward(query.contiguous(), key.contiguous(), index.contiguous())[0]
