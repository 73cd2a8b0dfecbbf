"""CPU restatements of the SURVEY section 8f "next" rows (test infrastructure only -- see oracle/__init__.py).

  coarse_match_stats        src/model/functions/coarse_matching.py:60-75   (CoarseMatching.forward, inference statistics)
  coarse_matches            src/model/functions/coarse_matching.py:91-153  (CoarseMatching.get_coarse_match, inference)
  fine_windows              src/model/functions/fine_matching.py:47-66     (CascadeFinePreprocess.forward)
  quadtree_attention_layer  src/model/modules/quadtree_attention.py:68-99  (QuadtreeAttention.forward, type A / B)
  cascade_attention_layer   src/model/modules/quadtree_attention.py:152-171 (CascadeQuadtreeAttention.forward)
  relative_pe               src/model/modules/transformer.py:473-509       (CascadeFeatureTransformer.get_relative_pe)

Pinned against outputs of the reference's own modules (tests/golden/widen_*.npz, tests/golden/make_golden.py)."""
import torch
import torch.nn.functional as tF

from . import qtatt


def coarse_match_stats(feat0, feat1, temperature, dtype=torch.float32, mask0=None, mask1=None):
    """feat0 [B,L,C], feat1 [B,S,C] -> next_conf01 [B,L], next_idx01 [B,L], next_conf10 [B,S], next_idx10 [B,S] and the
    dense similarity (for tie analysis).  :60 normalise, :63 einsum / T, :64-65 padding masks (bool [B,L] / [B,S]) filled with
    -INF = -1e9 (:9), :66-67 the two softmaxes, :70-71 max."""
    C = feat0.shape[-1]
    sim = torch.einsum('nlc,nsc->nls', feat0.to(dtype) / C ** 0.5, feat1.to(dtype) / C ** 0.5) / temperature
    if mask0 is not None:
        sim = sim.masked_fill(~(mask0[..., None] * mask1[:, None]).bool(), -1e9)
    p10, p01 = torch.softmax(sim, 1), torch.softmax(sim, 2)
    c01, i01 = p01.max(dim=2)
    c10, i10 = p10.max(dim=1)
    return c01, i01, c10, i10, sim


def coarse_matches(feat0, feat1, temperature, thr, border_rm, hw0, hw1, hw0_i, mask0=None, mask1=None, pad_mask0=None, pad_mask1=None,
                   scale0=None, scale1=None):
    """CoarseMatching.get_coarse_match, inference (coarse_matching.py:91-153) on the dense confidence matrix conf = softmax_i * softmax_j
    (:66-68).  :113 conf > thr; :114-119 border removal on the source AND target grids (mask_border: rows / cols < b or >= size - b;
    mask_border_with_padding when pad_mask0 / pad_mask1 [B,h,w] are given: the valid extents per sample replace the sizes);
    :122 mutual nearest neighbours (conf equals its row maximum and its column maximum); :126-129 at most one match per row, in
    torch.where order; :135-139 keypoints.  Returns dict b_ids, i_ids, j_ids, mconf, mkpts0_c, mkpts1_c."""
    from .cascade import valid_extent
    C = feat0.shape[-1]
    sim = torch.einsum('nlc,nsc->nls', feat0 / C ** 0.5, feat1 / C ** 0.5) / temperature
    if mask0 is not None:
        sim = sim.masked_fill(~(mask0[..., None] * mask1[:, None]).bool(), -1e9)
    conf = torch.softmax(sim, 1) * torch.softmax(sim, 2)
    B, L, S = conf.shape
    (h0, w0), (h1, w1) = hw0, hw1
    mask = conf > thr
    b = border_rm
    if b > 0:
        y0, x0 = torch.arange(L) // w0, torch.arange(L) % w0
        y1, x1 = torch.arange(S) // w1, torch.arange(S) % w1
        if pad_mask0 is None:
            bad0 = ((y0 < b) | (x0 < b) | (y0 >= h0 - b) | (x0 >= w0 - b)).unsqueeze(0).expand(B, L)
            bad1 = ((y1 < b) | (x1 < b) | (y1 >= h1 - b) | (x1 >= w1 - b)).unsqueeze(0).expand(B, S)
        else:
            h0s, w0s = valid_extent(pad_mask0)
            h1s, w1s = valid_extent(pad_mask1)
            bad0 = (y0 < b)[None] | (x0 < b)[None] | (y0[None] >= h0s[:, None] - b) | (x0[None] >= w0s[:, None] - b)
            bad1 = (y1 < b)[None] | (x1 < b)[None] | (y1[None] >= h1s[:, None] - b) | (x1[None] >= w1s[:, None] - b)
        mask = mask & ~bad0[:, :, None] & ~bad1[:, None, :]
    mask = mask & (conf == conf.max(dim=2, keepdim=True)[0]) & (conf == conf.max(dim=1, keepdim=True)[0])
    mask_v, all_j = mask.max(dim=2)
    b_ids, i_ids = torch.where(mask_v)
    j_ids = all_j[b_ids, i_ids]
    mconf = conf[b_ids, i_ids, j_ids]
    scale = hw0_i[0] / h0
    s0 = scale * scale0[b_ids] if scale0 is not None else scale
    s1 = scale * scale1[b_ids] if scale1 is not None else scale
    mk0 = torch.stack([i_ids % w0, torch.div(i_ids, w0, rounding_mode='trunc')], dim=1) * s0
    mk1 = torch.stack([j_ids % w1, torch.div(j_ids, w1, rounding_mode='trunc')], dim=1) * s1
    return {'b_ids': b_ids, 'i_ids': i_ids, 'j_ids': j_ids, 'mconf': mconf, 'mkpts0_c': mk0, 'mkpts1_c': mk1}


def fine_windows(feat_f, b_ids, ids, stride, W):
    """feat_f [B,C,Hf,Wf] -> [M, W*W, C]: :47 F.unfold(kernel W, stride, padding W//2), :49 'n (c ww) l -> n l ww c', :55 select."""
    B, C = feat_f.shape[:2]
    u = tF.unfold(feat_f, kernel_size=(W, W), stride=stride, padding=W // 2)          # [B, C*WW, L]
    return u.reshape(B, C, W * W, -1).permute(0, 3, 2, 1)[b_ids, ids]


def fine_preprocess(feat_f0, feat_f1, feat_c0, feat_c1, b_ids, i_ids, j_ids, stride, W, down_proj=None, merge_feat=None):
    """:47-66 incl. the optional coarse-feature concat (down_proj / merge_feat = (weight, bias) of the two nn.Linear)."""
    f0, f1 = fine_windows(feat_f0, b_ids, i_ids, stride, W), fine_windows(feat_f1, b_ids, j_ids, stride, W)
    if down_proj is not None:
        cw = tF.linear(torch.cat([feat_c0[b_ids, i_ids], feat_c1[b_ids, j_ids]], 0), *down_proj)
        cf = tF.linear(torch.cat([torch.cat([f0, f1], 0), cw.unsqueeze(1).expand(-1, W * W, -1)], -1), *merge_feat)
        f0, f1 = torch.chunk(cf, 2, dim=0)
    return f0, f1


def _nchw(t, H, W):
    B, N, C = t.shape
    return t.permute(0, 2, 1).reshape(B, C, H, W).contiguous()


def _pyramid(x, levels):
    out = []
    for i in range(levels):
        out.append(x.float())
        if i != levels - 1:
            x = tF.avg_pool2d(x, kernel_size=2, stride=2)            # :86-89
    return out


def quadtree_attention_layer(x, target, H, W, sd, nhead, topks, scale, attn_type='B', H1=None, W1=None):
    """sd: state dict with q_proj.weight [C,C,1,1] ... proj.weight / proj.bias, py_att.weight."""
    H1, W1 = H1 or H, W1 or W
    B, N, C = x.shape
    q = tF.conv2d(_nchw(x, H, W), sd['q_proj.weight'], sd.get('q_proj.bias'))          # :77-79
    k = tF.conv2d(_nchw(target, H1, W1), sd['k_proj.weight'], sd.get('k_proj.bias'))
    v = tF.conv2d(_nchw(target, H1, W1), sd['v_proj.weight'], sd.get('v_proj.bias'))
    qs, ks, vs = _pyramid(q, scale), _pyramid(k, scale), _pyramid(v, scale)
    if attn_type == 'A':
        msg = qtatt.qtatt_a(qs, ks, vs, topks, nhead)
    else:
        msg = qtatt.qtatt_b(qs, ks, vs, sd['py_att.weight'], topks, nhead)
    return tF.linear(msg.reshape(B, -1, C), sd['proj.weight'], sd['proj.bias'])          # :96


def cascade_attention_layer(x, target, H, W, idx, sd, nhead, rel_pos=None, H1=None, W1=None, dilated=1):
    H1, W1 = H1 or H, W1 or W
    B, N, C = x.shape
    q = tF.conv2d(_nchw(x, H, W), sd['q_proj.weight'], sd.get('q_proj.bias'))          # :159-161
    k = tF.conv2d(_nchw(target, H1, W1), sd['k_proj.weight'], sd.get('k_proj.bias'))
    v = tF.conv2d(_nchw(target, H1, W1), sd['v_proj.weight'], sd.get('v_proj.bias'))
    msg, up = qtatt.cascade_qtatt_b(q.float(), k.float(), v.float(), idx, rel_pos, nhead, dilated)
    return tF.linear(msg.reshape(B, -1, C), sd['proj.weight'], sd['proj.bias']), up


def relative_pe(window_pos, tgt_idx, w_table, h_table, LB, hw8, w8_other, H):
    """window_pos [B,(H/2)*(W/2),k,2] int64 (row, col on the previous level of the other image), tgt_idx [B,h8*w8] int64 (the
    query image's 1/8 match), w_table / h_table [n_emb,nhead] -> [B,nhead,H*W,4k] fp32.
    :474-478 s, W1 and the query's position inside its 1/8 cell (x = col % s, y = row % s); :480-485 the match, brought to the
    current level (* s + s//2 - 1); :487-499 the window's 2x2 children as flat indices, back to (x, y); :500-508
    index = src - (tgt - window + LB) + 2 LB into the two tables, x table + y table."""
    h8, w8 = hw8
    s = H // h8
    W, W1 = w8 * s, w8_other * s
    B, Np, k, _ = window_pos.shape
    Y, X = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    Y, X = Y.reshape(-1), X.reshape(-1)                                              # [HW]
    t = tgt_idx[:, (Y // s) * w8 + X // s]                                           # [B,HW]
    tx, ty = (t % w8_other) * s + (s // 2 - 1), torch.div(t, w8_other, rounding_mode='trunc') * s + (s // 2 - 1)
    wp = window_pos[:, (Y // 2) * (W // 2) + X // 2] * 2                             # [B,HW,k,2]
    kids = torch.stack([(wp[..., 0] + x) * W1 + wp[..., 1] + y for x in (0, 1) for y in (0, 1)], dim=3).reshape(B, H * W, 4 * k)
    kx, ky = kids % W1, torch.div(kids, W1, rounding_mode='trunc')
    rx = (X % s)[None, :, None] - (tx[:, :, None] - kx + LB) + 2 * LB
    ry = (Y % s)[None, :, None] - (ty[:, :, None] - ky + LB) + 2 * LB
    return (w_table[rx] + h_table[ry]).permute(0, 3, 1, 2).contiguous()               # [B,nhead,HW,4k]
