"""CPU restatements of the SURVEY section 8f "next" rows (test infrastructure only -- see oracle/__init__.py).

  coarse_match_stats        src/model/functions/coarse_matching.py:60-75   (CoarseMatching.forward, inference statistics)
  fine_windows              src/model/functions/fine_matching.py:47-66     (CascadeFinePreprocess.forward)
  quadtree_attention_layer  src/model/modules/quadtree_attention.py:68-99  (QuadtreeAttention.forward, type A / B)
  cascade_attention_layer   src/model/modules/quadtree_attention.py:152-171 (CascadeQuadtreeAttention.forward)

Pinned against outputs of the reference's own modules (tests/golden/widen_*.npz, tests/golden/make_golden.py)."""
import torch
import torch.nn.functional as tF

from . import qtatt


def coarse_match_stats(feat0, feat1, temperature, dtype=torch.float32):
    """feat0 [B,L,C], feat1 [B,S,C] -> next_conf01 [B,L], next_idx01 [B,L], next_conf10 [B,S], next_idx10 [B,S] and the
    dense similarity (for tie analysis).  :60 normalise, :63 einsum / T, :66-67 the two softmaxes, :70-71 max."""
    C = feat0.shape[-1]
    sim = torch.einsum('nlc,nsc->nls', feat0.to(dtype) / C ** 0.5, feat1.to(dtype) / C ** 0.5) / temperature
    p10, p01 = torch.softmax(sim, 1), torch.softmax(sim, 2)
    c01, i01 = p01.max(dim=2)
    c10, i10 = p10.max(dim=1)
    return c01, i01, c10, i10, sim


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
