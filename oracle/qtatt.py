"""Oracle restatement of the QuadTree attention modules (torch CPU, fp32).

Test infrastructure only -- see oracle/__init__.py.

Everything is computed head-major ([B, nh, tokens, D]) with explicit gathers,
which is a different formulation from the reference's (token-major + custom
ops) but the same arithmetic; results are returned in the reference's layouts.
Reference: cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py
(abbreviated ``qta.py`` below).
"""
import torch


# ----------------------------------------------------------------------------- layout helpers
def tokens(x, nhead):
    """[B,C,H,W] -> [B,nh,H*W,D]; channel c = head*D + d (qta.py:166-168)."""
    B, C, H, W = x.shape
    return x.reshape(B, nhead, C // nhead, H * W).transpose(2, 3)


def children(x, nhead):
    """[B,C,H,W] -> [B,nh,(H/2*W/2),4,D]; child f = 2*t1 + t2 (qta.py:188-189)."""
    B, C, H, W = x.shape
    D = C // nhead
    x = x.reshape(B, nhead, D, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 4, 6, 2)
    return x.reshape(B, nhead, (H // 2) * (W // 2), 4, D)


def quad_to_raster(x, hp, wp):
    """[B,nh,hp*wp,4,*] (child-major) -> [B,nh,(2hp*2wp),*] raster
    ('b (h w) (t1 t2) .. -> b (h t1 w t2) ..', qta.py:226,284)."""
    B, nh = x.shape[:2]
    rest = tuple(x.shape[4:])
    x = x.reshape(B, nh, hp, wp, 2, 2, *rest)
    nd = len(rest)
    x = x.permute(0, 1, 2, 4, 3, 5, *range(6, 6 + nd))
    return x.reshape(B, nh, 4 * hp * wp, *rest)


def expand_candidates(prev_idx, w_prev, w_cur, dil=1):
    """Previous-level key indices [..., k] -> the 4 children of each at the
    current key grid, flat order k*4+f, f = 2*x + y (qta.py:191-199)."""
    r = torch.div(prev_idx, w_prev, rounding_mode='trunc') * 2
    c = (prev_idx % w_prev) * 2
    cand = torch.stack([(r + x) * w_cur + c + y for x in (0, dil) for y in (0, dil)], dim=-1)
    return cand.reshape(*prev_idx.shape[:-1], prev_idx.shape[-1] * 4)


def gather_rows(t, idx):
    """t [B,nh,L,D], idx [B,nh,N,K] -> [B,nh,N,K,D]."""
    B, nh, L, D = t.shape
    N, K = idx.shape[2:]
    g = torch.gather(t, 2, idx.reshape(B, nh, N * K, 1).expand(B, nh, N * K, D))
    return g.reshape(B, nh, N, K, D)


def _api_msg(m):      # [B,nh,L,D] -> [B,L,nh,D]
    return m.permute(0, 2, 1, 3).contiguous()


def _api_idx(i):      # [B,nh,L,k] -> [B,L,k,nh]
    return i.permute(0, 2, 3, 1).contiguous()


# ----------------------------------------------------------------------------- QTAttB
def qtatt_b(queries, keys, values, weight, topks, nhead, rel_pos=None, return_aux=False):
    """QTAttB.forward (qta.py:231-286; coarse :161-178, fine :180-229, merge :262-284).

    queries/keys/values: lists finest->coarsest of [B,C,H_l,W_l]; weight [levels].
    Returns message [B, L_finest, nh, D]; with return_aux also a dict holding
    per-level 'topk_idx' ([B,L_l,k,nh] int64, raster, key index at level l),
    'topk_score' and 'messages' (per-level, child-major like the reference).
    rel_pos (optional): list per level (coarsest first): level 0 broadcastable to
    [B,L,S,nh]; finer levels [1,nh,L0,L1] (qta.py:172-173, 211-215).
    """
    n_lv = len(queries)
    msgs, aux_idx, aux_score = [], [], []
    prev_idx = None
    w_prev = None
    acc = None
    wsm = torch.softmax(weight, dim=0)
    for i in range(n_lv):
        q, k, v = queries[n_lv - 1 - i], keys[n_lv - 1 - i], values[n_lv - 1 - i]
        B, C, h0, w0 = q.shape
        h1, w1 = k.shape[2:]
        D = C // nhead
        temp = 1.0 / D ** 0.5
        kt, vt = tokens(k, nhead), tokens(v, nhead)
        if i == 0:
            qt = tokens(q, nhead)
            logits = torch.matmul(qt, kt.transpose(-1, -2)) * temp          # [B,nh,L,S]
            if rel_pos is not None and rel_pos[0] is not None:
                logits = logits + rel_pos[0].expand(B, h0 * w0, h1 * w1, nhead).permute(0, 3, 1, 2)
            A = torch.softmax(logits, dim=-1)
            sc, ix = torch.topk(A, k=topks[0], dim=-1, largest=True)      # sorted descending
            m = torch.matmul(A, vt)                                        # all S keys (type B)
            acc = m * wsm[0]
            msgs.append(_api_msg(m))
        else:
            hp, wp = h0 // 2, w0 // 2
            qc = children(q, nhead)                                        # [B,nh,Np,4,D]
            cand = expand_candidates(prev_idx, w_prev, w1)                 # [B,nh,Np,K]
            kc, vc = gather_rows(kt, cand), gather_rows(vt, cand)
            logits = torch.matmul(qc, kc.transpose(-1, -2)) * temp         # [B,nh,Np,4,K]
            if rel_pos is not None and rel_pos[i] is not None:
                rp = children_bias(rel_pos[i].expand(B, nhead, h0 * w0, h1 * w1), h0, w0)   # [B,nh,Np,4,L1]
                logits = logits + torch.gather(rp, 4, cand.unsqueeze(3).expand(-1, -1, -1, 4, -1))
            A = torch.softmax(logits, dim=-1)
            sc, lx = torch.topk(A, k=topks[i], dim=-1, largest=True)
            m = torch.matmul(A, vc)                                        # [B,nh,Np,4,D]
            ix = torch.gather(cand.unsqueeze(3).expand(-1, -1, -1, 4, -1), 4, lx)
            ix = quad_to_raster(ix, hp, wp)                                # [B,nh,L0,k]
            sc = quad_to_raster(sc, hp, wp)
            acc = quad_to_raster(acc.unsqueeze(3) + m * wsm[i], hp, wp)    # parent broadcast (qta.py:280-284)
            msgs.append(m.permute(0, 2, 3, 1, 4).contiguous())             # [B,Np,4,nh,D]
        prev_idx, w_prev = ix, w1
        aux_idx.append(_api_idx(ix))
        aux_score.append(_api_idx(sc))
    out = _api_msg(acc)
    if return_aux:
        return out, {'topk_idx': aux_idx, 'topk_score': aux_score, 'messages': msgs}
    return out


def children_bias(rp, h0, w0):
    """rel_pos [B,nh,h0*w0,L1] raster -> [B,nh,Np,4,L1] child-major (qta.py:212-213)."""
    B, nh, _, L1 = rp.shape
    rp = rp.reshape(B, nh, h0 // 2, 2, w0 // 2, 2, L1).permute(0, 1, 2, 4, 3, 5, 6)
    return rp.reshape(B, nh, (h0 // 2) * (w0 // 2), 4, L1)


# ----------------------------------------------------------------------------- QTAttA
def qtatt_a(queries, keys, values, topks, nhead, return_aux=False):
    """QTAttA.forward (qta.py:101-140; coarse :25-44, fine :46-99).

    Differences from B: the coarse/intermediate messages exclude the selected
    top-k entries (:37-42, :81-84); the fine softmax runs over the 4 children of
    each parent candidate (:72-74) and is multiplied by the parent's score
    (:76-77); the merge is a plain sum (:130-138).
    """
    n_lv = len(queries)
    aux_idx, aux_score = [], []
    prev_idx = prev_sc = w_prev = acc = None
    for i in range(n_lv):
        q, k, v = queries[n_lv - 1 - i], keys[n_lv - 1 - i], values[n_lv - 1 - i]
        B, C, h0, w0 = q.shape
        h1, w1 = k.shape[2:]
        D = C // nhead
        temp = 1.0 / D ** 0.5
        kt, vt = tokens(k, nhead), tokens(v, nhead)
        final = (i == n_lv - 1) and i > 0
        if i == 0:
            qt = tokens(q, nhead)
            A = torch.softmax(torch.matmul(qt, kt.transpose(-1, -2)) * temp, dim=-1)
            sc, ix = torch.topk(A, k=topks[0], dim=-1, largest=True)
            keep = torch.ones_like(A).scatter_(-1, ix, 0.0)
            acc = torch.matmul(A * keep, vt)
        else:
            hp, wp = h0 // 2, w0 // 2
            kp = prev_idx.shape[-1]
            qc = children(q, nhead)
            cand = expand_candidates(prev_idx, w_prev, w1)                 # [B,nh,Np,kp*4]
            kc, vc = gather_rows(kt, cand), gather_rows(vt, cand)
            logits = torch.matmul(qc, kc.transpose(-1, -2)) * temp         # [B,nh,Np,4,kp*4]
            Np = logits.shape[2]
            A = torch.softmax(logits.reshape(B, nhead, Np, 4, kp, 4), dim=-1)
            A = (A * prev_sc.reshape(B, nhead, Np, 1, kp, 1)).reshape(B, nhead, Np, 4, kp * 4)
            sc, lx = torch.topk(A, k=topks[i], dim=-1, largest=True)
            if not final:
                keep = torch.ones_like(A).scatter_(-1, lx, 0.0)
                m = torch.matmul(A * keep, vc)
            else:
                m = torch.matmul(A, vc)
            ix = torch.gather(cand.unsqueeze(3).expand(-1, -1, -1, 4, -1), 4, lx)
            ix = quad_to_raster(ix, hp, wp)
            sc = quad_to_raster(sc, hp, wp)
            acc = quad_to_raster(acc.unsqueeze(3) + m, hp, wp)
        prev_idx, prev_sc, w_prev = ix, sc, w1
        aux_idx.append(_api_idx(ix))
        aux_score.append(_api_idx(sc))
    out = _api_msg(acc)
    if return_aux:
        return out, {'topk_idx': aux_idx, 'topk_score': aux_score}
    return out


# ----------------------------------------------------------------------------- CascadeQTAttB
def cascade_window_idx(topk_pos, h1, w1, dil=1):
    """topk_pos [B,Np,k,2] (row,col at the previous level) -> candidate key indices
    [B,Np,4k] int64, order k*4+f, clamped to the key grid (qta.py:418-429)."""
    r = topk_pos[..., 0] * 2
    c = topk_pos[..., 1] * 2
    cand = torch.stack([(r + x) * w1 + c + y for x in (0, dil) for y in (0, dil)], dim=-1)
    cand = cand.reshape(*topk_pos.shape[:2], -1)
    return torch.clamp(cand, min=0, max=h1 * w1 - 1)


def cascade_qtatt_b(query, key, value, topk_pos, rel_pos, nhead, dilated=1):
    """CascadeQTAttB.forward (qta.py:400-452).

    query [B,C,h0,w0]; key/value [B,C,h1,w1]; topk_pos [B,(h0/2*w0/2),k,2] int64;
    rel_pos None or [B,nh,h0*w0,4k].  Returns (message [B,h0*w0,C] raster,
    upsampled_idx [B,h0*w0,4k] int64).
    """
    B, C, h0, w0 = query.shape
    h1, w1 = key.shape[2:]
    D = C // nhead
    hp, wp = h0 // 2, w0 // 2
    dil = 1 if dilated is None else dilated
    cand = cascade_window_idx(topk_pos, h1, w1, dil)                      # [B,Np,K]
    K = cand.shape[-1]
    cand_h = cand.unsqueeze(1).expand(B, nhead, hp * wp, K)
    kt, vt = tokens(key, nhead), tokens(value, nhead)
    qc = children(query, nhead)
    kc, vc = gather_rows(kt, cand_h), gather_rows(vt, cand_h)
    logits = torch.matmul(qc, kc.transpose(-1, -2)) * (1.0 / D ** 0.5)      # [B,nh,Np,4,K]
    if rel_pos is not None:
        logits = logits + children_bias(rel_pos.reshape(B, nhead, h0 * w0, K), h0, w0)
    A = torch.softmax(logits, dim=-1)
    m = quad_to_raster(torch.matmul(A, vc), hp, wp)                        # [B,nh,L0,D]
    message = m.permute(0, 2, 1, 3).reshape(B, h0 * w0, C).contiguous()
    up = quad_to_raster(cand.reshape(B, 1, hp * wp, 1, K).expand(B, 1, hp * wp, 4, K), hp, wp)
    return message, up.reshape(B, h0 * w0, K).contiguous()


# ----------------------------------------------------------------------------- QTAttGuided (single level)
def qtatt_guided(query, key, value, topk_pos, weight, nhead, reference_order=True):
    """QTAttGuided.forward on a ONE-level pyramid (qta.py:289-389; the only case the reference's merge can run, see :385).
    query [B,C,h0,w0], key / value [B,C,h1,w1], topk_pos [2,B,(h0/2*w0/2),K,nh] (row, col at half the key resolution), weight
    [scale].  :311-318 candidates = the 2x2 children of every key cell, order k*4+f; :325-327 scaled logits; :335 softmax over the 4K
    candidates; :342 message = A.V; :375-380 message * softmax(weight)[0]; :385 the rearrange with H = queries[-0].shape[2] = h0
    (reference_order=True reproduces it; False gives the raster order of the query grid).  Returns [B, h0*w0, nh, D]."""
    B, C, h0, w0 = query.shape
    h1, w1 = key.shape[2:]
    D = C // nhead
    hp, wp = h0 // 2, w0 // 2
    Np = hp * wp
    r2, c2 = topk_pos[0] * 2, topk_pos[1] * 2                                   # [B,Np,K,nh]
    cand = torch.stack([(r2 + x) * w1 + c2 + y for x in (0, 1) for y in (0, 1)], dim=3)          # [B,Np,K,4,nh]
    K4 = cand.shape[2] * 4
    cand = cand.reshape(B, Np, K4, nhead).permute(0, 3, 1, 2)                    # [B,nh,Np,4K]
    kt, vt = tokens(key, nhead), tokens(value, nhead)
    qc = children(query, nhead)                                                 # [B,nh,Np,4,D]
    kc, vc = gather_rows(kt, cand), gather_rows(vt, cand)
    A = torch.softmax(torch.matmul(qc, kc.transpose(-1, -2)) * (1.0 / D ** 0.5), dim=-1)
    m = torch.matmul(A, vc) * torch.softmax(weight, dim=0)[0]                   # [B,nh,Np,4,D]
    m = m.permute(0, 2, 3, 1, 4)                                                # [B,Np,4,nh,D]
    H = h0 if reference_order else hp
    W = Np // H
    out = m.reshape(B, H, W, 2, 2, nhead, D).permute(0, 1, 3, 2, 4, 5, 6)       # b (H t1 W t2) h d
    return out.reshape(B, h0 * w0, nhead, D).contiguous()
