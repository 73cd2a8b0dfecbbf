"""Oracle restatement of CascadeMatching (inference branch), NMS and match extraction.

Test infrastructure only -- see oracle/__init__.py.
Reference: src/model/functions/cascade_matching.py (``cm.py``),
src/model/functions/post_processing.py (``pp.py``),
src/model/functions/cascade_functions.py (``cf.py``).
"""
import torch

from . import ops

NEG = -1e9  # cm.py:8


def cascade_match(feat0, feat1, idx01, idx10, mask0=None, mask1=None, temperature=1.0):
    """cm.py:87-149 (inference): normalise, window masks, sparse correlation both
    directions, masked softmax over the K window candidates, max/argmax, gather.

    feat0 [B,L0,C], feat1 [B,L1,C], idx01 [B,L0,K], idx10 [B,L1,K] int64,
    mask0 [B,L0] / mask1 [B,L1] bool or None.
    Returns dict: conf01 [B,L0,K], next_conf01 [B,L0], next_idx01 [B,L0] (int64,
    index into image-1 tokens), conf10, next_conf10, next_idx10.
    """
    C = feat0.shape[-1]
    f0 = feat0 / C ** 0.5                                                   # cm.py:88
    f1 = feat1 / C ** 0.5
    out = {}
    for tag, (qa, kb, idx, mq, mk) in {'01': (f0, f1, idx01, mask0, mask1),
                                       '10': (f1, f0, idx10, mask1, mask0)}.items():
        sim = ops.score3d(qa, kb, idx) / temperature                       # cm.py:119-120, 140-141
        if mq is not None and mk is not None:
            win = torch.gather(mk, 1, idx.reshape(idx.shape[0], -1)).reshape(idx.shape) & mq.unsqueeze(-1)  # cm.py:108-112
            sim = sim.masked_fill(~win, NEG)                               # cm.py:125
        conf = torch.softmax(sim, dim=2)                                   # cm.py:126
        nconf, arg = torch.max(conf, dim=2)                                # cm.py:128
        nidx = torch.gather(idx, 2, arg.unsqueeze(-1)).squeeze(-1)         # cm.py:129
        out['conf' + tag], out['next_conf' + tag], out['next_idx' + tag] = conf, nconf, nidx
    return out


def nms_mask(conf, h, w, window, test_thr):
    """'maxpool_nms' (pp.py:111-121): keep a pixel iff F.max_pool2d(kernel=window,
    stride 1, pad window//2, -inf padding, return_indices) returns the pixel's own
    flat index, i.e. it is the FIRST maximum of its window in row-major scan order;
    then drop conf <= test_thr.  conf [B,h*w] -> bool [B,h*w].
    """
    B = conf.shape[0]
    c = conf.reshape(B, h, w)
    r = window // 2
    pad = torch.full((B, h + 2 * r, w + 2 * r), float('-inf'), dtype=c.dtype, device=c.device)
    pad[:, r:r + h, r:r + w] = c
    keep = torch.ones(B, h, w, dtype=torch.bool, device=c.device)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            if dy == 0 and dx == 0:
                continue
            nb = pad[:, r + dy:r + dy + h, r + dx:r + dx + w]
            if (dy, dx) < (0, 0):      # scanned before the centre: centre must be strictly greater
                keep &= c > nb
            else:                      # scanned after: a later value replaces only if strictly greater
                keep &= c >= nb
    keep = keep.reshape(B, h * w)
    return keep & ~(conf <= test_thr)


def nearest_upsample(x, h_in, w_in, h_out, w_out):
    """F.interpolate(mode='nearest') of [B,h_in*w_in] to [B,h_out*w_out] (cm.py:201-203):
    src = min(floor(dst * float32(in/out)), in-1)."""
    B = x.shape[0]
    dev = x.device
    sy = torch.tensor(h_in / h_out, dtype=torch.float32, device=dev)
    sx = torch.tensor(w_in / w_out, dtype=torch.float32, device=dev)
    ys = torch.clamp((torch.arange(h_out, dtype=torch.float32, device=dev) * sy).floor().long(), max=h_in - 1)
    xs = torch.clamp((torch.arange(w_out, dtype=torch.float32, device=dev) * sx).floor().long(), max=w_in - 1)
    return x.reshape(B, h_in, w_in)[:, ys][:, :, xs].reshape(B, h_out * w_out)


def valid_extent(pm):
    """cf.py:156-157: valid (unpadded) height/width per sample of a [B,H,W] bool mask."""
    hs = pm.sum(1).max(-1)[0].int()
    ws = pm.sum(-1).max(-1)[0].int()
    return hs, ws


def extract_matches(next_conf01, next_idx01, next_idx10, hw0, hw1, hw0_i, *, test_thr, border_rm,
                    nms_window=None, pre_confs=(), pre_thrs=(), double_check=True,
                    pad_mask0=None, pad_mask1=None, scale0=None, scale1=None):
    """CascadeMatching.get_coarse_match, inference branch (cm.py:170-261, 316-331).

    next_conf01 [B,L0] fp32; next_idx01 [B,L0], next_idx10 [B,L1] int64.
    hw0=(h0,w0), hw1=(h1,w1) grid sizes; hw0_i image size.
    nms_window None -> plain threshold (pp.py:43-44) else maxpool NMS.
    pre_confs: list of (conf [B,hp*wp], hp, wp) of previous stages, with pre_thrs (cm.py:199-206).
    pad_mask0/1 [B,h,w] bool: padded variant of border removal (cf.py:142-172).
    scale0/scale1 [B,2] optional per-sample image scales (cm.py:318-319).
    Returns dict b_ids,i_ids,j_ids (int64), mconf, mkpts0_c, mkpts1_c, mask.
    """
    B, L0 = next_conf01.shape
    h0, w0 = hw0
    h1, w1 = hw1
    if nms_window is None:
        mask = next_conf01 > test_thr
    else:
        mask = nms_mask(next_conf01, h0, w0, nms_window, test_thr)
    for (pc, hp, wp), thr in zip(pre_confs, pre_thrs):
        mask = mask & ~(nearest_upsample(pc, hp, wp, h0, w0) <= thr)

    # border removal on the source grid and on the target coordinate (cf.py:120-172)
    b = border_rm
    if b > 0:
        ys = torch.arange(h0, device=next_conf01.device).reshape(1, h0, 1)
        xs = torch.arange(w0, device=next_conf01.device).reshape(1, 1, w0)
        ty = torch.div(next_idx01, w1, rounding_mode='trunc').reshape(B, h0, w0)
        tx = (next_idx01 % w1).reshape(B, h0, w0)
        if pad_mask0 is None:
            src_bad = (ys < b) | (xs < b) | (ys >= h0 - b) | (xs >= w0 - b)
            tgt_bad = (tx < b) | (tx > w1 - b) | (ty < b) | (ty > h1 - b)
        else:
            h0s, w0s = valid_extent(pad_mask0)
            h1s, w1s = valid_extent(pad_mask1)
            h0s, w0s, h1s, w1s = (t.reshape(B, 1, 1).long() for t in (h0s, w0s, h1s, w1s))
            src_bad = (ys < b) | (xs < b) | (ys >= h0s - b) | (xs >= w0s - b)
            tgt_bad = (tx < b) | (tx > w1s - b) | (ty < b) | (ty > h1s - b)
        mask = mask & ~(src_bad | tgt_bad).reshape(B, L0)

    if double_check:   # cm.py:244-251: mutual nearest neighbour
        back = torch.gather(next_idx10, 1, next_idx01)
        mask = mask & (back == torch.arange(L0, device=next_conf01.device).unsqueeze(0))

    keep = mask            # flags before the fallback (what the CUDA path reports as mask_out)
    if mask.sum() == 0:   # cm.py:254-255
        mask = mask.clone()
        mask[:, 0] = True

    b_ids, i_ids = torch.where(mask)
    j_ids = next_idx01[mask]
    mconf = next_conf01[mask]
    scale = hw0_i[0] / h0                                                  # cm.py:317
    s0 = scale * scale0[b_ids] if scale0 is not None else scale
    s1 = scale * scale1[b_ids] if scale1 is not None else scale
    mk0 = torch.stack([i_ids % w0, torch.div(i_ids, w0, rounding_mode='trunc')], dim=1) * s0
    mk1 = torch.stack([j_ids % w1, torch.div(j_ids, w1, rounding_mode='trunc')], dim=1) * s1
    return {'b_ids': b_ids, 'i_ids': i_ids, 'j_ids': j_ids, 'mconf': mconf,
            'mkpts0_c': mk0, 'mkpts1_c': mk1, 'mask': keep}
