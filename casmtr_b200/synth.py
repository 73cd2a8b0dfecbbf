"""Synthetic hot-path inputs: seeded feature maps, no backbone, no weights.

Used by bench.py, the tests and the golden-vector generator so that every leg
(CUDA path, CPU oracle, reference) sees identical tensors.  CPU generators only
(``torch.Generator`` on CPU is reproducible across devices); callers move the
tensors where they need them.
"""
import torch
import torch.nn.functional as F


def _gen(seed):
    return torch.Generator().manual_seed(int(seed))


def qtatt_inputs(B, C, h, w, levels=3, seed=1234):
    """q/k/v pyramids the way the caller of QTAttB builds them
    (reference src/model/modules/quadtree_attention.py:81-89): finest level
    ~N(0,1), coarser levels by avg_pool2d(2,2).  Returns (queries, keys, values,
    weight) with lists ordered finest -> coarsest and weight ~ N(0,1) [levels]."""
    g = _gen(seed)
    out = []
    for _ in range(3):
        x = torch.randn(B, C, h, w, generator=g)
        pyr = [x]
        for _ in range(levels - 1):
            x = F.avg_pool2d(x, kernel_size=2, stride=2)
            pyr.append(x)
        out.append(pyr)
    weight = torch.randn(levels, generator=g)
    return out[0], out[1], out[2], weight


def window_positions(idx, H, W, window=5):
    """Previous-stage argmax index [B,L] -> window positions [B,L,window^2,2] (row, col)
    at that stage's grid (H, W), rigidly shifted so the whole window stays inside the
    grid.  Same result as the reference's CascadeFeatureTransformer.get_window_warp_idx
    (src/model/modules/transformer.py:416-433) with the 'window' propagation table
    (src/model/modules/propagations.py:12-15); written independently."""
    r = window // 2
    cy = torch.div(idx, W, rounding_mode='trunc').clamp(r, H - 1 - r)
    cx = (idx % W).clamp(r, W - 1 - r)
    off = torch.arange(-r, r + 1, device=idx.device)
    oy, ox = torch.meshgrid(off, off, indexing='ij')
    rows = cy.unsqueeze(-1) + oy.reshape(-1)
    cols = cx.unsqueeze(-1) + ox.reshape(-1)
    return torch.stack([rows, cols], dim=-1)


def cascade_inputs(B, C, h, w, seed=1234, max_shift=4, corrupt=0.1, pad=False, shifts=None):
    """Structured cascade-stage inputs: feat1 is feat0 rolled by a per-pair integer
    shift plus noise, so the window correlation has a real peak (random features
    would threshold every match away).

    Returns dict: feat0, feat1 [B,C,h,w]; next_idx01, next_idx10 [B,(h/2*w/2)] int64
    (previous-stage correspondences, ``corrupt`` fraction randomised); pre_conf01
    [B,h/2*w/2] ~U(0,1); topk_pos01/topk_pos10 [B,h/2*w/2,25,2]; shifts [B,2];
    optional pad masks mask0/mask1 [B,h,w] bool (bottom/right bands).
    shifts: optional [B,2] even integer shifts to use instead of drawing them (the same physical motion at another level)."""
    g = _gen(seed)
    hp, wp = h // 2, w // 2
    feat0 = 3.0 * torch.randn(B, C, h, w, generator=g)
    drawn = torch.randint(-max_shift, max_shift + 1, (B, 2), generator=g) * 2    # even => exact at the previous level
    shifts = drawn if shifts is None else shifts.to(torch.int64)
    feat1 = torch.stack([torch.roll(feat0[b], (int(shifts[b, 0]), int(shifts[b, 1])), (1, 2)) for b in range(B)])
    feat1 = feat1 + 0.3 * torch.randn(B, C, h, w, generator=g)
    py, px = torch.meshgrid(torch.arange(hp), torch.arange(wp), indexing='ij')
    nidx01, nidx10 = [], []
    for b in range(B):
        sy, sx = int(shifts[b, 0]) // 2, int(shifts[b, 1]) // 2
        nidx01.append((((py + sy) % hp) * wp + (px + sx) % wp).reshape(-1))
        nidx10.append((((py - sy) % hp) * wp + (px - sx) % wp).reshape(-1))
    nidx01, nidx10 = torch.stack(nidx01), torch.stack(nidx10)
    for t in (nidx01, nidx10):
        bad = torch.rand(t.shape, generator=g) < corrupt
        t[bad] = torch.randint(0, hp * wp, (int(bad.sum()),), generator=g)
    out = {
        'feat0': feat0, 'feat1': feat1, 'next_idx01': nidx01, 'next_idx10': nidx10,
        'pre_conf01': torch.rand(B, hp * wp, generator=g), 'shifts': shifts,
        'topk_pos01': window_positions(nidx01, hp, wp), 'topk_pos10': window_positions(nidx10, hp, wp),
    }
    if pad:
        m0 = torch.ones(B, h, w, dtype=torch.bool)
        m1 = torch.ones(B, h, w, dtype=torch.bool)
        m0[:, h - max(2, h // 8):] = False
        m1[:, :, w - max(2, w // 8):] = False
        out['mask0'], out['mask1'] = m0, m1
    return out


def fine_inputs(M, WW=25, C=64, seed=1234):
    g = _gen(seed)
    return torch.randn(M, WW, C, generator=g), torch.randn(M, WW, C, generator=g)


def topk_gap(scores, k):
    """Smallest gap across the top-k boundary and between consecutive selected entries
    along the last dim; the tie guard regenerates inputs when it is below 1e-6 relative."""
    s, _ = torch.sort(scores, dim=-1, descending=True)
    return (s[..., :k] - s[..., 1:k + 1]).min()
