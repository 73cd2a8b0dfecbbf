"""Tensor-level wrappers over the C ABI: argument checks, output/workspace allocation through
torch's caching allocator, current-stream plumbing.  Everything here requires CUDA tensors --
a CPU tensor raises, nothing falls back.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ExtractDesc, QtattDesc, RelpeDesc, check, lib


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _chk(t, name, dtype=None):
    # same conditions as the reference's CHECK_INPUT (is_cuda + is_contiguous), plus dtype
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name} must be a torch.Tensor')
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be contiguous')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'{name} must be {dtype}, got {t.dtype}')
    return t


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------- launch accounting
def launch_count():
    """Kernels launched by libcasmtr_b200.so in this process so far."""
    return int(lib().casmtr_launch_count())


def profile_enable(on=True):
    """Bracket every kernel launch of the library with a CUDA event pair on its launch stream."""
    check(lib().casmtr_profile_enable(1 if on else 0), 'casmtr_profile_enable')


def set_pdl(on):
    """Programmatic dependent launch of the hot-path kernels on/off for the calling thread (casmtr_set_pdl); returns the
    previous setting."""
    return bool(lib().casmtr_set_pdl(1 if on else 0))


def set_overlap(on):
    """Side-stream overlap of the finer levels' transposes with the coarsest level inside casmtr_qtatt_fwd (casmtr_set_overlap);
    returns the previous setting.  Call it once outside any CUDA-graph capture: it creates the library's side streams."""
    return bool(lib().casmtr_set_overlap(1 if on else 0))


def profile_collect():
    """-> {kind_name: (device_ms, launches)} accumulated since the last collect (synchronises)."""
    ms = (C.c_double * _lib.K_COUNT)()
    n = (C.c_uint64 * _lib.K_COUNT)()
    check(lib().casmtr_profile_collect(ms, n), 'casmtr_profile_collect')
    return {lib().casmtr_kernel_kind_name(i).decode(): (ms[i], int(n[i])) for i in range(_lib.K_COUNT)}


# ------------------------------------------------------------------------------------- op-level
def score5d(query, key, index):
    """[B,N1,4,H,D] x [B,N2,H,D] gathered by index [B,N1,K,H] -> [B,N1,4,K,H]."""
    _chk(query, 'query', torch.float32), _chk(key, 'key', torch.float32), _chk(index, 'index', torch.int64)
    B, N1, F, H, D = query.shape
    if F != 4:
        raise RuntimeError('query must be [B, N1, 4, H, D]')
    N2, K = key.shape[1], index.shape[2]
    out = torch.empty(B, N1, 4, K, H, dtype=torch.float32, device=query.device)
    with torch.cuda.device(query.device):
        check(lib().casmtr_score5d_fwd(_ptr(query), _ptr(key), _ptr(index), _ptr(out), B, N1, N2, H, D, K, _stream(query)),
              'casmtr_score5d_fwd')
    return out


def value_agg(score, value, index, output=None):
    """score/index [B,N,K,H], value [B,M,H,D] -> output [B,N,H,D] (written in place if given)."""
    _chk(score, 'score', torch.float32), _chk(value, 'value', torch.float32), _chk(index, 'index', torch.int64)
    B, N, K, H = score.shape
    M, D = value.shape[1], value.shape[3]
    if output is None:
        output = torch.empty(B, N, H, D, dtype=torch.float32, device=score.device)
    _chk(output, 'output', torch.float32)
    with torch.cuda.device(score.device):
        check(lib().casmtr_value_agg_fwd(_ptr(score), _ptr(value), _ptr(index), _ptr(output), B, N, K, H, M, D, _stream(score)),
              'casmtr_value_agg_fwd')
    return output


def score3d(query, key, index):
    """[B,N1,C] x [B,N2,C] gathered by index [B,N1,K] -> [B,N1,K]."""
    _chk(query, 'query', torch.float32), _chk(key, 'key', torch.float32), _chk(index, 'index', torch.int64)
    B, N1, Cc = query.shape
    N2, K = key.shape[1], index.shape[2]
    out = torch.empty(B, N1, K, dtype=torch.float32, device=query.device)
    with torch.cuda.device(query.device):
        check(lib().casmtr_score3d_fwd(_ptr(query), _ptr(key), _ptr(index), _ptr(out), B, N1, N2, Cc, K, _stream(query)),
              'casmtr_score3d_fwd')
    return out


def score5d_backward(grad_output, query, key, index):
    """-> (grad_query [B,N1,4,H,D], grad_key [B,N2,H,D]) of score5d for grad_output [B,N1,4,K,H]."""
    _chk(grad_output, 'grad_output', torch.float32), _chk(query, 'query', torch.float32), _chk(key, 'key', torch.float32)
    _chk(index, 'index', torch.int64)
    B, N1, _, H, D = query.shape
    N2, K = key.shape[1], index.shape[2]
    if grad_output.shape != (B, N1, 4, K, H):
        raise RuntimeError(f'grad_output must be [B,N1,4,K,H], got {tuple(grad_output.shape)}')
    gq, gk = torch.empty_like(query), torch.empty_like(key)
    with torch.cuda.device(query.device):
        check(lib().casmtr_score5d_bwd(_ptr(grad_output), _ptr(query), _ptr(key), _ptr(index), _ptr(gq), _ptr(gk), B, N1, N2, H, D, K,
                                       _stream(query)), 'casmtr_score5d_bwd')
    return gq, gk


def value_agg_backward(grad_output, score, value, index):
    """-> (grad_score [B,N,K,H], grad_value [B,M,H,D]) of value_agg for grad_output [B,N,H,D]."""
    _chk(grad_output, 'grad_output', torch.float32), _chk(score, 'score', torch.float32), _chk(value, 'value', torch.float32)
    _chk(index, 'index', torch.int64)
    B, N, K, H = score.shape
    M, D = value.shape[1], value.shape[3]
    if grad_output.shape != (B, N, H, D):
        raise RuntimeError(f'grad_output must be [B,N,H,D], got {tuple(grad_output.shape)}')
    gs, gv = torch.empty_like(score), torch.empty_like(value)
    with torch.cuda.device(score.device):
        check(lib().casmtr_value_agg_bwd(_ptr(grad_output), _ptr(score), _ptr(value), _ptr(index), _ptr(gs), _ptr(gv), B, N, K, H, M, D,
                                         _stream(score)), 'casmtr_value_agg_bwd')
    return gs, gv


def score3d_backward(grad_output, query, key, index):
    """-> (grad_query [B,N1,C], grad_key [B,N2,C]) of score3d for grad_output [B,N1,K]."""
    _chk(grad_output, 'grad_output', torch.float32), _chk(query, 'query', torch.float32), _chk(key, 'key', torch.float32)
    _chk(index, 'index', torch.int64)
    B, N1, Cc = query.shape
    N2, K = key.shape[1], index.shape[2]
    if grad_output.shape != (B, N1, K):
        raise RuntimeError(f'grad_output must be [B,N1,K], got {tuple(grad_output.shape)}')
    gq, gk = torch.empty_like(query), torch.empty_like(key)
    with torch.cuda.device(query.device):
        check(lib().casmtr_score3d_bwd(_ptr(grad_output), _ptr(query), _ptr(key), _ptr(index), _ptr(gq), _ptr(gk), B, N1, N2, Cc, K,
                                       _stream(query)), 'casmtr_score3d_bwd')
    return gq, gk


def nchw_to_tokens(x):
    _chk(x, 'x', torch.float32)
    B, Cc = x.shape[:2]
    HW = x[0, 0].numel()
    out = torch.empty(B, HW, Cc, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().casmtr_nchw_to_tokens(_ptr(x), _ptr(out), B, Cc, HW, _stream(x)), 'casmtr_nchw_to_tokens')
    return out


# ------------------------------------------------------------------------------------- fused QuadTree attention
def qtatt_forward(queries, keys, values, topks, nhead, weight=None, attn_type='B', return_topk=False, flags=0, concurrent_calls=0):
    """Fused QTAttA / QTAttB forward.  queries/keys/values: lists finest->coarsest of [B,C,H,W] fp32.
    Returns message [B, L_finest, nhead, D] (and, with return_topk, per-level lists of top-k key
    indices [B,L,k,nhead] int64 and scores, processing order, last level omitted).
    weight: the whole QTAttB.weight parameter (its soft-max runs over all entries, level i uses entry i).
    flags: _lib.QT_* bits; concurrent_calls: launch-geometry hint (casmtr_qtatt_desc)."""
    n = len(queries)
    if not (len(keys) == n and len(values) == n and 1 <= n <= _lib.MAX_LEVELS):
        raise RuntimeError(f'need 1..{_lib.MAX_LEVELS} pyramid levels, got {n}')
    if len(topks) < n:
        raise RuntimeError(f'topks {topks} shorter than the pyramid ({n} levels)')
    qs = [_chk(q, f'queries[{i}]', torch.float32) for i, q in enumerate(queries)]
    ks = [_chk(k, f'keys[{i}]', torch.float32) for i, k in enumerate(keys)]
    vs = [_chk(v, f'values[{i}]', torch.float32) for i, v in enumerate(values)]
    dev = qs[0].device
    B, Cc = qs[0].shape[:2]
    d = QtattDesc()
    d.B, d.nhead, d.D, d.levels, d.type = B, nhead, Cc // nhead, n, 1 if attn_type == 'A' else 0
    d.flags, d.concurrent_calls = int(flags), int(concurrent_calls)
    for l in range(n):
        if ks[l].shape != vs[l].shape or qs[l].shape[:2] != (B, Cc) or ks[l].shape[:2] != (B, Cc):
            raise RuntimeError(f'inconsistent shapes at level {l}')
        d.qh[l], d.qw[l] = qs[l].shape[2:]
        d.kh[l], d.kw[l] = ks[l].shape[2:]
        d.topks[l] = int(topks[l])
    if d.type == 0:
        if weight is None:
            raise RuntimeError('QTAttB needs the level weight parameter')
        weight = _chk(weight.detach().to(torch.float32).contiguous(), 'weight', torch.float32)
        if weight.numel() < n:
            raise RuntimeError('weight shorter than the pyramid')
        d.weight_len = weight.numel()
    L0 = d.qh[0] * d.qw[0]
    out = torch.empty(B, L0, nhead, Cc // nhead, dtype=torch.float32, device=dev)
    arr = C.c_void_p * n
    qa, ka, va = arr(*[q.data_ptr() for q in qs]), arr(*[k.data_ptr() for k in ks]), arr(*[v.data_ptr() for v in vs])
    tk_idx, tk_sc, ia, sa = [], [], None, None
    if return_topk:
        ia, sa = arr(), arr()
        for i in range(n - 1 if n > 1 else 1):
            l = n - 1 - i
            shp = (B, d.qh[l] * d.qw[l], int(topks[i]), nhead)
            tk_idx.append(torch.empty(shp, dtype=torch.int64, device=dev))
            tk_sc.append(torch.empty(shp, dtype=torch.float32, device=dev))
            ia[i], sa[i] = tk_idx[-1].data_ptr(), tk_sc[-1].data_ptr()
    with torch.cuda.device(dev):
        nbytes = lib().casmtr_qtatt_workspace_bytes(C.byref(d))
        if nbytes == 0:
            check(-1, 'casmtr_qtatt_workspace_bytes')
        ws = _workspace(nbytes, dev)
        check(lib().casmtr_qtatt_fwd(C.byref(d), qa, ka, va, _ptr(weight if d.type == 0 else None), _ptr(out),
                                     ia, sa, _ptr(ws), ws.numel(), _stream(out)), 'casmtr_qtatt_fwd')
    if return_topk:
        return out, tk_idx, tk_sc
    return out


def qtatt_tokens_forward(q0, k0, v0, hw_q, hw_k, topks, nhead, weight=None, attn_type='B', return_topk=False, flags=0, concurrent_calls=0):
    """QTAttA / QTAttB from the finest level only, token-major: q0 [B, hq*wq, C], k0 / v0 [B, hk*wk, C] fp32; the
    avg-pool pyramid of len(topks) levels is built inside (casmtr_qtatt_tokens_fwd).  Returns as qtatt_forward."""
    _chk(q0, 'q0', torch.float32), _chk(k0, 'k0', torch.float32), _chk(v0, 'v0', torch.float32)
    n = len(topks)
    if not 1 <= n <= _lib.MAX_LEVELS:
        raise RuntimeError(f'need 1..{_lib.MAX_LEVELS} pyramid levels, got {n}')
    B, Lq, Cc = q0.shape
    (hq, wq), (hk, wk) = hw_q, hw_k
    if Lq != hq * wq or k0.shape != (B, hk * wk, Cc) or v0.shape != k0.shape:
        raise RuntimeError(f'token maps {tuple(q0.shape)} / {tuple(k0.shape)} / {tuple(v0.shape)} do not match grids {hw_q} / {hw_k}')
    if any(s % (1 << (n - 1)) for s in (hq, wq, hk, wk)):
        raise RuntimeError(f'grids {hw_q} / {hw_k} are not divisible by 2^{n - 1}')
    dev = q0.device
    d = QtattDesc()
    d.B, d.nhead, d.D, d.levels, d.type = B, nhead, Cc // nhead, n, 1 if attn_type == 'A' else 0
    d.flags, d.concurrent_calls = int(flags), int(concurrent_calls)
    for l in range(n):
        d.qh[l], d.qw[l], d.kh[l], d.kw[l], d.topks[l] = hq >> l, wq >> l, hk >> l, wk >> l, int(topks[l])
    if d.type == 0:
        if weight is None:
            raise RuntimeError('QTAttB needs the level weight parameter')
        weight = _chk(weight.detach().to(torch.float32).contiguous(), 'weight', torch.float32)
        if weight.numel() < n:
            raise RuntimeError('weight shorter than the pyramid')
        d.weight_len = weight.numel()
    out = torch.empty(B, Lq, nhead, Cc // nhead, dtype=torch.float32, device=dev)
    arr = C.c_void_p * n
    tk_idx, tk_sc, ia, sa = [], [], None, None
    if return_topk:
        ia, sa = arr(), arr()
        for i in range(n - 1 if n > 1 else 1):
            l = n - 1 - i
            shp = (B, d.qh[l] * d.qw[l], int(topks[i]), nhead)
            tk_idx.append(torch.empty(shp, dtype=torch.int64, device=dev))
            tk_sc.append(torch.empty(shp, dtype=torch.float32, device=dev))
            ia[i], sa[i] = tk_idx[-1].data_ptr(), tk_sc[-1].data_ptr()
    with torch.cuda.device(dev):
        nbytes = lib().casmtr_qtatt_workspace_bytes(C.byref(d))
        if nbytes == 0:
            check(-1, 'casmtr_qtatt_workspace_bytes')
        ws = _workspace(nbytes, dev)
        check(lib().casmtr_qtatt_tokens_fwd(C.byref(d), _ptr(q0), _ptr(k0), _ptr(v0), _ptr(weight if d.type == 0 else None), _ptr(out),
                                            ia, sa, _ptr(ws), ws.numel(), _stream(out)), 'casmtr_qtatt_tokens_fwd')
    if return_topk:
        return out, tk_idx, tk_sc
    return out


def qtatt_guided_forward(query, key, value, topk_pos, weight, nhead):
    """One guided quadtree level (QTAttGuided, single-level pyramid): query [B,C,h0,w0], key / value [B,C,h1,w1], topk_pos
    [2,B,(h0/2)*(w0/2),K,nhead] int64 (row, col of the K key cells per query cell and head on the (h1/2 x w1/2) grid), weight = the
    module's raw level weights.  -> message [B, h0*w0, nhead, D] in raster order, scaled by softmax(weight)[0]."""
    _chk(query, 'query', torch.float32), _chk(key, 'key', torch.float32), _chk(value, 'value', torch.float32)
    _chk(topk_pos, 'topk_pos', torch.int64)
    B, Cc, h0, w0 = query.shape
    h1, w1 = key.shape[2:]
    if topk_pos.dim() != 5 or topk_pos.shape[:3] != (2, B, (h0 // 2) * (w0 // 2)) or topk_pos.shape[4] != nhead:
        raise RuntimeError(f'topk_pos must be [2,B,(h0/2)*(w0/2),K,nhead], got {tuple(topk_pos.shape)}')
    K = topk_pos.shape[3]
    weight = _chk(weight.detach().to(torch.float32).contiguous(), 'weight', torch.float32)
    dev = query.device
    out = torch.empty(B, h0 * w0, nhead, Cc // nhead, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib().casmtr_qtatt_guided_workspace_bytes(B, Cc, nhead, h0, w0, h1, w1, K)
        if nbytes == 0:
            check(-1, 'casmtr_qtatt_guided_workspace_bytes')
        ws = _workspace(nbytes, dev)
        check(lib().casmtr_qtatt_guided_fwd(_ptr(query), _ptr(key), _ptr(value), _ptr(topk_pos), _ptr(weight), weight.numel(), _ptr(out),
                                            B, nhead, Cc // nhead, h0, w0, h1, w1, K, _ptr(ws), ws.numel(), _stream(out)), 'casmtr_qtatt_guided_fwd')
    return out


def window_warp_idx(next_idx, H, W, window=5):
    """next_idx [B,L] int64 on an H x W grid -> [B,L,window^2,2] (row, col) of the border-shifted window around each match
    (CascadeFeatureTransformer.get_window_warp_idx, reference transformer.py:416-440)."""
    _chk(next_idx, 'next_idx', torch.int64)
    B, L = next_idx.shape
    pos = torch.empty(B, L, window * window, 2, dtype=torch.int64, device=next_idx.device)
    with torch.cuda.device(next_idx.device):
        check(lib().casmtr_window_idx_fwd(_ptr(next_idx), _ptr(pos), B, L, int(H), int(W), int(window), _stream(next_idx)), 'casmtr_window_idx_fwd')
    return pos


class RelativePE:
    """What CascadeFeatureTransformer.get_relative_pe (reference transformer.py:473-509) reads besides the window: the two
    embedding tables (w_pos_bias.weight / h_pos_bias.weight, [2*LB + sr_ratio, nhead]), LB, the query image's 1/8 match
    tgt_idx [B,h8*w8] (data['stage_8c']['next_idx_c01' / 'next_idx_c10']) and the two 1/8 grids.  Pass it as `rel_pos` to
    cascade_qtatt_forward and the bias is computed inside the attention kernels; relative_pe() materialises the tensor."""

    def __init__(self, w_table, h_table, LB, tgt_idx, hw8, w8_other):
        self.w_table = _chk(w_table.detach().to(torch.float32).contiguous(), 'w_table', torch.float32)
        self.h_table = _chk(h_table.detach().to(torch.float32).contiguous(), 'h_table', torch.float32)
        self.tgt_idx = _chk(tgt_idx, 'tgt_idx', torch.int64)
        if self.w_table.dim() != 2 or self.w_table.shape != self.h_table.shape:
            raise RuntimeError('w_table / h_table must both be [n_emb, nhead]')
        self.LB, self.hw8, self.w8_other = int(LB), (int(hw8[0]), int(hw8[1])), int(w8_other)
        if self.tgt_idx.dim() != 2 or self.tgt_idx.shape[1] != self.hw8[0] * self.hw8[1]:
            raise RuntimeError(f'tgt_idx must be [B, h8*w8], got {tuple(self.tgt_idx.shape)}')

    def desc(self, nhead, B):
        if self.w_table.shape[1] != nhead or self.tgt_idx.shape[0] != B:
            raise RuntimeError('relative PE tables / tgt_idx do not match nhead / batch')
        d = RelpeDesc()
        d.w_table, d.h_table, d.tgt_idx = self.w_table.data_ptr(), self.h_table.data_ptr(), self.tgt_idx.data_ptr()
        d.n_emb, d.LB = self.w_table.shape[0], self.LB
        d.h8, d.w8, d.w8_other = self.hw8[0], self.hw8[1], self.w8_other
        return d


def relative_pe(pe, window_pos, hw):
    """get_relative_pe as a tensor: window_pos [B,(H/2)*(W/2),k,2] int64 (get_window_warp_idx output) -> [B,nhead,H*W,4k] fp32."""
    _chk(window_pos, 'window_pos', torch.int64)
    H, W = int(hw[0]), int(hw[1])
    B, Np, k, two = window_pos.shape
    if Np != (H // 2) * (W // 2) or two != 2:
        raise RuntimeError(f'window_pos must be [B,(H/2)*(W/2),k,2], got {tuple(window_pos.shape)}')
    nhead = pe.w_table.shape[1]
    out = torch.empty(B, nhead, H * W, 4 * k, dtype=torch.float32, device=window_pos.device)
    d = pe.desc(nhead, B)
    with torch.cuda.device(window_pos.device):
        check(lib().casmtr_relative_pe_fwd(C.byref(d), _ptr(window_pos), _ptr(out), B, nhead, H, W, k, _stream(out)), 'casmtr_relative_pe_fwd')
    return out


def cascade_qtatt_forward(query, key, value, topk_pos, rel_pos, nhead, dilated=1, need_idx=True, hw_q=None, hw_k=None, window=5):
    """Fused CascadeQTAttB forward -> (message [B,h0*w0,C], upsampled_idx [B,h0*w0,4k] int64 or None).
    query / key / value: NCHW maps, or - with hw_q / hw_k given - token-major [B, h*w, C] (no transposes).
    topk_pos: [B,(h0/2)*(w0/2),k,2] window positions, or the 2-D next_idx [B,(h0/2)*(w0/2)] they are derived from
    (the window x window expansion then happens inside the kernels; dilated must be 1).
    rel_pos: None, the reference's [B,nhead,h0*w0,4k] tensor, or a RelativePE (bias computed inside the kernels)."""
    _chk(query, 'query', torch.float32), _chk(key, 'key', torch.float32), _chk(value, 'value', torch.float32)
    _chk(topk_pos, 'topk_pos', torch.int64)
    from_idx = topk_pos.dim() == 2
    tokens = hw_q is not None
    if tokens:
        (h0, w0), (h1, w1) = hw_q, hw_k if hw_k is not None else hw_q
        B, _, Cc = query.shape
        if query.shape != (B, h0 * w0, Cc) or key.shape != (B, h1 * w1, Cc) or value.shape != key.shape:
            raise RuntimeError('token-major query/key/value do not match hw_q / hw_k')
    else:
        B, Cc, h0, w0 = query.shape
        h1, w1 = key.shape[2:]
    k = window * window if from_idx else topk_pos.shape[2]
    if from_idx:
        if topk_pos.shape != (B, (h0 // 2) * (w0 // 2)) or (dilated or 1) != 1:
            raise RuntimeError(f'next_idx must be [B,(h0/2)*(w0/2)] (got {tuple(topk_pos.shape)}) and dilated 1')
    elif topk_pos.shape != (B, (h0 // 2) * (w0 // 2), k, 2):
        raise RuntimeError(f'topk_pos must be [B,(h0/2)*(w0/2),k,2], got {tuple(topk_pos.shape)}')
    fused_pe = rel_pos if isinstance(rel_pos, RelativePE) else None
    if fused_pe is not None:
        rel_pos = None
        if (dilated or 1) != 1:
            raise RuntimeError('the fused relative PE needs dilated == 1')
    if rel_pos is not None:
        rel_pos = _chk(rel_pos.to(torch.float32).contiguous(), 'rel_pos', torch.float32)
        if rel_pos.numel() != B * nhead * h0 * w0 * 4 * k:
            raise RuntimeError('rel_pos must hold B*nhead*h0*w0*4k values')
    dev = query.device
    msg = torch.empty(B, h0 * w0, Cc, dtype=torch.float32, device=dev)
    up = torch.empty(B, h0 * w0, 4 * k, dtype=torch.int64, device=dev) if need_idx else None
    with torch.cuda.device(dev):
        nbytes = lib().casmtr_cascade_qtatt_workspace_bytes(B, Cc, h0, w0, h1, w1)
        ws = _workspace(nbytes, dev)
        if fused_pe is not None:
            d = fused_pe.desc(nhead, B)
            check(lib().casmtr_cascade_qtatt_relpe_fwd(_ptr(query), _ptr(key), _ptr(value), _ptr(None if from_idx else topk_pos),
                                                       _ptr(topk_pos if from_idx else None), int(window), C.byref(d), _ptr(msg), _ptr(up),
                                                       B, nhead, Cc // nhead, h0, w0, h1, w1, k, 1 if tokens else 0,
                                                       _ptr(ws), ws.numel(), _stream(msg)), 'casmtr_cascade_qtatt_relpe_fwd')
            return msg, up
        if from_idx:
            check(lib().casmtr_cascade_qtatt_window_fwd(_ptr(query), _ptr(key), _ptr(value), _ptr(topk_pos), int(window), _ptr(rel_pos), _ptr(msg), _ptr(up),
                                                        B, nhead, Cc // nhead, h0, w0, h1, w1, 1 if tokens else 0,
                                                        _ptr(ws), ws.numel(), _stream(msg)), 'casmtr_cascade_qtatt_window_fwd')
            return msg, up
        fn = lib().casmtr_cascade_qtatt_tokens_fwd if tokens else lib().casmtr_cascade_qtatt_fwd
        check(fn(_ptr(query), _ptr(key), _ptr(value), _ptr(topk_pos), _ptr(rel_pos), _ptr(msg), _ptr(up),
                 B, nhead, Cc // nhead, h0, w0, h1, w1, k, 1 if dilated is None else int(dilated),
                 _ptr(ws), ws.numel(), _stream(msg)), 'casmtr_cascade_qtatt_fwd')
    return msg, up


# ------------------------------------------------------------------------------------- cascade matching
def cascade_match_forward(feat0, feat1, idx01, idx10, mask0=None, mask1=None, temperature=1.0, need_conf=True,
                          need_conf10=None, w0=0, w1=0):
    """Fused sparse correlation + softmax + argmax, both directions.  Returns dict with conf01/conf10
    (None unless need_conf / need_conf10; need_conf10 defaults to need_conf), next_conf01/10 [B,L] fp32,
    next_idx01/10 [B,L] int64.  w0 / w1: token-grid row lengths (0 = unknown); they only select the faster
    sibling-sharing kernel, results are identical."""
    if need_conf10 is None:
        need_conf10 = need_conf
    _chk(feat0, 'feat0', torch.float32), _chk(feat1, 'feat1', torch.float32)
    _chk(idx01, 'idx01', torch.int64), _chk(idx10, 'idx10', torch.int64)
    B, L0, Cc = feat0.shape
    L1, K = feat1.shape[1], idx01.shape[2]
    dev = feat0.device
    m0 = m1 = None
    if mask0 is not None and mask1 is not None:
        m0 = _chk(mask0.reshape(B, L0).to(torch.uint8).contiguous(), 'mask0', torch.uint8)
        m1 = _chk(mask1.reshape(B, L1).to(torch.uint8).contiguous(), 'mask1', torch.uint8)
    o = {
        'conf01': torch.empty(B, L0, K, dtype=torch.float32, device=dev) if need_conf else None,
        'conf10': torch.empty(B, L1, K, dtype=torch.float32, device=dev) if need_conf10 else None,
        'next_conf01': torch.empty(B, L0, dtype=torch.float32, device=dev),
        'next_conf10': torch.empty(B, L1, dtype=torch.float32, device=dev),
        'next_idx01': torch.empty(B, L0, dtype=torch.int64, device=dev),
        'next_idx10': torch.empty(B, L1, dtype=torch.int64, device=dev),
    }
    with torch.cuda.device(dev):
        ws = _workspace(lib().casmtr_cascade_match_workspace_bytes(B, L0, L1), dev)
        check(lib().casmtr_cascade_match_fwd(_ptr(feat0), _ptr(feat1), _ptr(idx01), _ptr(idx10), _ptr(m0), _ptr(m1),
                                             float(temperature), _ptr(o['conf01']), _ptr(o['next_conf01']), _ptr(o['next_idx01']),
                                             _ptr(o['conf10']), _ptr(o['next_conf10']), _ptr(o['next_idx10']),
                                             B, L0, L1, Cc, K, int(w0), int(w1), _ptr(ws), ws.numel(), _stream(feat0)),
              'casmtr_cascade_match_fwd')
    return o


def coarse_match_forward(feat0, feat1, temperature=0.1, mask0=None, mask1=None, mutual=False):
    """Dense dual-softmax statistics (tcgen05): feat0 [B,L0,C], feat1 [B,L1,C] -> dict next_conf01 [B,L0], next_idx01 [B,L0]
    (int64), next_conf10 [B,L1], next_idx10 [B,L1]; the L0 x L1 similarity matrix is never materialised.
    mask0 [B,L0] / mask1 [B,L1] (bool, both or neither): padding masks as in the reference's masked_fill_(-1e9) (padded rows: uniform soft-max, i.e. 1 / columns and index 0).
    mutual=True adds mconf_row [B,L0] (row maximum of conf = softmax_i * softmax_j), midx_row [B,L0] and midx_col [B,L1] (its row /
    column arg-maxima, int64): what get_coarse_match's mutual-nearest-neighbour test needs (casmtr_coarse_match_mutual_fwd)."""
    _chk(feat0, 'feat0', torch.float32), _chk(feat1, 'feat1', torch.float32)
    B, L0, Cc = feat0.shape
    L1 = feat1.shape[1]
    dev = feat0.device
    if (mask0 is None) != (mask1 is None):
        raise RuntimeError('give both masks or neither')
    if mask0 is not None:
        if mask0.shape != (B, L0) or mask1.shape != (B, L1):
            raise RuntimeError(f'masks must be [B,L0] / [B,L1], got {tuple(mask0.shape)} / {tuple(mask1.shape)}')
        mask0 = _chk((mask0 != 0).to(torch.uint8).contiguous(), 'mask0', torch.uint8)
        mask1 = _chk((mask1 != 0).to(torch.uint8).contiguous(), 'mask1', torch.uint8)
    o = {'next_conf01': torch.empty(B, L0, dtype=torch.float32, device=dev), 'next_idx01': torch.empty(B, L0, dtype=torch.int64, device=dev),
         'next_conf10': torch.empty(B, L1, dtype=torch.float32, device=dev), 'next_idx10': torch.empty(B, L1, dtype=torch.int64, device=dev)}
    with torch.cuda.device(dev):
        nbytes = lib().casmtr_coarse_match_workspace_bytes(B, L0, L1, Cc)
        ws = _workspace(nbytes, dev)
        if mutual:
            o.update({'mconf_row': torch.empty(B, L0, dtype=torch.float32, device=dev), 'midx_row': torch.empty(B, L0, dtype=torch.int64, device=dev),
                      'midx_col': torch.empty(B, L1, dtype=torch.int64, device=dev)})
            check(lib().casmtr_coarse_match_mutual_fwd(_ptr(feat0), _ptr(feat1), _ptr(mask0), _ptr(mask1), float(temperature),
                                                       _ptr(o['next_conf01']), _ptr(o['next_idx01']), _ptr(o['next_conf10']), _ptr(o['next_idx10']),
                                                       _ptr(o['mconf_row']), _ptr(o['midx_row']), _ptr(o['midx_col']), B, L0, L1, Cc,
                                                       _ptr(ws), ws.numel(), _stream(feat0)), 'casmtr_coarse_match_mutual_fwd')
            return o
        check(lib().casmtr_coarse_match_masked_fwd(_ptr(feat0), _ptr(feat1), _ptr(mask0), _ptr(mask1), float(temperature),
                                                   _ptr(o['next_conf01']), _ptr(o['next_idx01']),
                                                   _ptr(o['next_conf10']), _ptr(o['next_idx10']), B, L0, L1, Cc, _ptr(ws), ws.numel(),
                                                   _stream(feat0)), 'casmtr_coarse_match_masked_fwd')
    return o


def match_extract(next_conf01, next_idx01, next_idx10, hw0, hw1, hw0_i, *, test_thr, border_rm, nms_window=None,
                  pre_confs=(), pre_thrs=(), double_check=True, pad_mask0=None, pad_mask1=None,
                  scale0=None, scale1=None, defer=False, coarse_mode=False):
    """NMS / thresholds / border / mutual check / ordered compaction.  Same keyword surface as
    oracle.cascade.extract_matches.  One host sync (the match count), like the reference's torch.where.
    defer=True skips that sync (CUDA-graph capture): the result holds full-capacity buffers plus the device-side
    'count'; slice them with trim_matches() once the count may be read."""
    _chk(next_conf01, 'next_conf01', torch.float32), _chk(next_idx01, 'next_idx01', torch.int64), _chk(next_idx10, 'next_idx10', torch.int64)
    B, L0 = next_conf01.shape
    dev = next_conf01.device
    d = ExtractDesc()
    d.B, (d.h0, d.w0), (d.h1, d.w1) = B, hw0, hw1
    d.nms_window = 0 if nms_window is None else int(nms_window)
    d.test_thr, d.border_rm, d.double_check = float(test_thr), int(border_rm), int(bool(double_check))
    d.coarse_mode = int(bool(coarse_mode))        # CoarseMatching.get_coarse_match: symmetric target border, no empty-list fallback
    keep = []
    d.n_pre = len(pre_confs)
    if d.n_pre > 2 or len(pre_thrs) < d.n_pre:
        raise RuntimeError('at most 2 previous-stage gates, each with a threshold')
    for s, (pc, hp, wp) in enumerate(pre_confs):
        pc = _chk(pc.detach().to(torch.float32).contiguous(), 'pre_conf', torch.float32)
        keep.append(pc)
        d.pre_conf[s], d.pre_h[s], d.pre_w[s], d.pre_thr[s] = pc.data_ptr(), hp, wp, float(pre_thrs[s])
    if pad_mask0 is not None and pad_mask1 is not None:
        pm0 = _chk(pad_mask0.to(torch.uint8).contiguous(), 'pad_mask0', torch.uint8)
        pm1 = _chk(pad_mask1.to(torch.uint8).contiguous(), 'pad_mask1', torch.uint8)
        keep += [pm0, pm1]
        d.pad_mask0, d.pad_mask1 = pm0.data_ptr(), pm1.data_ptr()
    d.scale = float(hw0_i[0] / hw0[0])
    if scale0 is not None:
        s0 = _chk(scale0.to(torch.float32).contiguous(), 'scale0', torch.float32)
        keep.append(s0)
        d.scale0 = s0.data_ptr()
    if scale1 is not None:
        s1 = _chk(scale1.to(torch.float32).contiguous(), 'scale1', torch.float32)
        keep.append(s1)
        d.scale1 = s1.data_ptr()
    cap = max(B * L0, B)
    mask = torch.empty(B, L0, dtype=torch.uint8, device=dev)
    ids = torch.empty(3, cap, dtype=torch.int64, device=dev)
    mconf = torch.empty(cap, dtype=torch.float32, device=dev)
    mk = torch.empty(2, cap, 2, dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib().casmtr_match_extract_workspace_bytes(C.byref(d))
        if nbytes == 0:
            check(-1, 'casmtr_match_extract_workspace_bytes')
        ws = _workspace(nbytes, dev)
        check(lib().casmtr_match_extract(C.byref(d), _ptr(next_conf01), _ptr(next_idx01), _ptr(next_idx10), _ptr(mask),
                                         _ptr(ids[0]), _ptr(ids[1]), _ptr(ids[2]), _ptr(mconf), _ptr(mk[0]), _ptr(mk[1]),
                                         cap, _ptr(count), _ptr(ws), ws.numel(), _stream(mask)), 'casmtr_match_extract')
    full = {'b_ids': ids[0], 'i_ids': ids[1], 'j_ids': ids[2], 'mconf': mconf, 'mkpts0_c': mk[0], 'mkpts1_c': mk[1],
            'mask': mask, 'count': count, '_keep': keep}
    return full if defer else trim_matches(full)


def trim_matches(full):
    """Reads the match count (host sync) and slices the deferred result of match_extract."""
    M = min(int(full['count'].item()), full['b_ids'].shape[0])
    out = {k: full[k][:M] for k in ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0_c', 'mkpts1_c')}
    out['mask'] = full['mask'].bool()
    return out


def pack_matches(m, pair_offset, cap):
    """dict(b_ids,i_ids,j_ids,mconf,mkpts0,mkpts1) of M matches -> uint8 [cap+1, 44] block (row 0 = count), one kernel.
    If m carries 'count' (int32 device tensor of a deferred extraction) the arrays are capacity-sized and the count is read
    on the device: no host synchronisation."""
    count = m.get('count')
    M = int(m['b_ids'].shape[0])
    if count is None and M > cap:
        raise RuntimeError(f'pack_matches: {M} matches exceed the static capacity {cap}')
    dev = m['b_ids'].device
    out = torch.empty(cap + 1, 44, dtype=torch.uint8, device=dev)      # always cap + 1 rows: every rank's block has the same size
    t = [_chk(m[k].contiguous(), k, dt) for k, dt in (('b_ids', torch.int64), ('i_ids', torch.int64), ('j_ids', torch.int64),
                                                      ('mconf', torch.float32), ('mkpts0', torch.float32), ('mkpts1', torch.float32))]
    with torch.cuda.device(dev):
        if count is not None:
            _chk(count, 'count', torch.int32)
            check(lib().casmtr_pack_matches_dev(*[_ptr(x) for x in t], _ptr(count), int(pair_offset), int(min(cap, M)), _ptr(out), _stream(out)),
                  'casmtr_pack_matches_dev')
        else:
            check(lib().casmtr_pack_matches(*[_ptr(x) for x in t], M, int(pair_offset), int(cap), _ptr(out), _stream(out)),
                  'casmtr_pack_matches')
    return out


def fine_window_gather(feat, b_ids, ids, wc, stride, W=5):
    """feat [B,C,Hf,Wf] fp32, b_ids / ids [M] int64 (ids on a coarse grid of width wc, fine = coarse * stride)
    -> windows [M, W*W, C]: F.unfold(feat, W, stride, padding=W//2) selected at (b_ids, ids), without the unfold."""
    _chk(feat, 'feat', torch.float32), _chk(b_ids, 'b_ids', torch.int64), _chk(ids, 'ids', torch.int64)
    B, Cc, Hf, Wf = feat.shape
    M = b_ids.shape[0]
    out = torch.empty(M, W * W, Cc, dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        check(lib().casmtr_fine_window_gather(_ptr(feat), _ptr(b_ids), _ptr(ids), _ptr(out), M, Cc, Hf, Wf, int(wc), int(stride), int(W),
                                              _stream(feat)), 'casmtr_fine_window_gather')
    return out


def fine_match_forward(feat_f0, feat_f1, mkpts1_c, scale, scale1_b=None, b_ids=None, count=None):
    """-> (expec_f [M,3], mkpts1_f [M,2]).  count: optional int32 device tensor -- only rows [0, count) are computed (the
    inputs are then capacity-sized buffers of a deferred extraction; rows past the count are left uninitialised)."""
    _chk(feat_f0, 'feat_f0', torch.float32), _chk(feat_f1, 'feat_f1', torch.float32)
    M, WW, Cc = feat_f0.shape
    dev = feat_f0.device
    mk = _chk(mkpts1_c.to(torch.float32).contiguous(), 'mkpts1_c', torch.float32)
    expec = torch.empty(M, 3, dtype=torch.float32, device=dev)
    out = torch.empty(M, 2, dtype=torch.float32, device=dev)
    s1 = bi = None
    if scale1_b is not None:
        s1 = _chk(scale1_b.to(torch.float32).contiguous(), 'scale1', torch.float32)
        bi = _chk(b_ids.contiguous(), 'b_ids', torch.int64)
    with torch.cuda.device(dev):
        if count is not None:
            _chk(count, 'count', torch.int32)
            check(lib().casmtr_fine_match_dev_fwd(_ptr(feat_f0), _ptr(feat_f1), _ptr(mk), _ptr(s1), _ptr(bi), float(scale),
                                                  _ptr(expec), _ptr(out), _ptr(count), M, WW, Cc, _stream(expec)), 'casmtr_fine_match_dev_fwd')
        else:
            check(lib().casmtr_fine_match_fwd(_ptr(feat_f0), _ptr(feat_f1), _ptr(mk), _ptr(s1), _ptr(bi), float(scale),
                                              _ptr(expec), _ptr(out), M, WW, Cc, _stream(expec)), 'casmtr_fine_match_fwd')
    return expec, out
