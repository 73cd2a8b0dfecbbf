"""Drop-in replacements for the reference's QuadTree attention modules
(cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py): same class names,
constructor signatures, forward signatures, return layouts and state-dict keys, so that
``src/model/modules/quadtree_attention.py:6`` can import them instead.  Each forward is ONE call
into libcasmtr_b200.so (fused per pyramid level) instead of ~40 torch ops + 2 extension kernels
per level.  Inference only.
"""
import torch
import torch.nn as nn
import torch.nn.functional as tF

from .. import functional as F


def _f32c(x):
    return x.to(torch.float32).contiguous()


class QTAttA(nn.Module):
    """reference :8-140."""

    def __init__(self, nhead, dim, topks=[32, 32, 32, 32], scale=None, use_dropout=False, attention_dropout=0.1):
        super().__init__()
        self.use_dropout = use_dropout
        self.topks = topks
        self.nhead = nhead
        self.dim = dim

    def forward(self, queries, keys, values, q_mask=None, kv_mask=None):
        """queries/keys/values: pyramids (finest first) of [N,C,H,W] -> message [N, H*W, nhead, dim].
        Masks are accepted and ignored, as in the reference (:101)."""
        return F.qtatt_forward([_f32c(q) for q in queries], [_f32c(k) for k in keys], [_f32c(v) for v in values],
                               self.topks, self.nhead, attn_type='A')


class QTAttB(nn.Module):
    """reference :144-286.  Owns the learned per-level merge weight ``weight`` [scale] (:159)."""

    def __init__(self, nhead, dim, scale, topks=[32, 32, 32, 32], use_dropout=False, attention_dropout=0.1, lepe=False):
        super().__init__()
        self.use_dropout = use_dropout
        self.topks = topks
        self.nhead = nhead
        self.dim = dim
        self.lepe = lepe
        if lepe:  # locally enhanced position encoding (:152-158); stays a cuDNN depth-wise conv
            self.get_vs = nn.ModuleList([
                nn.Conv2d(dim * nhead, dim * nhead, kernel_size=3, stride=1, padding=1, groups=dim * nhead)
                for _ in range(scale)])
        self.register_parameter('weight', nn.Parameter(torch.randn(scale)))

    def forward(self, queries, keys, values, q_mask=None, kv_mask=None, rel_pos=None):
        if rel_pos is not None:
            # dead in every shipped config (RELATIVE_PE False at 1/8; reference transformer.py:210-216 builds
            # an unusable table) -- refuse loudly rather than silently ignore the bias.
            raise NotImplementedError('QTAttB rel_pos is not implemented by casmtr_b200')
        n = len(queries)
        out = F.qtatt_forward([_f32c(q) for q in queries], [_f32c(k) for k in keys], [_f32c(v) for v in values],
                              self.topks, self.nhead, weight=self.weight, attn_type='B')     # soft-max over the WHOLE parameter (:264)
        if self.lepe:    # (m_i + lepe_i) * w_i == m_i * w_i + lepe_i * w_i; the second term is added here (:266-282)
            w = torch.softmax(self.weight, dim=0)
            B, C, H, W = values[0].shape
            for i in range(n):
                lp = self.get_vs[i](values[-(i + 1)])
                if lp.shape[-2:] != (H, W):
                    lp = tF.interpolate(lp, size=(H, W), mode='nearest')      # parent -> children broadcast
                out = out + (lp * w[i]).flatten(2).transpose(1, 2).reshape(B, H * W, self.nhead, C // self.nhead)
        return out


class QTAttGuided(nn.Module):
    """reference :289-389.  Quadtree levels seeded by an external `topk_pos` instead of a dense coarsest level; reachable with
    SELF_ATTN_TYPE='topk' (src/model/modules/quadtree_attention.py:37), which no shipped config selects.

    The reference's merge loop reshapes every level with `queries[-i]` (:385): for i = 0 that is `queries[0]`, the FINEST map's
    height, where half of the current level's height is meant.  With more than one level the shapes no longer fit and the reference
    itself raises; with ONE level it runs, but the wrong height permutes the tokens.  This drop-in supports exactly what the
    reference can run -- a single-level pyramid -- and reproduces its output INCLUDING that permutation (`reference_order=True`, the
    default); `reference_order=False` returns the tokens in raster order, which is what the merge was written to produce."""

    def __init__(self, nhead, dim, scale, topks=[32], use_dropout=False):
        super().__init__()
        self.use_dropout = use_dropout
        self.topks = topks
        self.nhead = nhead
        self.dim = dim
        self.register_parameter('weight', nn.Parameter(torch.randn(scale)))
        self.reference_order = True
        self._perm = {}

    def _reference_permutation(self, h0, w0, device):
        """dst -> src token map of the reference's rearrange 'b (H W) (t1 t2) h d -> b (H t1 W t2) h d' with H = h0 (:385) applied to
        per-cell messages, relative to the raster order of the (h0 x w0) query grid."""
        key = (h0, w0, str(device))
        if key not in self._perm:
            Np = (h0 // 2) * (w0 // 2)
            if Np % h0:
                raise RuntimeError(f"QTAttGuided: the reference's merge cannot reshape {Np} cells with H={h0} (quadtree_attention.py:385)")
            W = Np // h0
            p = torch.arange(Np)
            t = torch.arange(4)
            y, x = p // (w0 // 2), p % (w0 // 2)
            src = ((2 * y[:, None] + t[None] // 2) * w0 + 2 * x[:, None] + t[None] % 2)                     # raster token of (cell, child)
            dst = (((p // W)[:, None] * 2 + t[None] // 2) * W + (p % W)[:, None]) * 2 + t[None] % 2           # where the reference puts it
            perm = torch.empty(h0 * w0, dtype=torch.int64)
            perm[dst.reshape(-1)] = src.reshape(-1)
            self._perm[key] = perm.to(device)
        return self._perm[key]

    def forward(self, queries, keys, values, q_mask=None, kv_mask=None, rel_pos=None, topk_pos=None):
        """queries / keys / values: one-level lists of [N,C,H,W]; topk_pos [2,N,(H/2*W/2),K,nhead] (row, col of the key cells at half
        the key resolution) -> message [N, H*W, nhead, dim]."""
        if len(queries) != 1:
            raise NotImplementedError("QTAttGuided: the reference's merge (quadtree_attention.py:372-385) only runs on a single-level "
                                      'pyramid; multi-level guided attention is not defined by it')
        if rel_pos is not None:
            raise NotImplementedError('QTAttGuided rel_pos is not implemented by casmtr_b200')
        if topk_pos is None:
            raise RuntimeError('QTAttGuided needs topk_pos')
        q = _f32c(queries[0])
        out = F.qtatt_guided_forward(q, _f32c(keys[0]), _f32c(values[0]), topk_pos.to(torch.int64).contiguous(), self.weight, self.nhead)
        if self.reference_order:
            out = out.index_select(1, self._reference_permutation(q.shape[2], q.shape[3], out.device))
        return out


class CascadeQTAttB(nn.Module):
    """reference :392-452.  Window (K = 4k) cross attention of the cascade stages."""

    def __init__(self, nhead, dim, dilated, use_dropout=False):
        super().__init__()
        self.use_dropout = use_dropout
        self.nhead = nhead
        self.dim = dim
        self.dilated = 1 if dilated is None else dilated

    def forward(self, query, key, value, topk_pos, rel_pos):
        """query [N,C,h0,w0], key/value [N,C,h1,w1], topk_pos [N,(h0/2*w0/2),k,2] (row,col), rel_pos None or
        [N,nhead,h0*w0,4k] -> (message [N,h0*w0,C], upsampled_idx [N,h0*w0,4k] int64).
        Extension: topk_pos may be the 2-D next_idx [N,(h0/2*w0/2)] of the previous stage; the 5x5 window expansion
        (reference transformer.py:416-440) then happens inside the kernel."""
        return F.cascade_qtatt_forward(_f32c(query), _f32c(key), _f32c(value), topk_pos.to(torch.int64).contiguous(),
                                       rel_pos, self.nhead, self.dilated)
