"""Drop-ins for the reference's attention layers one level above QTAttB / CascadeQTAttB
(src/model/modules/quadtree_attention.py: QuadtreeAttention :9-99, CascadeQuadtreeAttention :102-171) -- SURVEY.md
section 8f "next" #2.  Same constructor arguments, forward signatures, return values and state-dict keys
(q_proj / k_proj / v_proj are nn.Conv2d 1x1 as in the reference, so reference checkpoints load unchanged).

What changes is the data movement around the projections.  The reference receives tokens [B,N,C], permutes them to
NCHW, runs three 1x1 convolutions, builds the pyramid with avg_pool2d launches (3 maps x (scale-1) levels) and hands
NCHW lists to QTAttB, which permutes everything back to token order.  A 1x1 convolution IS a linear layer on the
tokens, so here the projections run on [B,N,C] directly (cuBLAS, like the reference's `proj`), and the token-major
result goes straight into casmtr_qtatt_tokens_fwd / casmtr_cascade_qtatt_tokens_fwd: the pyramid is built inside in
the layout the kernels gather from, and no NCHW tensor, transpose or pooling launch exists.  Inference only.
"""
import torch
import torch.nn as nn
import torch.nn.functional as tF

from .. import functional as F
from .quadtree_attention import CascadeQTAttB, QTAttA, QTAttB, QTAttGuided


def _trunc_normal_(w, std=0.02):
    return nn.init.trunc_normal_(w, std=std, a=-2.0, b=2.0)


def _init_weights(m):                       # reference :44-66 (timm's trunc_normal_ == torch's)
    if isinstance(m, nn.Linear):
        _trunc_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        _trunc_normal_(m.weight)
        m.init = True
        if m.bias is not None:
            m.bias.data.zero_()


def _project(conv, tokens):
    """1x1 convolution applied as a linear layer on [B,N,C] tokens -> fp32 token-major."""
    return tF.linear(tokens, conv.weight.flatten(1), conv.bias).to(torch.float32).contiguous()


class QuadtreeAttention(nn.Module):
    """reference :9-99."""

    def __init__(self, dim, num_heads, topks, value_branch=False, act=nn.GELU(), qkv_bias=False, qk_scale=None,
                 attn_drop=0.0, proj_drop=0.0, scale=1, attn_type='B'):
        super().__init__()
        assert dim % num_heads == 0, f'dim {dim} should be divided by num_heads {num_heads}.'
        self.dim = dim
        self.num_heads = num_heads
        self.q_proj = nn.Conv2d(dim, dim, kernel_size=1, stride=1, bias=qkv_bias)
        self.k_proj = nn.Conv2d(dim, dim, kernel_size=1, stride=1, bias=qkv_bias)
        self.v_proj = nn.Conv2d(dim, dim, kernel_size=1, stride=1, bias=qkv_bias)
        self.attn_type = attn_type
        if attn_type == 'Guided':
            self.py_att = QTAttGuided(num_heads, dim // num_heads, scale=scale, topks=topks)
        elif attn_type == 'A':
            self.py_att = QTAttA(num_heads, dim // num_heads, scale=scale, topks=topks)
        else:
            self.py_att = QTAttB(num_heads, dim // num_heads, scale=scale, topks=topks)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.scale = scale
        self.apply(_init_weights)

    def forward(self, x, target, H, W, H1=None, W1=None, rel_pos=None, topk_pos=None):
        H1 = H if H1 is None else H1
        W1 = W if W1 is None else W1
        B, N, C = x.shape
        if self.attn_type == 'Guided':        # scale 1 only (see QTAttGuided): NCHW maps, as the guided module consumes them
            if self.scale != 1:
                raise NotImplementedError("QTAttGuided: the reference's merge only runs on a single-level pyramid (scale = 1)")
            q, k, v = (_project(c, t) for c, t in ((self.q_proj, x), (self.k_proj, target), (self.v_proj, target)))
            nchw = lambda t, h, w: t.transpose(1, 2).reshape(B, C, h, w).contiguous()
            msg = self.py_att([nchw(q, H, W)], [nchw(k, H1, W1)], [nchw(v, H1, W1)], rel_pos=rel_pos, topk_pos=topk_pos).reshape(B, -1, C)
            return self.proj_drop(self.proj(msg))
        if rel_pos is not None:
            raise NotImplementedError('QTAttB rel_pos is not implemented by casmtr_b200')
        if getattr(self.py_att, 'lepe', False):
            raise NotImplementedError('lepe needs the NCHW value pyramid: call QTAttB through its own forward')
        q, k, v = _project(self.q_proj, x), _project(self.k_proj, target), _project(self.v_proj, target)
        topks = list(self.py_att.topks)[:self.scale]
        weight = self.py_att.weight if self.attn_type != 'A' else None
        msg = F.qtatt_tokens_forward(q, k, v, (H, W), (H1, W1), topks, self.num_heads, weight=weight,
                                     attn_type='A' if self.attn_type == 'A' else 'B').view(B, -1, C)
        return self.proj_drop(self.proj(msg))


class CascadeQuadtreeAttention(nn.Module):
    """reference :102-171."""

    def __init__(self, dim, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0, scale=2, dilated=1):
        super().__init__()
        assert dim % num_heads == 0, f'dim {dim} should be divided by num_heads {num_heads}.'
        self.dim = dim
        self.num_heads = num_heads
        self.q_proj = nn.Conv2d(dim, dim, kernel_size=1, stride=1, bias=qkv_bias)
        self.k_proj = nn.Conv2d(dim, dim, kernel_size=1, stride=1, bias=qkv_bias)
        self.v_proj = nn.Conv2d(dim, dim, kernel_size=1, stride=1, bias=qkv_bias)
        self.cross_attn = CascadeQTAttB(num_heads, dim // num_heads, dilated=dilated)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.scale = scale
        self.apply(_init_weights)

    def forward(self, x, target, H, W, H1=None, W1=None, idx=None, rel_pos=None):
        H1 = H if H1 is None else H1
        W1 = W if W1 is None else W1
        B, N, C = x.shape
        q, k, v = _project(self.q_proj, x), _project(self.k_proj, target), _project(self.v_proj, target)
        msg, upsampled_idx = F.cascade_qtatt_forward(q, k, v, idx.contiguous(), rel_pos, self.num_heads,
                                                     dilated=self.cross_attn.dilated, hw_q=(H, W), hw_k=(H1, W1))
        return self.proj_drop(self.proj(msg)), upsampled_idx


class CascadeRelativePE(nn.Module):
    """The relative position bias of CascadeFeatureTransformer (reference src/model/modules/transformer.py: tables and LB
    :356-362, get_window_warp_idx :416-440, get_relative_pe :473-509; indoor config only) -- SURVEY.md section 8f "next" #3.
    Same parameter names (`h_pos_bias.weight`, `w_pos_bias.weight`), so the two tables load from a reference checkpoint
    with the `coarse2.` / `coarse3.` prefix stripped.

    get_relative_pe() is the drop-in (same arguments, returns the [B,nhead,HW,4ww] tensor, one kernel instead of ~25 torch
    ops); fused() returns a handle to pass as `rel_pos` to CascadeQuadtreeAttention / CascadeQTAttB, in which case the bias
    is computed inside the attention kernels and the tensor never exists."""

    def __init__(self, nhead, window_size=5, sr_ratio=2):
        super().__init__()
        self.nhead, self.window_size, self.sr_ratio = nhead, window_size, sr_ratio
        self.LB = window_size * 2 if sr_ratio == 2 else window_size * 6          # :357-360
        self.h_pos_bias = nn.Embedding(self.LB * 2 + sr_ratio, nhead)
        self.w_pos_bias = nn.Embedding(self.LB * 2 + sr_ratio, nhead)

    def get_window_warp_idx(self, idx, B, H, W):
        """idx [B,HW] -> ([B,HW,ww,2], None) (:416-440, 'window' propagation)."""
        return F.window_warp_idx(idx.reshape(B, -1).contiguous(), H, W, self.window_size), None

    def fused(self, data, i=0):
        tgt = data['stage_8c']['next_idx_c01'] if i == 0 else data['stage_8c']['next_idx_c10']
        return F.RelativePE(self.w_pos_bias.weight, self.h_pos_bias.weight, self.LB, tgt.contiguous(),
                            data[f'hw{i}_8c'], data[f'hw{1 - i}_8c'][1])

    def get_relative_pe(self, data, H, window_idx, device=None, i=0):
        h, w = data[f'hw{i}_8c'][0], data[f'hw{i}_8c'][1]
        return F.relative_pe(self.fused(data, i), window_idx.contiguous(), (H, w * (H // h)))
