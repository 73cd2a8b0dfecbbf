"""Mirror of cuda_imp/QuadTreeAttention/QuadtreeAttention/functions/quadtree_attention.py (the op-level
API the reference modules call): same names, argument meaning and return layouts, backed by
libcasmtr_b200.so -- forward and backward (casmtr_score5d_bwd / casmtr_value_agg_bwd), so the reference's
un-fused Python modules train on it.  (The FUSED modules of casmtr_b200.modules are inference only.)"""
import torch

from .. import functional as F


class ScoreComputation(torch.autograd.Function):
    """reference :7-19.  query [B,N1,4,H,D], key [B,N2,H,D], index [B,N1,K,H] -> [B,N1,4,K,H]"""

    @staticmethod
    def forward(ctx, query, key, index):
        ctx.save_for_backward(query, key, index)
        return F.score5d(query, key, index)

    @staticmethod
    def backward(ctx, grad_output):
        query, key, index = ctx.saved_tensors
        gq, gk = F.score5d_backward(grad_output.contiguous(), query, key, index)
        return gq, gk, None


score_computation_op = ScoreComputation.apply


class value_aggregation(torch.autograd.Function):
    """reference :25-51.  score/index [B,N,f,K,H], value [B,M,H,D] -> [B,N,f,H,D]"""

    @staticmethod
    def forward(ctx, score, value, index):
        ctx.save_for_backward(score, value, index)
        B, N, f, K, H = score.shape
        out = F.value_agg(score.reshape(B, N * f, K, H), value, index.reshape(B, N * f, K, H))
        return out.reshape(B, N, f, H, value.shape[-1])

    @staticmethod
    def backward(ctx, grad_output):
        score, value, index = ctx.saved_tensors
        B, N, f, K, H = score.shape
        gs, gv = F.value_agg_backward(grad_output.contiguous().reshape(B, N * f, H, value.shape[-1]), score.reshape(B, N * f, K, H),
                                      value, index.reshape(B, N * f, K, H))
        return gs.reshape(B, N, f, K, H), gv, None


value_aggregation_op = value_aggregation.apply
