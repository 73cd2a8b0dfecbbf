"""Mirror of cuda_imp/QuadTreeAttention/QuadtreeAttention/functions/quadtree_attention.py (the op-level
API the reference modules call): same names, argument meaning and return layouts, backed by
libcasmtr_b200.so.  Inference only -- a backward pass raises."""
import torch

from .. import functional as F


class ScoreComputation(torch.autograd.Function):
    """reference :7-19.  query [B,N1,4,H,D], key [B,N2,H,D], index [B,N1,K,H] -> [B,N1,4,K,H]"""

    @staticmethod
    def forward(ctx, query, key, index):
        return F.score5d(query, key, index)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError('casmtr_b200 implements the inference (forward) path only')


score_computation_op = ScoreComputation.apply


class value_aggregation(torch.autograd.Function):
    """reference :25-38.  score/index [B,N,f,K,H], value [B,M,H,D] -> [B,N,f,H,D]"""

    @staticmethod
    def forward(ctx, score, value, index):
        B, N, f, K, H = score.shape
        out = F.value_agg(score.reshape(B, N * f, K, H), value, index.reshape(B, N * f, K, H))
        return out.reshape(B, N, f, H, value.shape[-1])

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError('casmtr_b200 implements the inference (forward) path only')


value_aggregation_op = value_aggregation.apply
