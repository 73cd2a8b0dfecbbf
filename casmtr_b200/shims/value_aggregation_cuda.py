"""Drop-in for the pybind module built from
cuda_imp/QuadTreeAttention/QuadtreeAttention/src/value_aggregation.cpp:62-65."""
from casmtr_b200 import functional as _F


def value_aggregation_forward(score, value, index, output):
    """Writes output [B,N,H,D] in place (the reference's contract: caller-allocated output)."""
    _F.value_agg(score, value, index, output)


def value_aggregation_backward(grad_output, score, value, index, grad_score, grad_value):
    """ACCUMULATES into the caller's grad_score [B,N,K,H] / grad_value [B,M,H,D] like the reference kernel (which the
    reference wrapper hands zero-filled tensors, functions/quadtree_attention.py:47-48)."""
    gs, gv = _F.value_agg_backward(grad_output, score, value, index)
    grad_score.add_(gs)
    grad_value.add_(gv)
