"""Drop-in for the pybind module built from
cuda_imp/QuadTreeAttention/QuadtreeAttention/src/value_aggregation.cpp:62-65."""
from casmtr_b200 import functional as _F


def value_aggregation_forward(score, value, index, output):
    """Writes output [B,N,H,D] in place (the reference's contract: caller-allocated output)."""
    _F.value_agg(score, value, index, output)


def value_aggregation_backward(grad_output, score, value, index, grad_score, grad_value):
    raise NotImplementedError('casmtr_b200 implements the inference (forward) path only')
