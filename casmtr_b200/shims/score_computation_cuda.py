"""Drop-in for the pybind module built from
cuda_imp/QuadTreeAttention/QuadtreeAttention/src/score_computation.cpp:35-38."""
from casmtr_b200 import functional as _F


def score_forward(query, key, index):
    """query [B,N1,4,H,D], key [B,N2,H,D], index [B,N1,K,H] int64 -> [out [B,N1,4,K,H]]"""
    return [_F.score5d(query, key, index)]


def score_backward(grad_output, query, key, index):
    """-> [grad_query [B,N1,4,H,D], grad_key [B,N2,H,D]] (reference score_computation.cpp:22-33)"""
    return list(_F.score5d_backward(grad_output, query, key, index))
