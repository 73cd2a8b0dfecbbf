"""Drop-in for the pybind module built from cuda_imp/score_cuda/src/score_computation.cpp:29-32."""
from casmtr_b200 import functional as _F


def score_forward(query, key, index):
    """query [B,N1,C], key [B,N2,C], index [B,N1,K] int64 -> [out [B,N1,K]]"""
    return [_F.score3d(query, key, index)]


def score_backward(grad_output, query, key, index):
    """-> [grad_query [B,N1,C], grad_key [B,N2,C]] (reference score_cuda/src/score_computation.cpp:19-27)"""
    return list(_F.score3d_backward(grad_output, query, key, index))
