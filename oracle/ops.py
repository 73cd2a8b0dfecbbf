"""Oracle restatements of the reference's three forward CUDA kernels (torch CPU).

Test infrastructure only -- see oracle/__init__.py.
"""
import torch


def score5d(query, key, index):
    """Sparse gathered Q.K^T shared by the 4 sibling queries of a parent.

    Follows ``ScoreData`` cuda_imp/QuadTreeAttention/QuadtreeAttention/src/score_computation_kernal.cu:22-62
    (host wrapper :65-92): out[b,n,f,k,h] = sum_d query[b,n,f,h,d] * key[b, index[b,n,k,h], h, d].

    query [B,N1,4,H,D] fp32, key [B,N2,H,D] fp32, index [B,N1,K,H] int64 -> [B,N1,4,K,H] fp32
    """
    B, N1, F, H, D = query.shape
    K = index.shape[2]
    key_hm = key.permute(0, 2, 1, 3)                               # [B,H,N2,D]
    flat = index.permute(0, 3, 1, 2).reshape(B, H, N1 * K)         # [B,H,N1*K]
    picked = torch.gather(key_hm, 2, flat.unsqueeze(-1).expand(B, H, N1 * K, D))
    picked = picked.reshape(B, H, N1, K, D)
    q_hm = query.permute(0, 3, 1, 2, 4)                            # [B,H,N1,F,D]
    out = torch.matmul(q_hm, picked.transpose(-1, -2))             # [B,H,N1,F,K]
    return out.permute(0, 2, 3, 4, 1).contiguous()


def value_agg(score, value, index):
    """Sparse gathered A.V.

    Follows ``ValueAggregationForwardFunc`` cuda_imp/QuadTreeAttention/QuadtreeAttention/src/value_aggregation_kernel.cu:21-42:
    out[b,n,h,d] = sum_k score[b,n,k,h] * value[b, index[b,n,k,h], h, d].

    score [B,N,K,H] fp32, value [B,M,H,D] fp32, index [B,N,K,H] int64 -> [B,N,H,D] fp32
    """
    B, N, K, H = score.shape
    D = value.shape[-1]
    val_hm = value.permute(0, 2, 1, 3)                             # [B,H,M,D]
    flat = index.permute(0, 3, 1, 2).reshape(B, H, N * K)
    picked = torch.gather(val_hm, 2, flat.unsqueeze(-1).expand(B, H, N * K, D))
    picked = picked.reshape(B, H, N, K, D)
    w = score.permute(0, 3, 1, 2).unsqueeze(-2)                    # [B,H,N,1,K]
    out = torch.matmul(w, picked).squeeze(-2)                      # [B,H,N,D]
    return out.permute(0, 2, 1, 3).contiguous()


def value_agg5(score, value, index):
    """The reference's Python wrapper around value aggregation
    (cuda_imp/QuadTreeAttention/QuadtreeAttention/functions/quadtree_attention.py:25-38):
    score/index [B,N,F,K,H] are flattened to [B,N*F,K,H]; result [B,N,F,H,D]."""
    B, N, F, K, H = score.shape
    out = value_agg(score.reshape(B, N * F, K, H), value, index.reshape(B, N * F, K, H))
    return out.reshape(B, N, F, H, value.shape[-1])


def score3d(query, key, index, chunk=8192):
    """Single-head sparse correlation volume.

    Follows ``score_computation_forward_kernel`` cuda_imp/score_cuda/src/score_computation_kernel.cu:23-40:
    out[b,n,k] = sum_c query[b,n,c] * key[b, index[b,n,k], c].

    query [B,N1,C], key [B,N2,C], index [B,N1,K] int64 -> [B,N1,K] fp32.
    Processed in chunks of query rows to bound the [rows,K,C] gather.
    """
    B, N1, C = query.shape
    K = index.shape[2]
    out = query.new_empty(B, N1, K)
    for b in range(B):
        for s in range(0, N1, chunk):
            e = min(N1, s + chunk)
            rows = key[b].index_select(0, index[b, s:e].reshape(-1)).reshape(e - s, K, C)
            out[b, s:e] = torch.bmm(rows, query[b, s:e].unsqueeze(-1)).squeeze(-1)
    return out
