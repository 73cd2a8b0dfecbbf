"""Oracle restatement of (Cascade)FineMatching (torch CPU, fp32).

Test infrastructure only -- see oracle/__init__.py.
Reference: src/model/functions/fine_matching.py:77-137 (CascadeFineMatching) and
:201-261 (legacy FineMatching, same arithmetic).  kornia 0.6.2 (requirements.txt:5,
not vendored) supplies ``dsnt.spatial_expectation2d`` and ``create_meshgrid``; their
published behaviour is restated here: normalised grid = linspace(-1, 1, W) with
[...,0]=x, [...,1]=y, expectation = (sum x*p, sum y*p).
"""
import math

import torch


def normalized_grid(W):
    lin = (torch.linspace(0, W - 1, W) / (W - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(lin, lin, indexing='ij')
    return torch.stack([gx, gy], dim=-1).reshape(W * W, 2)       # [WW,2] (x,y)


def fine_match(feat_f0, feat_f1, mkpts1_c, scale, scale1=None):
    """feat_f0/feat_f1 [M,WW,C] fp32; mkpts1_c [M,2]; scale = hw0_i[0]/hw0_f[0];
    scale1 optional [M,2] per-match image scale (already indexed by b_ids).
    Returns expec_f [M,3] (x, y, std) and mkpts1_f [M,2] (fine_matching.py:104-137).
    """
    M, WW, C = feat_f0.shape
    W = int(math.sqrt(WW))
    centre = feat_f0[:, WW // 2, :]                                        # :105
    sim = torch.bmm(feat_f1, centre.unsqueeze(-1)).squeeze(-1)             # :106
    heat = torch.softmax(sim * (1.0 / C ** 0.5), dim=1)                    # :107-108
    grid = normalized_grid(W).to(feat_f0.device)
    coords = heat @ grid                                                   # :111 spatial_expectation2d
    var = heat @ (grid ** 2) - coords ** 2                                 # :115
    std = torch.sqrt(torch.clamp(var, min=1e-10)).sum(-1)                  # :116
    expec = torch.cat([coords, std.unsqueeze(1)], dim=-1)
    s1 = scale * scale1 if scale1 is not None else scale                   # :130
    mkpts1_f = mkpts1_c + coords * (W // 2) * s1                           # :131
    return expec, mkpts1_f
