"""Drop-in replacements for the reference's fine matching heads (src/model/functions/fine_matching.py):
``CascadeFineMatching`` (:70-137) and the legacy ``FineMatching`` (:195-261).  One kernel instead of
einsum + softmax + kornia dsnt + meshgrid + variance ops."""
import math

import torch
import torch.nn as nn

from . import functional as F


def _run(feat_f0, feat_f1, data, mk0, mk1, b_ids, n_keep):
    M, WW, C = feat_f0.shape
    scale = data['hw0_i'][0] / data['hw0_f'][0]
    s1 = data['scale1'] if 'scale0' in data else None          # the reference tests 'scale0' and reads 'scale1' (:130)
    if mk1.shape[0] < M:        # training-style padding: more windows than matches, extra rows are sliced off below
        mk1 = torch.cat([mk1, mk1.new_zeros(M - mk1.shape[0], 2)], 0)
        b_ids = torch.cat([b_ids, b_ids.new_zeros(M - b_ids.shape[0])], 0)
    expec, mk1f = F.fine_match_forward(feat_f0.to(torch.float32).contiguous(), feat_f1.to(torch.float32).contiguous(),
                                       mk1[:M], scale, s1, b_ids[:M])
    data.update({'expec_f': expec, 'mkpts0_f': mk0, 'mkpts1_f': mk1f[:n_keep]})


class CascadeFineMatching(nn.Module):
    """FineMatching with s2d paradigm (reference :70-137)."""

    def __init__(self, coarse_level='4c'):
        super().__init__()
        self.coarse_level = coarse_level

    def forward(self, feat_f0, feat_f1, data):
        """feat_f0/feat_f1 [M,WW,C]; updates data['expec_f'] [M,3], data['mkpts0_f'], data['mkpts1_f'] [M,2]."""
        M, WW, C = feat_f0.shape
        self.M, self.W, self.WW, self.C = M, int(math.sqrt(WW)), WW, C
        self.scale = data['hw0_i'][0] / data['hw0_f'][0]
        st = data[f'stage_{self.coarse_level}']
        if M == 0:
            assert self.training is False, 'M is always >0, when training, see coarse_matching.py'
            data.update({'expec_f': torch.empty(0, 3, device=feat_f0.device),
                         'mkpts0_f': st['mkpts0_c'], 'mkpts1_f': st['mkpts1_c']})
            return
        _run(feat_f0, feat_f1, data, st['mkpts0_c'], st['mkpts1_c'], st['b_ids'], len(st['mconf']))


class FineMatching(nn.Module):
    """Legacy head (reference :195-261): same arithmetic, reads the coarse matches from the top level of ``data``."""

    def __init__(self):
        super().__init__()

    def forward(self, feat_f0, feat_f1, data):
        M, WW, C = feat_f0.shape
        self.M, self.W, self.WW, self.C = M, int(math.sqrt(WW)), WW, C
        self.scale = data['hw0_i'][0] / data['hw0_f'][0]
        if M == 0:
            assert self.training is False, 'M is always >0, when training, see coarse_matching.py'
            data.update({'expec_f': torch.empty(0, 3, device=feat_f0.device),
                         'mkpts0_f': data['mkpts0_c'], 'mkpts1_f': data['mkpts1_c']})
            return
        _run(feat_f0, feat_f1, data, data['mkpts0_c'], data['mkpts1_c'], data['b_ids'], len(data['mconf']))
