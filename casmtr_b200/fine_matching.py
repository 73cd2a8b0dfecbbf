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


class CascadeFinePreprocess(nn.Module):
    """reference :13-67 (SURVEY section 8f "next" #4): crops the W x W fine-level windows of the predicted matches.  Same
    constructor and parameters (down_proj / merge_feat when fine_concat_coarse_feat) as the reference; the F.unfold of the
    whole fine map is replaced by a gather of the M windows (casmtr_fine_window_gather), the two nn.Linear stay cuBLAS."""

    def __init__(self, config, config_fine, config_coarse, coarse_level):
        super().__init__()
        self.config = config
        self.cat_c_feat = config['fine_concat_coarse_feat']
        self.W = self.config['fine_window_size']
        self.coarse_level = coarse_level
        d_model_c, d_model_f = config_coarse['d_model'], config_fine['d_model']
        self.d_model_f = d_model_f
        if self.cat_c_feat:
            self.down_proj = nn.Linear(d_model_c, d_model_f, bias=True)
            self.merge_feat = nn.Linear(2 * d_model_f, d_model_f, bias=True)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.kaiming_normal_(p, mode='fan_out', nonlinearity='relu')

    def forward(self, feat_f0, feat_f1, feat_c0, feat_c1, data):
        W = self.W
        stride = data['hw0_f'][0] // data[f'hw0_{self.coarse_level}'][0]
        data.update({'W': W})
        st = data[f'stage_{self.coarse_level}']
        if st['b_ids'].shape[0] == 0:
            z = torch.empty(0, W ** 2, self.d_model_f, device=feat_f0.device)
            return z, z.clone()
        b, i, j = st['b_ids'].contiguous(), st['i_ids'].contiguous(), st['j_ids'].contiguous()
        f0 = F.fine_window_gather(feat_f0.float().contiguous(), b, i, data[f'hw0_{self.coarse_level}'][1], stride, W)
        f1 = F.fine_window_gather(feat_f1.float().contiguous(), b, j, data[f'hw1_{self.coarse_level}'][1], stride, W)
        if self.cat_c_feat:
            c_win = self.down_proj(torch.cat([feat_c0[b, i], feat_c1[b, j]], 0))                         # [2n, c]
            cf = self.merge_feat(torch.cat([torch.cat([f0, f1], 0), c_win.unsqueeze(1).expand(-1, W ** 2, -1)], -1))
            f0, f1 = torch.chunk(cf, 2, dim=0)
        return f0, f1


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
