"""Drop-in for the inference statistics of the reference's dense coarse matcher (src/model/functions/coarse_matching.py:22-89,
SURVEY.md section 8f "next" #1).  The reference forms sim [B,L,S], two soft-maxes and their product (four 468 MB tensors at
832x832); the cascade stages only consume next_idx_c01/c10 and next_conf_c01/c10, which libcasmtr_b200 computes on the
tensor cores without materialising the matrix (casmtr_coarse_match_fwd).

The mutual-nearest-neighbour match list of get_coarse_match (:91-153; read by a stage-1 model's temp_outputs,
cascade_model_stage3.py:71-74) comes from a second tensor-core pass: conf = softmax_i * softmax_j is a function of sim and of the row /
column log-sum-exps the first pass produced, so its row and column arg-maxima need no matrix either; casmtr_match_extract (coarse
mode) then applies the threshold, the border removal and the mutual test and emits the list in torch.where order.
What is NOT produced: `conf_matrix` itself (training supervision only); it is set to None.
Padding masks (mask_c0 / mask_c1, reference :64-65) are applied inside the kernel: padded columns take no part, padded rows
come out as the reference's constant -1e9 rows do (uniform soft-max: next_conf 1 / columns, next_idx 0)."""
import torch.nn as nn

from . import functional as F


class CoarseMatching(nn.Module):
    def __init__(self, config, coarse_config=None):
        super().__init__()
        self.config = config
        self.thr = config['thr']
        self.border_rm = config['border_rm']
        self.train_coarse_percent = config['train_coarse_percent']
        self.train_pad_num_gt_min = config['train_pad_num_gt_min']
        self.next_topk = coarse_config.get('next_topk', None) if coarse_config is not None else None
        self.match_type = config['match_type']
        self.temperature = config['dsmax_temperature']
        assert self.match_type == 'dual_softmax'
        self.match_list = True      # False: skip the second pass and leave the 1/8 match list None (the cascade stages do not read it)

    def forward(self, feat_c0, feat_c1, data, mask_c0=None, mask_c1=None, level='8c'):
        """feat_c0 [N,L,C], feat_c1 [N,S,C] -> data[f'stage_{level}'] with next_idx_c01/c10 [N,L]/[N,S] int64 and
        next_conf_c01/c10 fp32 (reference :70-84)."""
        if self.training:
            raise NotImplementedError('casmtr_b200.CoarseMatching implements the inference statistics only')
        if (mask_c0 is None) != (mask_c1 is None):
            raise RuntimeError('CoarseMatching: the reference masks with mask_c0 * mask_c1 (:65), give both or neither')
        o = F.coarse_match_forward(feat_c0.float().contiguous(), feat_c1.float().contiguous(), self.temperature, mask_c0, mask_c1,
                                   mutual=self.match_list)
        m = {k: None for k in ('b_ids', 'i_ids', 'j_ids', 'gt_mask', 'm_bids', 'mkpts0_c', 'mkpts1_c', 'mconf')}
        if self.match_list:
            m = self.get_coarse_match(o, data, level)
        data[f'stage_{level}'] = {
            'conf_matrix': None, 'next_conf_c01_topk': None, 'next_idx_c01_topk': None,
            'next_conf_c10_topk': None, 'next_idx_c10_topk': None,
            'next_idx_c01': o['next_idx01'], 'next_idx_c10': o['next_idx10'],
            'next_conf_c01': o['next_conf01'], 'next_conf_c10': o['next_conf10'],
            'next_conf_c01_s': None, 'next_idx_c01_s': None, **m}

    def get_coarse_match(self, o, data, level):
        """reference :91-153 (inference): conf > thr, border removal on both grids (padded variant when data holds mask_{level}0/1),
        mutual nearest neighbours of conf, matches in torch.where order, keypoints at image resolution."""
        padded = f'mask_{level}0' in data
        r = F.match_extract(o['mconf_row'], o['midx_row'], o['midx_col'], tuple(data[f'hw0_{level}']), tuple(data[f'hw1_{level}']),
                            tuple(data['hw0_i']), test_thr=self.thr, border_rm=self.border_rm, nms_window=None, double_check=True,
                            pad_mask0=data[f'mask_{level}0'] if padded else None, pad_mask1=data[f'mask_{level}1'] if padded else None,
                            scale0=data.get('scale0'), scale1=data.get('scale1'), coarse_mode=True)
        return {'b_ids': r['b_ids'], 'i_ids': r['i_ids'], 'j_ids': r['j_ids'], 'gt_mask': r['mconf'] == 0, 'm_bids': r['b_ids'],
                'mkpts0_c': r['mkpts0_c'], 'mkpts1_c': r['mkpts1_c'], 'mconf': r['mconf']}
