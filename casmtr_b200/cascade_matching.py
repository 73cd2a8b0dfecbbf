"""Drop-in replacements for the reference's cascade matching stage (inference):
``ScoreComputation`` (src/model/functions/cascade_functions.py:8-22), ``PostProcess``
(src/model/functions/post_processing.py:35-44,111-121) and ``CascadeMatching``
(src/model/functions/cascade_matching.py:36-331, inference branch).  Same constructor / forward
signatures and the same keys written into ``data``; the ~40 torch kernels and three host syncs of
the reference collapse into two library calls (casmtr_cascade_match_fwd, casmtr_match_extract).
"""
import torch
import torch.nn as nn

from . import functional as F


class ScoreComputation(torch.autograd.Function):
    """reference cascade_functions.py:8-22; query [B,N1,C], key [B,N2,C], index [B,N1,K] -> [B,N1,K]."""

    @staticmethod
    def forward(ctx, query, key, index):
        # the reference's launch-geometry guards (:11-12) do not apply to this kernel and are lifted
        ctx.save_for_backward(query, key, index)
        return F.score3d(query, key, index)

    @staticmethod
    def backward(ctx, grad_output):                                      # reference :17-22
        query, key, index = ctx.saved_tensors
        gq, gk = F.score3d_backward(grad_output.contiguous(), query, key, index)
        return gq, gk, None


class PostProcess(object):
    """reference post_processing.py:35-44, 111-121: method None (threshold) or 'maxpool_nms'.  The other
    detectors (sift / local_window_nms / softargmax_nms / d2d) are selected by no shipped config."""

    def __init__(self, post_config):
        self.config = post_config
        self.method = post_config['method']
        if self.method not in (None, 'maxpool_nms'):
            raise NotImplementedError(f"post-processing method {self.method!r} is not implemented by casmtr_b200")
        if self.method == 'maxpool_nms' and post_config.get('stride', 1) != 1:
            raise NotImplementedError('maxpool_nms stride != 1')

    @property
    def nms_window(self):
        return self.config['window_size'] if self.method == 'maxpool_nms' else None

    def apply(self, data, axes_lengths, next_idx_c01, next_conf_c01, test_thr, level):
        """-> bool mask [B, L0]: detection only (NMS/threshold), exactly the reference's method."""
        B = next_conf_c01.shape[0]
        hw0 = (axes_lengths['h0c'], axes_lengths['w0c'])
        hw1 = (axes_lengths['h1c'], axes_lengths['w1c'])
        zeros = torch.zeros(B, hw1[0] * hw1[1], dtype=torch.int64, device=next_conf_c01.device)
        r = F.match_extract(next_conf_c01.contiguous(), next_idx_c01.contiguous(), zeros, hw0, hw1, (hw0[0], hw0[1]),
                            test_thr=test_thr, border_rm=0, nms_window=self.nms_window, double_check=False)
        return r['mask']            # flags before the "keep element 0" fallback, which the reference applies later


class CascadeMatching(nn.Module):
    def __init__(self, config, cas_config, stage=None):
        super().__init__()
        self.config = config
        self.cas_config = cas_config
        self.thr = config['thr']
        self.test_thr = config['test_thr']
        self.pre_thr = config['pre_thr']
        self.border_rm = config['border_rm']
        self.double_check = config['double_check']
        self.train_pad_num_gt_min = config['train_pad_num_gt_min']
        self.propagation = cas_config['propagation']
        self.dilated = cas_config['dilated']
        self.post_process = PostProcess(post_config=cas_config['post_config'])
        self.detector_mode = cas_config.get('detector_mode', None)
        self.grid_size = cas_config.get('grid_size', None)
        self.rt = cas_config['post_config'].get('rt', None)
        self.rd = cas_config['post_config'].get('rd', None)
        self.stage = stage
        self.next_topk = cas_config.get('next_topk', None)
        self.match_type = config['match_type']
        assert self.match_type == 'softmax'
        self.temperature = config['dsmax_temperature']
        if self.rt is not None or self.rd is not None:
            raise NotImplementedError('ratio tests rt/rd are dead in every shipped config and not implemented')
        self.store_conf_matrix = True       # set False to skip writing the [B,L,K] softmax volume
        self.defer_sync = False             # True: no host sync in forward (CUDA-graph capture); call finalize(data, level) later

    def forward(self, feat_c0, feat_c1, idx_c01, idx_c10, data, mask_c0=None, mask_c1=None,
                heatmap_c0=None, level='4c', pre_level='8c'):
        """feat_c0 [B,HW0,C], feat_c1 [B,HW1,C], idx_c01 [B,HW0,4ww], idx_c10 [B,HW1,4ww] (int64 key
        indices), optional masks [B,HW].  Updates data[f'stage_{level}'] like the reference (:151-168)."""
        if self.training:
            raise NotImplementedError('casmtr_b200.CascadeMatching implements the inference branch only')
        o = F.cascade_match_forward(feat_c0.to(torch.float32).contiguous(), feat_c1.to(torch.float32).contiguous(),
                                    idx_c01.contiguous(), idx_c10.contiguous(), mask_c0, mask_c1,
                                    temperature=self.temperature, need_conf=self.store_conf_matrix,
                                    need_conf10=False,       # the reference keeps only conf_matrix01 (:155)
                                    w0=data[f'hw0_{level}'][1] if f'hw0_{level}' in data else 0,
                                    w1=data[f'hw1_{level}'][1] if f'hw1_{level}' in data else 0)
        data[f'stage_{level}'] = {
            'conf_matrix': o['conf01'], 'detector_matrix01': None,
            'next_conf_c01_topk': None, 'next_idx_c01_topk': None,
            'next_conf_c10_topk': None, 'next_idx_c10_topk': None,
            'idx_c01': idx_c01, 'idx_c10': idx_c10,
            'next_idx_c01': o['next_idx01'], 'next_idx_c10': o['next_idx10'],
            'next_conf_c01': o['next_conf01'], 'next_conf_c10': o['next_conf10'],
            'next_conf_c01_s': None, 'next_idx_c01_s': None}
        match_result = self.get_coarse_match(o['conf01'], idx_c01, o['next_conf01'], o['next_idx01'], o['next_idx10'],
                                             data, level, pre_level)
        if self.defer_sync:
            data[f'stage_{level}']['_deferred'] = match_result
            return
        data[f'stage_{level}'].update(**match_result)
        if 'm_bids' in match_result:
            data['m_bids'] = match_result['m_bids']

    def finalize(self, data, level='4c'):
        """Completes a defer_sync forward: reads the match count and fills the match list keys of data[f'stage_{level}']."""
        r = F.trim_matches(data[f'stage_{level}'].pop('_deferred'))
        res = {'b_ids': r['b_ids'], 'i_ids': r['i_ids'], 'j_ids': r['j_ids'], 'm_bids': r['b_ids'],
               'mkpts0_c': r['mkpts0_c'], 'mkpts1_c': r['mkpts1_c'], 'mconf': r['mconf']}
        data[f'stage_{level}'].update(**res)
        data['m_bids'] = res['m_bids']

    def get_coarse_match(self, conf_matrix01, idx_c01, next_conf_c01, next_idx_c01, next_idx_c10, data, level, pre_level):
        """reference :170-261, 316-331 (inference): one fused, sync-free-until-the-count extraction."""
        if type(pre_level) != list:
            pre_level = [pre_level]
        hw0, hw1 = tuple(data[f'hw0_{level}']), tuple(data[f'hw1_{level}'])
        pre = [(data[f'stage_{p}']['next_conf_c01'].detach(), data[f'hw0_{p}'][0], data[f'hw0_{p}'][1]) for p in pre_level]
        padded = f'mask_{level}0' in data
        r = F.match_extract(next_conf_c01, next_idx_c01, next_idx_c10, hw0, hw1, tuple(data['hw0_i']),
                            test_thr=self.test_thr, border_rm=self.border_rm, nms_window=self.post_process.nms_window,
                            pre_confs=pre, pre_thrs=self.pre_thr, double_check=self.double_check,
                            pad_mask0=data[f'mask_{level}0'] if padded else None,
                            pad_mask1=data[f'mask_{level}1'] if padded else None,
                            scale0=data.get('scale0'), scale1=data.get('scale1'), defer=self.defer_sync)
        if self.defer_sync:
            return r
        return {'b_ids': r['b_ids'], 'i_ids': r['i_ids'], 'j_ids': r['j_ids'], 'm_bids': r['b_ids'],
                'mkpts0_c': r['mkpts0_c'], 'mkpts1_c': r['mkpts1_c'], 'mconf': r['mconf']}
