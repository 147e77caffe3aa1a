"""Drop-in ``ResultParser`` (reference: /root/reference/acr/result_parser.py:7-190).

``parse(outputs, meta_data, cfg)`` takes the reference's dict of NCHW maps; inside the fused
pipeline ``parse_engine`` reads the engine's NHWC fp32 maps in place.  Either way the work is three
small kernels (acr_b200_parse) and exactly one device->host read (the two hand counts) instead
of the reference's >=6 implicit syncs.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from acr.config import args
from acr_b200 import ops as _ops


class ResultParser(nn.Module):
    def __init__(self):
        super().__init__()
        self.map_size = args().centermap_size
        self.part_name = ['cam', 'global_orient', 'hand_pose', 'betas']
        self.part_idx = [args().cam_dim, args().rot_dim, (args().mano_theta_num - 1) * args().rot_dim, 10]
        self.kps_num = 21
        self.params_num = int(np.array(self.part_idx).sum())
        if (args().prior_mode, args().inter_prior, args().Rot_type, self.map_size) != ('cross', True, '6D', 64):
            raise ValueError("only prior_mode='cross', inter_prior=True, Rot_type='6D', centermap_size=64 "
                             "(the reference's shipped configuration) are supported")
        self._pbufs = {}

    def _parse_buffers(self, B, device):
        key = (B, str(device))
        if key not in self._pbufs:
            self._pbufs[key] = _ops.ParseBuffers(B, device)
        return self._pbufs[key]

    # ------------------------------------------------------------------ kernels
    def launch(self, maps, B, meta_data, device):
        """Enqueue the parse kernels; returns the worst-case buffers (no sync)."""
        bufs = self._parse_buffers(B, device)
        ids = meta_data.get('batch_ids') if meta_data is not None else None
        offs = meta_data.get('offsets') if meta_data is not None else None
        _ops.parse_maps(maps, B, bufs, ids, offs, args().centermap_conf_thresh)
        return bufs

    @staticmethod
    def collect(bufs, outputs, meta_data):
        """One D2H read of (L, R), then slice the buffers into the reference's output schema."""
        L, R = (int(v) for v in bufs.counts[:2].tolist())
        N = L + R
        outputs['params_pred'] = bufs.params_pred[:N]
        outputs['l_params_pred'], outputs['r_params_pred'] = bufs.params_pred[:L], bufs.params_pred[L:N]
        outputs['detection_flag'] = bufs.detection_flag[:N]
        outputs['detection_flag_cache'] = bufs.detection_flag[:N].bool()
        outputs['l_centers_pred'], outputs['r_centers_pred'] = bufs.centers_pred[:L], bufs.centers_pred[L:N]
        outputs['l_centers_conf'] = bufs.centers_conf[:L].unsqueeze(1)
        outputs['r_centers_conf'] = bufs.centers_conf[L:N].unsqueeze(1)
        dev = bufs.counts.device
        outputs['left_hand_num'] = torch.tensor([L], device=dev)
        outputs['right_hand_num'] = torch.tensor([R], device=dev)
        outputs['reorganize_idx'] = bufs.reorganize_idx[:N]
        outputs['output_hand_type'] = bufs.hand_type[:N]
        outputs['params_dict'] = dict(cam=bufs.cam[:N], global_orient=bufs.global_orient[:N],
                                      hand_pose=bufs.hand_pose[:N], betas=bufs.betas[:N], poses=bufs.poses[:N])
        if meta_data is not None:
            bi = bufs.batch_ids[:N]
            for key in ('image', 'offsets', 'imgpath'):      # result_parser.py:186-187
                if key in meta_data:
                    v = meta_data[key]
                    if isinstance(v, torch.Tensor):
                        meta_data[key] = v[bi.to(v.device)]
                    elif isinstance(v, list):
                        meta_data[key] = np.array(v)[bi.cpu().numpy()]
        return outputs, meta_data

    # ---------------------------------------------------------- reference entry
    @torch.no_grad()
    def parse(self, outputs, meta_data, cfg=None):
        """Reference signature: NCHW fp32 maps in ``outputs`` (result_parser.py:21-40)."""
        names = dict(l_center='l_center_map', r_center='r_center_map', l_params='l_params_maps',
                     r_params='r_params_maps', l_prior='l_prior_maps', r_prior='r_prior_maps')
        maps = {}
        for k, n in names.items():
            t = outputs[n].float().permute(0, 2, 3, 1).contiguous()
            maps[k] = (t, t.shape[-1])
        B = outputs['l_center_map'].shape[0]
        bufs = self.launch(maps, B, meta_data, outputs['l_center_map'].device)
        return self.collect(bufs, outputs, meta_data)
