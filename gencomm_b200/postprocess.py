"""Drop-in ``VoxelPostprocessor`` inference half: anchors, decode, rotated NMS.

Mirrors ``opencood/data_utils/post_processor/voxel_postprocessor.py``: ``VoxelPostprocessor(anchor_params, train)``,
``generate_anchor_box()`` (:68-121, host numpy like the reference -- it runs once per dataset), ``post_process(data_dict,
output_dict)`` (:1084-1244) returning ``(pred_box3d_tensor [K,8,3], scores [K])`` or ``(None, None)``, and the static
``delta_to_boxes3d`` is folded into the kernels.  The whole of post_process -- sigmoid / threshold, box decode, direction
fix, corners, projection, size and z filters, top-1000 rotated NMS, range mask -- runs on the GPU through ``gc_postprocess``
(csrc/postprocess.cu); the reference moves the candidates to the host and loops over shapely polygons in Python.
``post_process_batch`` is the sync-free form (padded outputs + counts) used for the per-frame detection all-gather
(SURVEY.md section 8e).  Label generation / training targets are out of scope.  No CPU path.
"""
import math

import numpy as np
import torch

from . import ops

TOP = 1000   # box_utils.py:941


class VoxelPostprocessor:
    def __init__(self, anchor_params, train=False, class_names=None):
        self.params = anchor_params
        self.train = train
        self.anchor_num = self.params['anchor_args']['num']
        self.max_num = self.params.get('max_num')
        self._post = None
        self._ws = None

    def generate_anchor_box(self):
        a = self.params['anchor_args']
        r = a['r']
        assert self.anchor_num == len(r)
        r = [math.radians(e) for e in r]
        vh, vw = a['vh'], a['vw']
        xrange = [a['cav_lidar_range'][0], a['cav_lidar_range'][3]]
        yrange = [a['cav_lidar_range'][1], a['cav_lidar_range'][4]]
        stride = a['feature_stride'] if 'feature_stride' in a else 2
        x = np.linspace(xrange[0] + vw, xrange[1] - vw, a['W'] // stride)
        y = np.linspace(yrange[0] + vh, yrange[1] - vh, a['H'] // stride)
        cx, cy = np.meshgrid(x, y)
        cx = np.tile(cx[..., np.newaxis], self.anchor_num)
        cy = np.tile(cy[..., np.newaxis], self.anchor_num)
        cz = np.ones_like(cx) * -1.0
        w, l, h = np.ones_like(cx) * a['w'], np.ones_like(cx) * a['l'], np.ones_like(cx) * a['h']
        r_ = np.ones_like(cx)
        for i in range(self.anchor_num):
            r_[..., i] = r[i]
        if self.params['order'] == 'hwl':
            return np.stack([cx, cy, cz, h, w, l, r_], axis=-1)
        if self.params['order'] == 'lhw':
            return np.stack([cx, cy, cz, l, h, w, r_], axis=-1)
        raise ValueError('Unknown bbx order.')          # the reference calls sys.exit here (:119)

    def _params(self):
        if self._post is None:
            p, d = self.params, self.params.get('dir_args', {'dir_offset': 0.0, 'num_bins': 1})
            self._post = ops.make_post_params(p['target_args']['score_threshold'], p['nms_thresh'], d['dir_offset'],
                                              d['num_bins'], p['order'], p['gt_range'], TOP)
        return self._post

    def post_process_batch(self, cls_preds, reg_preds, dir_preds, anchor_box, transformation_matrix=None):
        """Sync-free, batched over frames: returns (boxes [B,1000,8,3], scores [B,1000], counts [B] i32) on the device."""
        dev = cls_preds.device
        anchors = torch.as_tensor(anchor_box).to(device=dev, dtype=torch.float32).contiguous()
        tfm = None
        if transformation_matrix is not None:
            tfm = torch.as_tensor(transformation_matrix).to(device=dev, dtype=torch.float32).reshape(-1, 4, 4).contiguous()
        return ops.postprocess(cls_preds.contiguous(), reg_preds.contiguous(),
                               None if dir_preds is None else dir_preds.contiguous(), anchors, self._params(), tfm)

    def post_process(self, data_dict, output_dict):
        """Reference contract (:1084-1244).  Intermediate / early fusion: ``output_dict`` holds the ego only; with
        several cavs (late fusion) their candidates would have to be merged before the NMS, which is outside the
        GenComm path -> NotImplementedError."""
        if len(output_dict) != 1:
            raise NotImplementedError("gencomm_b200 VoxelPostprocessor.post_process: one cav (the ego of an intermediate-"
                                      "fusion model) per call; late fusion is outside the GenComm hot path")
        (cav_id, out), = output_dict.items()
        assert cav_id in data_dict
        cav = data_dict[cav_id]
        # key handling exactly as the reference (:1122-1127): 'psm' renames cls_preds; the 'rm' / 'dm' renames test the
        # OUTER dict ('rm' in output_dict), where they can only match a cav id -- kept, so psm/rm/dm-style models read
        # 'reg_preds' and get the direction fix only when 'dir_preds' itself is present
        cls = out['psm'] if 'psm' in out else out['cls_preds']
        reg = out['rm'] if 'rm' in output_dict else out['reg_preds']
        dr = out['dm'] if 'dm' in output_dict else out.get('dir_preds')
        assert cls.shape[0] == 1                                      # batch size 1 during testing (:1153)
        if 'iou_preds' in out:
            raise NotImplementedError("gencomm_b200 VoxelPostprocessor: iou_preds rescoring is not on the GenComm path")
        boxes, scores, counts = self.post_process_batch(cls, reg, dr, cav['anchor_box'], cav['transformation_matrix'])
        k = int(counts[0].item())                                     # the one host sync: the result is ragged
        if k == 0:
            return None, None
        return boxes[0, :k].clone(), scores[0, :k].clone()
