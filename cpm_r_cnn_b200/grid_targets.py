"""Grid-point training targets on the device: GridLossComputation.prepare_target
(pet/rcnn/modeling/grid_cascade_rcnn/loss.py:178-258) without its CPU triple loop and H2D upload.  The reference's
subsample() moves the positive boxes to the host (`.cpu()`, :161-162) only because prepare_target loops over them in
Python; with this op they stay where they are."""
import ctypes

import torch

from . import _lib
from .grid_decode import STAGE_MAPPING_RATIO, calc_sub_regions


def prepare_grid_target(pos_bboxes, pos_gt_bboxes, mapping_ratio, pos_radius=1, grid_points=9, roi_feat_size=14,
                        target_refine=False):
    """pos_bboxes, pos_gt_bboxes: (R,4) xyxy CUDA tensors (positive RoIs, their matched ground truth)
    -> (R, grid_points, 2*roi_feat_size, 2*roi_feat_size) float32 targets of 0/1 on the same device."""
    _lib.require_cuda(pos_bboxes, "pos_bboxes")
    _lib.require_cuda(pos_gt_bboxes, "pos_gt_bboxes")
    assert pos_bboxes.shape == pos_gt_bboxes.shape and pos_bboxes.dim() == 2 and pos_bboxes.shape[1] == 4
    R = pos_bboxes.shape[0]
    grid_size = int(round(grid_points ** 0.5))
    map_size = roi_feat_size * 4
    half = map_size // 4 * 2
    sub = calc_sub_regions(grid_points, grid_size, map_size)
    pos = pos_bboxes.float().contiguous()
    gt = pos_gt_bboxes.float().contiguous()
    out = torch.empty((R, grid_points, half, half), dtype=torch.float32, device=pos.device)
    if R:
        sub_xy = (ctypes.c_int32 * (2 * grid_points))(*[int(v) for s in sub for v in (s[0], s[1])])
        with _lib.device_of(pos):
            _lib.check(_lib.lib().cpm_grid_targets(_lib.ptr(pos), _lib.ptr(gt), R, grid_points, map_size, sub_xy,
                                                   float(mapping_ratio), int(pos_radius), int(bool(target_refine)),
                                                   _lib.ptr(out), _lib.stream_ptr(pos.device)))
    return out


class GridTargetGenerator(object):
    """The target half of GridLossComputation (loss.py:113-132, :178-258): same constructor fields; prepare_target takes
    the (pos_bboxes, pos_gt_bboxes) pair the reference keeps in self.pos_result."""

    def __init__(self, stage, pos_radius=1, grid_points=9, roi_feat_size=14, mapping_ratios=STAGE_MAPPING_RATIO,
                 target_refine=False):
        self.stage = stage
        self.pos_radius = pos_radius
        self.grid_points = grid_points
        self.roi_feat_size = roi_feat_size
        self.mapping_ratio = mapping_ratios[stage]
        self.target_refine = target_refine

    def prepare_target(self, pos_bboxes, pos_gt_bboxes):
        return prepare_grid_target(pos_bboxes, pos_gt_bboxes, self.mapping_ratio, self.pos_radius, self.grid_points,
                                   self.roi_feat_size, self.target_refine)
