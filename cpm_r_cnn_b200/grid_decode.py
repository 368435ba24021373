"""Grid-point heat-map -> box decode (GridPostProcessor.get_boxes, grid_cascade_rcnn/inference.py:189-279) on the
device, without the reference's .cpu() round trip (:195-196, :278)."""
import ctypes

import torch

from . import _lib

# cfg.GRID_RCNN.CASCADE_MAPPING_OPTION.STAGE_MAPPING_RATIO (config.py:997)
STAGE_MAPPING_RATIO = (1.0, 0.5, 0.25)


def calc_sub_regions(grid_points, grid_size, whole_map_size):
    """pet/rcnn/modeling/grid_rcnn/loss.py:244-273: (x1, y1, x2, y2) of every grid point's half-size sub-region."""
    half = whole_map_size // 4 * 2
    res = []
    for i in range(grid_points):
        xi, yi = i // grid_size, i % grid_size
        sx = 0 if xi == 0 else half if xi == grid_size - 1 else max(int((xi / (grid_size - 1) - 0.25) * whole_map_size), 0)
        sy = 0 if yi == 0 else half if yi == grid_size - 1 else max(int((yi / (grid_size - 1) - 0.25) * whole_map_size), 0)
        res.append((sx, sy, sx + half, sy + half))
    return res


def grid_decode(grid_logits, boxes, sub_regions, mapping_ratio, return_scores=False):
    """grid_logits (R,P,h,w) PRE-sigmoid fp32, boxes (R,4) xyxy -> refined boxes (R,4) (un-clamped, as the reference:
    its clamp_ at :275-276 acts on a copy)."""
    _lib.require_cuda(grid_logits, "grid_logits")
    _lib.require_cuda(boxes, "boxes")
    logits = grid_logits.float().contiguous()
    boxes = boxes.float().contiguous()
    R, P, h, w = logits.shape
    assert boxes.shape == (R, 4)
    out = torch.empty((R, 4), dtype=torch.float32, device=logits.device)
    scores = torch.empty((R, P), dtype=torch.float32, device=logits.device) if return_scores else None
    if R:
        sub = (ctypes.c_int32 * (2 * P))(*[int(v) for s in sub_regions for v in (s[0], s[1])])
        with _lib.device_of(logits):
            _lib.check(_lib.lib().cpm_grid_decode(_lib.ptr(logits), _lib.ptr(boxes), R, P, h, w, sub,
                                                  float(mapping_ratio), _lib.ptr(out), _lib.ptr(scores),
                                                  _lib.stream_ptr(logits.device)))
    return (out, scores) if return_scores else out


class GridPostProcessor(object):
    """The decode half of the reference's GridPostProcessor (inference.py:127-143, :189-279): same constructor
    arguments, same get_boxes(proposals, grid_pred, is_train) signature -- but grid_pred are the LOGITS' sigmoid input
    or probabilities?  The reference applies .sigmoid() itself (:196), so grid_pred here too is pre-sigmoid."""

    def __init__(self, stage, grid_points=9, roi_feat_size=14, mapping_ratios=STAGE_MAPPING_RATIO, extend_roi=False):
        self.stage = stage
        self.grid_points = grid_points
        self.grid_size = int(round(grid_points ** 0.5))
        self.whole_map_size = roi_feat_size * 4
        self.sub_regions = calc_sub_regions(grid_points, self.grid_size, self.whole_map_size)
        self.mapping_ratio = 1.0 if extend_roi else mapping_ratios[stage]

    def get_boxes(self, proposals, grid_pred, is_train=False):
        det_bboxes = proposals.bbox
        assert det_bboxes.shape[0] > 0
        assert det_bboxes.shape[0] == grid_pred.shape[0]
        R, c, h, w = grid_pred.shape
        half_size = self.whole_map_size // 4 * 2
        assert h == w == half_size
        assert c == self.grid_points
        return grid_decode(grid_pred, det_bboxes, self.sub_regions, self.mapping_ratio)
