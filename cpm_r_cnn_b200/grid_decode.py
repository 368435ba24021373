"""Grid-point heat-map -> box decode (GridPostProcessor.get_boxes, grid_cascade_rcnn/inference.py:189-279) on the
device, without the reference's .cpu() round trip (:195-196, :278)."""
import ctypes

import torch

from . import _lib
from .structures import BoxList

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


def filter_not_gt(boxes, gt_boxes):
    """GridPostProcessor._filter_boxes (inference.py:281-290) without its .cpu(): every coordinate of a proposal that
    equals the same coordinate of some ground-truth box is replaced by -1, a proposal is kept when its four coordinates
    then sum to something positive (sum in the reference's order, ((x1 + y1) + x2) + y2).  Returns a bool mask."""
    hit = (boxes[:, None, :] == gt_boxes[None, :, :]).any(dim=1)
    last = torch.where(hit, torch.full_like(boxes, -1), boxes)
    s = ((last[:, 0] + last[:, 1]) + last[:, 2]) + last[:, 3]
    return s > 0


class GridPostProcessor(object):
    """The decode half of the reference's GridPostProcessor (inference.py:127-143, :189-279): same constructor
    arguments, same get_boxes(proposals, grid_pred, is_train) signature -- but grid_pred are the LOGITS' sigmoid input
    or probabilities?  The reference applies .sigmoid() itself (:196), so grid_pred here too is pre-sigmoid."""

    def __init__(self, stage, grid_points=9, roi_feat_size=14, mapping_ratios=STAGE_MAPPING_RATIO, extend_roi=False,
                 nms_on=True, fused_on=False, iou_helper=True, iou_helper_merge=True, stage_num=3):
        # the last five are the cfg.GRID_RCNN switches the reference reads at call time (config.py:935-943, :985); the
        # defaults are the published CPM models' values
        self.nms_on = nms_on
        self.fused_on = fused_on
        self.iou_helper = iou_helper
        self.iou_helper_merge = iou_helper_merge
        self.stage_num = stage_num
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

    def add_gt_proposals(self, proposal, target):
        """inference.py:292-298: the ground truth joins the refined proposals (fields: labels, objectness = 1)."""
        gt = target.bbox.to(proposal.bbox.device, proposal.bbox.dtype)
        out = BoxList(torch.cat([proposal.bbox, gt], 0), proposal.size, proposal.mode)
        extra = {"labels": target.get_field("labels").to(gt.device), "objectness": torch.ones(gt.shape[0], device=gt.device)}
        assert set(proposal.fields()) == set(extra), "proposals must carry exactly the fields labels and objectness"
        for k, v in extra.items():
            out.add_field(k, torch.cat([proposal.get_field(k), v.to(proposal.get_field(k).dtype)], 0))
        return out

    def forward(self, grid_logits, proposals, iou_logits=None, is_train=False, targets=None):
        """inference.py:145-187.  grid_logits: {'fused' | 'unfused': (R_total, P, h, w)}; proposals: list[BoxList].
        Training: proposals that coincide with ground truth are dropped (_filter_boxes), the rest are refined, the ground
        truth is appended.  Testing: every proposal is refined; on the last cascade stage the scores are multiplied (or
        replaced) by the IoU head's foreground probability.  The reference decodes image by image, each time through the
        host; here all images' boxes go through ONE decode launch."""
        grid_pred = grid_logits["fused"] if self.fused_on else grid_logits["unfused"]
        counts = [p.bbox.shape[0] for p in proposals]
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        if is_train:
            masks = [filter_not_gt(p.bbox, t.bbox.to(p.bbox.device, p.bbox.dtype)) for p, t in zip(proposals, targets)]
            kept = [p[m] for p, m in zip(proposals, masks)]
            sel = torch.cat([torch.nonzero(m).squeeze(1) + s for m, s in zip(masks, starts)])
            boxes = torch.cat([k.bbox for k in kept], 0)
            refined = (grid_decode(grid_pred[sel], boxes, self.sub_regions, self.mapping_ratio) if boxes.shape[0]
                       else boxes)
            out, pos = [], 0
            for k, t in zip(kept, targets):
                n = k.bbox.shape[0]
                k.bbox = refined[pos:pos + n]
                pos += n
                out.append(self.add_gt_proposals(k, t))
            return out
        boxes = torch.cat([p.bbox for p in proposals], 0)
        refined = grid_decode(grid_pred[:starts[-1]], boxes, self.sub_regions, self.mapping_ratio) if boxes.shape[0] else boxes
        last = self.iou_helper and self.stage == self.stage_num - 1
        for i, p in enumerate(proposals):
            if last:
                score, iou_score = p.get_field("scores"), iou_logits[:, 1]
                assert score.shape == iou_score.shape          # the reference indexes the whole tensor (:176-178)
                p.add_field("scores", score * iou_score if self.iou_helper_merge else iou_score)
            p.bbox = refined[starts[i]:starts[i + 1]]
        return list(proposals)
