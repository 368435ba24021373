"""Detection post-processing around ml_nms as one batched device pipeline: same interface as the
reference's CLSPostProcessor (pet/rcnn/modeling/grid_cascade_rcnn/inference.py:32-124).

The reference splits the class probabilities per image and, for every image, repeats the boxes per class, builds the
label arrays on the host with numpy (:113-118), masks and calls `_C.ml_nms` (mask D2H + host sweep inside,
ml_nms.cu:117-140).  Here all images go through one pass: softmax, clip, score gate and label build for the whole batch,
ONE batched NMS launch with segment = image x class (`_C.ml_nms` compares only equal labels, so a (image, class) segment is
exactly its comparison set), then the per-image ordering by decreasing score that ml_nms returns (ml_nms.cu:143-145).
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .nms import batched_nms
from .structures import BoxList


class CLSPostProcessor(nn.Module):
    """inference.py:32-58: CLSPostProcessor(score_thresh, nms); forward(x, boxes, rescore=False)."""

    def __init__(self, score_thresh, nms):
        super(CLSPostProcessor, self).__init__()
        self.score_thresh = score_thresh
        self.nms = nms

    def forward(self, x, boxes, rescore=False):
        """x: (R_total, num_classes) class logits of all images' proposals; boxes: list[BoxList] (one per image).
        rescore=True is the reference's rescoring branch (:62-76): scores <- scores^0.8 * prob[label]^0.2 in place."""
        class_prob = F.softmax(x, -1)
        if rescore:
            # :62-76 -- the geometric blend of the detection score with the rescoring head's probability of its label
            # (weights 0.8 / 0.2, 'pow' mode); every BoxList indexes the full probability table, as in the reference
            rows = torch.arange(class_prob.shape[0], device=class_prob.device)
            for bl in boxes:
                p_label = class_prob[rows, bl.get_field("labels")]
                bl.add_field("scores", bl.get_field("scores").pow(0.8) * p_label.pow(0.2))
            return boxes
        _lib.require_cuda(x, "class logits")
        dev = x.device
        num_classes = class_prob.shape[1]
        image_shapes = [b.size for b in boxes]
        counts = [len(b) for b in boxes]
        B = len(boxes)
        concat = torch.cat([b.bbox for b in boxes], dim=0).float()
        img_of_box = torch.repeat_interleave(torch.arange(B, device=dev), torch.tensor(counts, device=dev))
        # clip_to_image(remove_empty=False), bounding_box.py:294-299, for every image at once
        lim = torch.tensor([[w - 1.0, h - 1.0, w - 1.0, h - 1.0] for (w, h) in image_shapes], dtype=torch.float32,
                           device=dev)[img_of_box]
        concat = torch.minimum(concat.clamp(min=0), lim)
        # filter_results (:107-124): score > thresh and label != 0, labels = column index
        mask = class_prob > self.score_thresh
        mask[:, 0] = False
        nz = mask.nonzero()                                   # row-major (box, class): the reference's candidate order
        cand_boxes = concat[nz[:, 0]]
        cand_scores = class_prob[nz[:, 0], nz[:, 1]]
        cand_labels = nz[:, 1]
        cand_img = img_of_box[nz[:, 0]]
        results = []
        if self.nms <= 0:                                     # boxlist_ml_nms returns its input (boxlist_ops.py:58-59)
            keep_sorted, kept_img = torch.arange(nz.shape[0], device=dev), cand_img
        else:
            seg = (cand_img * num_classes + cand_labels).to(torch.int32)
            keep, _ = batched_nms(cand_boxes, cand_scores, seg, B * num_classes, self.nms, 0,
                                  iou_flavor=_lib.IOU_ML_CUDA, return_counts=True)
            # ml_nms returns every image's survivors by decreasing score over all labels (ml_nms.cu:92-94,143-145):
            # order by (image, -score), ties by candidate index
            keep = keep.sort().values
            order = torch.argsort(cand_scores[keep], descending=True, stable=True)
            keep = keep[order]
            keep_sorted = keep[torch.argsort(cand_img[keep], stable=True)]
            kept_img = cand_img[keep_sorted]
        per_img = torch.bincount(kept_img, minlength=B).tolist()
        pos = 0
        for i in range(B):
            idx = keep_sorted[pos:pos + per_img[i]]
            pos += per_img[i]
            bl = BoxList(cand_boxes[idx], image_shapes[i], mode="xyxy")
            bl.add_field("scores", cand_scores[idx])
            bl.add_field("labels", cand_labels[idx])
            results.append(bl)
        return results
