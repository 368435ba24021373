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
        # per-box image index and clip limits are functions of the (host-known) box counts: built on the host, one upload
        img_host = torch.repeat_interleave(torch.arange(B), torch.tensor(counts))
        lim_host = torch.tensor([[w - 1.0, h - 1.0, w - 1.0, h - 1.0] for (w, h) in image_shapes], dtype=torch.float32)[img_host]
        img_of_box = img_host.to(dev, non_blocking=True)
        # clip_to_image(remove_empty=False), bounding_box.py:294-299, for every image at once
        concat = torch.minimum(concat.clamp(min=0), lim_host.to(dev, non_blocking=True))
        # filter_results (:107-124): score > thresh and label != 0, labels = column index
        mask = class_prob > self.score_thresh
        mask[:, 0] = False
        nz = mask.nonzero()                                   # row-major (box, class): the reference's candidate order
        cand_boxes = concat[nz[:, 0]]
        cand_scores = class_prob[nz[:, 0], nz[:, 1]]
        cand_labels = nz[:, 1]
        cand_img = img_of_box[nz[:, 0]]
        if self.nms <= 0:                                     # boxlist_ml_nms returns its input (boxlist_ops.py:58-59)
            keep_sorted = torch.arange(nz.shape[0], device=dev)
            per_img = torch.bincount(cand_img, minlength=B).tolist()
        else:
            seg = (cand_img * num_classes + cand_labels).to(torch.int32)
            keep_buf, seg_counts, total = batched_nms(cand_boxes, cand_scores, seg, B * num_classes, self.nms, 0,
                                                      iou_flavor=_lib.IOU_ML_CUDA, sync=False)
            # the one host round trip after the candidate count: survivors in total and per image
            sizes = torch.cat([total.reshape(1), seg_counts.reshape(B, num_classes).sum(1)]).tolist()
            keep = keep_buf[:sizes[0]]
            per_img = sizes[1:]
            # ml_nms returns every image's survivors by decreasing score over all labels (ml_nms.cu:92-94,143-145): order by
            # (image, -score), ties by candidate index.  `keep` is grouped by (image, class) segment, i.e. by image already.
            keep = keep.sort().values
            order = torch.argsort(cand_scores[keep], descending=True, stable=True)
            keep = keep[order]
            keep_sorted = keep[torch.argsort(cand_img[keep], stable=True)]
        # one gather per field for the whole batch; the per-image BoxLists are views of it
        all_boxes, all_scores, all_labels = cand_boxes[keep_sorted], cand_scores[keep_sorted], cand_labels[keep_sorted]
        results = []
        pos = 0
        for i in range(B):
            n = int(per_img[i])
            bl = BoxList(all_boxes[pos:pos + n], image_shapes[i], mode="xyxy")
            bl.add_field("scores", all_scores[pos:pos + n])
            bl.add_field("labels", all_labels[pos:pos + n])
            results.append(bl)
            pos += n
        return results
