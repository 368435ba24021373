"""BoxList NMS helpers with the reference's signatures.

    boxlist_nms(boxlist, nms_thresh, topk=0, score_field="scores", idxs=None)            pet/lib/ops/boxlist_ops.py:15-43
    boxlist_ml_nms(boxlist, nms_thresh, topk=0, score_field="scores", label_field="labels")              ... :46-67
    boxlist_nms_legacy(boxlist, nms_thresh, max_proposals=-1, score_field="scores")
                                                            pet/utils/data/structures/boxlist_ops.py:10-32 (RPN path)
    boxlist_ml_nms_legacy(boxlist, nms_thresh, max_proposals=-1, score_field, label_field)               ... :35-59
    batched_boxlist_nms(boxlists, nms_thresh, topk)   new: every BoxList of a batch in one launch
"""
import torch

from .nms import batched_nms, ml_nms as _box_ml_nms, nms as _box_nms


def boxlist_nms(boxlist, nms_thresh, topk=0, score_field="scores", idxs=None):
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    boxes = boxlist.bbox
    score = boxlist.get_field(score_field)
    if idxs is not None:
        # the reference offsets the coordinates by idx * (max + 1) (boxlist_ops.py:34-38), which perturbs the fp32 IoU;
        # segments give the same partition without touching the coordinates
        idxs = idxs.to(torch.int64)
        nseg = int(idxs.max().item()) + 1 if idxs.numel() else 1
        keep = batched_nms(boxes, score, idxs, nseg, nms_thresh)
        keep = keep[torch.argsort(score[keep], descending=True, stable=True)]
    else:
        keep = _box_nms(boxes, score, nms_thresh)
    if keep.size(0) > topk > 0:
        keep = keep[:topk]
    boxlist = boxlist[keep]
    return boxlist.convert(mode)


def boxlist_ml_nms(boxlist, nms_thresh, topk=0, score_field="scores", label_field="labels"):
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    boxes = boxlist.bbox
    scores = boxlist.get_field(score_field)
    labels = boxlist.get_field(label_field)
    keep = _box_ml_nms(boxes, scores, labels, nms_thresh, topk)
    boxlist = boxlist[keep]
    return boxlist.convert(mode)


def boxlist_nms_legacy(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = _box_nms(boxlist.bbox, boxlist.get_field(score_field), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    boxlist = boxlist[keep]
    return boxlist.convert(mode)


def boxlist_ml_nms_legacy(boxlist, nms_thresh, max_proposals=-1, score_field="scores", label_field="labels"):
    """The stale twin at pet/utils/data/structures/boxlist_ops.py:35-59 passes 4 arguments and float labels and would
    throw against ml_nms.h:16-21; this keeps its signature and does what it meant."""
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = _box_ml_nms(boxlist.bbox, boxlist.get_field(score_field), boxlist.get_field(label_field), nms_thresh,
                       max_proposals if max_proposals > 0 else 0)
    boxlist = boxlist[keep]
    return boxlist.convert(mode)


def batched_boxlist_nms(boxlists, nms_thresh, topk=0, score_field="scores"):
    """NMS of every BoxList (one per image / per image-level) in ONE launch; returns the list of filtered BoxLists,
    each identical to boxlist_nms_legacy(b, nms_thresh, topk)."""
    if nms_thresh <= 0 or not boxlists:
        return list(boxlists)
    modes = [b.mode for b in boxlists]
    xy = [b.convert("xyxy") for b in boxlists]
    sizes = [len(b) for b in xy]
    boxes = torch.cat([b.bbox for b in xy], dim=0)
    scores = torch.cat([b.get_field(score_field) for b in xy], dim=0)
    seg = torch.repeat_interleave(torch.arange(len(xy), dtype=torch.int32), torch.tensor(sizes)).to(boxes.device)
    keep, counts = batched_nms(boxes, scores, seg, len(xy), nms_thresh, topk, return_counts=True)
    counts = counts.tolist()
    out, pos, base = [], 0, 0
    for b, n, m, c in zip(xy, sizes, modes, counts):
        out.append(b[keep[pos:pos + c] - base].convert(m))
        pos += c
        base += n
    return out
