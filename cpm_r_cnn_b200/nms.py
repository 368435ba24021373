"""NMS op layer with the reference's names (pet/lib/ops/nms.py:10-11):

    nms(boxes, scores, iou_threshold)                      <- torchvision.ops.nms (pet/lib/ops/nms.py:2,10)
    ml_nms(boxes, scores, labels, iou_threshold, topk)     <- _C.ml_nms (ml_nms.h:16-21)
    batched_nms(boxes, scores, segments, num_segments, iou_threshold, topk_per_segment=0)
        one call for all (image x FPN level) or (image x class) segments -- replaces the Python loops at
        rpn/inference.py:102-113 and grid_cascade_rcnn/inference.py:91-97.

Everything (sort, IoU tests, suppression sweep, compaction) runs on the device; the only host round trip is reading the
number of kept boxes to size the returned tensor.  Inputs are computed in fp32 (apex.amp.float_function semantics).
"""
import torch

from . import _lib


def _ws(nbytes, device):
    return torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=device)


def _prep(boxes, scores):
    _lib.require_cuda(boxes, "boxes")
    _lib.require_cuda(scores, "scores")
    if boxes.dim() != 2 or boxes.size(1) != 4:
        raise RuntimeError("cpm_ops: boxes should be a 2d tensor of shape (N, 4), got %s" % (tuple(boxes.shape),))
    if scores.dim() != 1 or scores.size(0) != boxes.size(0):
        raise RuntimeError("cpm_ops: boxes and scores should have the same number of elements")
    return boxes.float().contiguous(), scores.float().contiguous()


def nms(boxes, scores, iou_threshold, iou_flavor=_lib.IOU_TV_CUDA):
    """keep indices (int64), sorted by decreasing score; box j is dropped when a kept higher-scored box i has
    IoU(i, j) > iou_threshold (strict, areas without +1)."""
    boxes, scores = _prep(boxes, scores)
    N = boxes.shape[0]
    dev = boxes.device
    keep = torch.empty((N,), dtype=torch.int64, device=dev)
    if N == 0:
        return keep
    count = torch.empty((1,), dtype=torch.int64, device=dev)
    with _lib.device_of(boxes):
        L = _lib.lib()
        nbytes = L.cpm_nms_workspace_bytes(N)
        ws = _ws(nbytes, dev)
        _lib.check(L.cpm_nms(_lib.ptr(boxes), _lib.ptr(scores), None, N, float(iou_threshold), 0, iou_flavor,
                             _lib.ptr(keep), _lib.ptr(count), _lib.ptr(ws), nbytes, _lib.stream_ptr(dev)))
    return keep[:int(count.item())]


def ml_nms(boxes, scores, labels, iou_threshold, topk=0, iou_flavor=_lib.IOU_ML_CUDA):
    """Per-label NMS (ml_nms.h:16-21): a pair is only compared when labels match; result ordered by decreasing score
    over all labels; topk > 0 keeps the topk best survivors (ml_nms.cu:134)."""
    boxes, scores = _prep(boxes, scores)
    _lib.require_cuda(labels, "labels")
    N = boxes.shape[0]
    dev = boxes.device
    keep = torch.empty((N,), dtype=torch.int64, device=dev)
    if N == 0:
        return keep                              # ml_nms.h:24-27
    labels = labels.to(torch.int64).contiguous()
    count = torch.empty((1,), dtype=torch.int64, device=dev)
    with _lib.device_of(boxes):
        L = _lib.lib()
        nbytes = L.cpm_nms_workspace_bytes(N)
        ws = _ws(nbytes, dev)
        _lib.check(L.cpm_nms(_lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(labels), N, float(iou_threshold), int(topk),
                             iou_flavor, _lib.ptr(keep), _lib.ptr(count), _lib.ptr(ws), nbytes, _lib.stream_ptr(dev)))
    return keep[:int(count.item())]


def batched_nms(boxes, scores, segments, num_segments, iou_threshold, topk_per_segment=0,
                iou_flavor=_lib.IOU_TV_CUDA, return_counts=False, sync=True):
    """Independent NMS problems in one launch.  segments: int tensor (N,) in [0, num_segments).
    Returns kept input indices grouped by segment (ascending id), decreasing score inside a segment; with
    return_counts also the (num_segments,) kept counts.  sync=False returns (keep_buffer, counts, total) device
    tensors without any host synchronisation (keep_buffer[:total] is valid)."""
    boxes, scores = _prep(boxes, scores)
    _lib.require_cuda(segments, "segments")
    N = boxes.shape[0]
    dev = boxes.device
    keep = torch.empty((N,), dtype=torch.int64, device=dev)
    counts = torch.zeros((num_segments,), dtype=torch.int64, device=dev)
    total = torch.zeros((1,), dtype=torch.int64, device=dev)
    if N:
        segments = segments.to(torch.int32).contiguous()
        with _lib.device_of(boxes):
            L = _lib.lib()
            nbytes = L.cpm_nms_batched_workspace_bytes(N, num_segments)
            ws = _ws(nbytes, dev)
            _lib.check(L.cpm_nms_batched(_lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(segments), N, int(num_segments),
                                         float(iou_threshold), int(topk_per_segment), iou_flavor, _lib.ptr(keep),
                                         _lib.ptr(counts), _lib.ptr(total), _lib.ptr(ws), nbytes, _lib.stream_ptr(dev)))
    if not sync:
        return keep, counts, total
    keep = keep[:int(total.item())]
    return (keep, counts) if return_counts else keep
