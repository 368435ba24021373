"""RPN proposal selection as one batched device pipeline: same interface as the reference's
RPNPostProcessor (pet/rcnn/modeling/rpn/inference.py:12-172).

The reference loops in Python over FPN levels (permute, sigmoid, top-k, two gathers per level) and, inside each level, over
images: clip -> remove_small_boxes (nonzero, host sync) -> boxlist_nms, i.e. 5 x B tiny NMS calls per iteration.  Here all
levels go through ONE top-k: cpm_rpn_flatten_objectness lays every (level, image)'s objectness out as a padded row in
permute_and_flatten order, torch applies sigmoid and top-k to the rows (the same values and the same per-row selection as
inference.py:84-88), cpm_rpn_select_decode fetches each winner's deltas and anchor straight from the head's (N, 4A, H, W)
output and decodes it (BoxCoder.decode + clip + size test), one cpm_nms_batched call (segments = level x image,
post_nms_top_n per segment) follows, then the cross-level selection.  The only host synchronisation is the read-back of
the per-segment keep counts that sizes the returned BoxLists.
"""
import ctypes
import math

import numpy as np
import torch
from torch import nn

from . import _lib
from .nms import batched_nms
from .structures import BoxList


class BoxCoder(object):
    """The two attributes of pet/rcnn/utils/box_coder.py the decode needs (weights, bbox_xform_clip)."""

    def __init__(self, weights=(1.0, 1.0, 1.0, 1.0), bbox_xform_clip=math.log(1000. / 16)):
        self.weights = weights
        self.bbox_xform_clip = bbox_xform_clip


def permute_and_flatten(layer, N, A, C, H, W):
    """pet/rcnn/utils/misc.py:6-10."""
    layer = layer.view(N, -1, C, H, W)
    layer = layer.permute(0, 3, 4, 1, 2)
    return layer.reshape(N, -1, C)


def rpn_decode(deltas, anchors, segments, segment_im_wh, weights, bbox_xform_clip, min_size):
    """(M,4) deltas + anchors -> decoded, clipped boxes (M,4) and output segments (trash ids for too-small boxes)."""
    _lib.require_cuda(deltas, "deltas")
    M = deltas.shape[0]
    S = segment_im_wh.shape[0]
    num_trash = max(1, (M + 32767) // 32768)
    deltas = deltas.float().contiguous()
    anchors = anchors.float().contiguous()
    segments = segments.to(torch.int32).contiguous()
    segment_im_wh = segment_im_wh.float().contiguous()
    boxes = torch.empty((M, 4), dtype=torch.float32, device=deltas.device)
    seg_out = torch.empty((M,), dtype=torch.int32, device=deltas.device)
    if M:
        w = (ctypes.c_float * 4)(*[float(v) for v in weights])
        with _lib.device_of(deltas):
            _lib.check(_lib.lib().cpm_rpn_decode(_lib.ptr(deltas), _lib.ptr(anchors), _lib.ptr(segments),
                                                 _lib.ptr(segment_im_wh), M, S, num_trash, w, float(bbox_xform_clip),
                                                 float(min_size), _lib.ptr(boxes), _lib.ptr(seg_out),
                                                 _lib.stream_ptr(deltas.device)))
    return boxes, seg_out, S + num_trash


class RPNPostProcessor(nn.Module):
    """inference.py:12-42: RPNPostProcessor(pre_nms_top_n, post_nms_top_n, nms_thresh, min_size, box_coder=None,
    fpn_post_nms_top_n=None, fpn_post_nms_per_batch=True); forward(anchors, objectness, box_regression, targets=None)."""

    def __init__(self, pre_nms_top_n, post_nms_top_n, nms_thresh, min_size, box_coder=None, fpn_post_nms_top_n=None,
                 fpn_post_nms_per_batch=True):
        super(RPNPostProcessor, self).__init__()
        self.pre_nms_top_n = pre_nms_top_n
        self.post_nms_top_n = post_nms_top_n
        self.nms_thresh = nms_thresh
        self.min_size = min_size
        if box_coder is None:
            box_coder = BoxCoder(weights=(1.0, 1.0, 1.0, 1.0))
        self.box_coder = box_coder
        if fpn_post_nms_top_n is None:
            fpn_post_nms_top_n = post_nms_top_n
        self.fpn_post_nms_top_n = fpn_post_nms_top_n
        self.fpn_post_nms_per_batch = fpn_post_nms_per_batch

    def add_gt_proposals(self, proposals, targets):
        """inference.py:44-66: ground-truth boxes join the proposals with objectness 1."""
        out = []
        for proposal, target in zip(proposals, targets):
            gt = target.bbox.to(proposal.bbox.device, proposal.bbox.dtype)
            merged = BoxList(torch.cat([proposal.bbox, gt], 0), proposal.size, proposal.mode)
            merged.add_field("objectness", torch.cat([proposal.get_field("objectness"),
                                                      torch.ones(gt.shape[0], device=gt.device)], 0))
            out.append(merged)
        return out

    def _level_candidates(self, anchors, objectness, box_regression):
        """One FPN level, all images at once (what inference.py:75-94 computes): objectness probabilities of the
        pre_nms_top_n best anchors per image, with their regression deltas and anchor boxes."""
        N, A, H, W = objectness.shape
        k = min(self.pre_nms_top_n, A * H * W)
        prob = permute_and_flatten(objectness, N, A, 1, H, W).view(N, -1).sigmoid()
        top_prob, top_idx = prob.topk(k, dim=1, sorted=True)
        pick = top_idx.unsqueeze(-1).expand(N, k, 4)
        deltas = permute_and_flatten(box_regression, N, A, 4, H, W).gather(1, pick)
        level_anchors = torch.cat([a.bbox for a in anchors], dim=0).reshape(N, -1, 4).gather(1, pick)
        return top_prob, deltas, level_anchors

    def _select_decode_all_levels(self, anchors, objectness, box_regression, im_wh):
        """Every level at once: padded objectness rows -> sigmoid -> one top-k -> gather + decode of the winners.
        Returns boxes (M,4), scores (M,), segments (M,) in (level, image, rank) order, and the segment count incl. trash."""
        L, N, dev = len(objectness), objectness[0].shape[0], objectness[0].device
        lv = _lib.RpnLevels()
        lv.num_levels, lv.num_images = L, N
        keep_alive, row, kmax, M = [], 0, 0, 0
        for l in range(L):
            o, b = objectness[l].contiguous(), box_regression[l].float().contiguous()
            _, A, H, W = o.shape
            per_level = [per_image[l].bbox for per_image in anchors]
            shared = all(t.data_ptr() == per_level[0].data_ptr() for t in per_level)      # AnchorGenerator shares them
            an = per_level[0] if shared else torch.stack(per_level, 0)
            an = an.float().contiguous()
            keep_alive += [o, b, an]
            lv.d_objectness[l], lv.d_regression[l], lv.d_anchors[l] = o.data_ptr(), b.data_ptr(), an.data_ptr()
            lv.anchors_per_image[l] = 0 if shared else 1
            lv.A[l], lv.HW[l] = A, H * W
            lv.k[l] = min(self.pre_nms_top_n, A * H * W)
            row, kmax, M = max(row, A * H * W), max(kmax, lv.k[l]), M + N * lv.k[l]
        lv.row = row
        rows = torch.empty((L * N, row), dtype=torch.float32, device=dev)
        num_trash = max(1, (M + 32767) // 32768)
        boxes = torch.empty((M, 4), dtype=torch.float32, device=dev)
        scores = torch.empty((M,), dtype=torch.float32, device=dev)
        seg_out = torch.empty((M,), dtype=torch.int32, device=dev)
        w = (ctypes.c_float * 4)(*[float(v) for v in self.box_coder.weights])
        with _lib.device_of(rows):
            st = _lib.stream_ptr(dev)
            _lib.check(_lib.lib().cpm_rpn_flatten_objectness(ctypes.byref(lv), _lib.ptr(rows), st))
            top_val, top_idx = rows.sigmoid_().topk(kmax, dim=1, sorted=True)
            _lib.check(_lib.lib().cpm_rpn_select_decode(ctypes.byref(lv), _lib.ptr(top_idx), _lib.ptr(top_val), kmax,
                                                        _lib.ptr(im_wh), num_trash, w, float(self.box_coder.bbox_xform_clip),
                                                        float(self.min_size), _lib.ptr(boxes), _lib.ptr(scores),
                                                        _lib.ptr(seg_out), st))
        return boxes, scores, seg_out, L * N + num_trash

    def forward(self, anchors, objectness, box_regression, targets=None):
        """anchors: list (images) of list (levels) of BoxList; objectness / box_regression: list (levels) of tensors
        (N, A, H, W) / (N, 4A, H, W).  Returns list[BoxList] with field "objectness" (inference.py:115-143)."""
        num_levels = len(objectness)
        N = objectness[0].shape[0]
        dev = objectness[0].device
        _lib.require_cuda(objectness[0], "objectness")
        image_sizes = [per_image[0].size for per_image in anchors]
        S = num_levels * N
        im_wh = torch.tensor([[float(w), float(h)] for (w, h) in image_sizes], dtype=torch.float32, device=dev)
        if num_levels <= _lib.CPM_MAX_LEVELS and all(o.dtype == torch.float32 for o in objectness):
            boxes, scores, seg_out, nseg_total = self._select_decode_all_levels(anchors, objectness, box_regression, im_wh)
        else:
            scores, deltas, boxes_a, segs = [], [], [], []
            for l, (a, o, b) in enumerate(zip(list(zip(*anchors)), objectness, box_regression)):
                s, d, an = self._level_candidates(a, o, b)
                k = s.shape[1]
                scores.append(s.reshape(-1))
                deltas.append(d.reshape(-1, 4))
                boxes_a.append(an.reshape(-1, 4))
                segs.append((l * N + torch.arange(N, device=dev, dtype=torch.int32))[:, None].expand(N, k).reshape(-1))
            scores, deltas, boxes_a, segs = torch.cat(scores), torch.cat(deltas), torch.cat(boxes_a), torch.cat(segs)
            boxes, seg_out, nseg_total = rpn_decode(deltas, boxes_a, segs, im_wh.repeat(num_levels, 1), self.box_coder.weights,
                                                    self.box_coder.bbox_xform_clip, self.min_size)
        topk = self.post_nms_top_n if self.post_nms_top_n > 0 else 0      # boxlist_nms(max_proposals=...), :106-111
        keep, counts, _ = batched_nms(boxes, scores, seg_out, nseg_total, self.nms_thresh, topk, sync=False)
        counts_h = counts[:S].cpu().tolist()                              # host sync: sizes of the results
        boxlists = self._assemble(keep, counts_h, boxes, scores, image_sizes, num_levels, N)
        if self.training and targets is not None:
            boxlists = self.add_gt_proposals(boxlists, targets)
        return boxlists

    def _assemble(self, keep, counts_h, boxes, scores, image_sizes, num_levels, N):
        """The kept boxes of every (level, image) segment -> one BoxList per image: cat_boxlist order (levels in order) and,
        for a pyramid, select_over_all_levels (inference.py:145-172), for all images at once: index arithmetic on the host
        from the segment counts, a handful of gathers on the device, the per-image results are slices."""
        dev = boxes.device
        offs = np.concatenate([[0], np.cumsum(counts_h)]).astype(np.int64)
        pieces = [np.arange(offs[l * N + i], offs[l * N + i + 1]) for i in range(N) for l in range(num_levels)]
        order = np.concatenate(pieces) if pieces else np.zeros((0,), np.int64)
        per_img = [int(sum(counts_h[l * N + i] for l in range(num_levels))) for i in range(N)]
        kidx = keep[torch.from_numpy(order).to(dev)]                      # image-major, level-minor: cat_boxlist order
        kb, ks = boxes[kidx], scores[kidx]
        starts = np.concatenate([[0], np.cumsum(per_img)]).astype(np.int64)
        total = int(starts[-1])
        if num_levels > 1 and total > 0:
            if self.training and self.fpn_post_nms_per_batch:
                # the fpn_post_nms_top_n best of the WHOLE batch survive, every image keeps its own in their current order
                chosen = torch.zeros((total,), dtype=torch.bool, device=dev)
                chosen[ks.topk(min(self.fpn_post_nms_top_n, total), dim=0, sorted=True).indices] = True
                chosen_h = chosen.cpu().numpy()                           # host sync: who survived
                sel = np.nonzero(chosen_h)[0]
                per_img = [int(chosen_h[starts[i]:starts[i + 1]].sum()) for i in range(N)]
                sel_d = torch.from_numpy(sel).to(dev)
                kb, ks = kb[sel_d], ks[sel_d]
            else:
                # every image keeps its fpn_post_nms_top_n best, by decreasing objectness: one sort over (image, -score)
                # (key = image in the high word, the score's bit pattern -- monotone for non-negative floats -- in the low one)
                hi = torch.from_numpy(np.repeat(np.arange(N - 1, -1, -1, dtype=np.int64) << 32, per_img)).to(dev)
                srt = (hi | ks.contiguous().view(torch.int32).to(torch.int64)).sort(descending=True).indices
                take = [min(self.fpn_post_nms_top_n, c) for c in per_img]
                pick = np.concatenate([np.arange(starts[i], starts[i] + take[i]) for i in range(N)])
                fin = srt[torch.from_numpy(pick).to(dev)] if sum(take) < total else srt
                kb, ks = kb[fin], ks[fin]
                per_img = take
            starts = np.concatenate([[0], np.cumsum(per_img)]).astype(np.int64)
        boxlists = []
        for i in range(N):
            bl = BoxList(kb[int(starts[i]):int(starts[i + 1])], image_sizes[i], mode="xyxy")
            bl.add_field("objectness", ks[int(starts[i]):int(starts[i + 1])])
            boxlists.append(bl)
        return boxlists

    def select_over_all_levels(self, boxlists):
        """inference.py:145-172.  Training (with fpn_post_nms_per_batch, the Detectron convention): the
        fpn_post_nms_top_n best proposals of the WHOLE batch survive, each image keeps its own in their current order.
        Testing: every image keeps its fpn_post_nms_top_n best, ordered by decreasing objectness."""
        scores = [b.get_field("objectness") for b in boxlists]
        if self.training and self.fpn_post_nms_per_batch:
            flat = torch.cat(scores, dim=0)
            chosen = torch.zeros_like(flat, dtype=torch.bool)
            chosen[flat.topk(min(self.fpn_post_nms_top_n, flat.numel()), dim=0, sorted=True).indices] = True
            return [b[m] for b, m in zip(boxlists, chosen.split([len(b) for b in boxlists]))]
        return [b[s.topk(min(self.fpn_post_nms_top_n, s.numel()), dim=0, sorted=True).indices]
                for b, s in zip(boxlists, scores)]
