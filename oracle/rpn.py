"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's RPN proposal selection
(pet/rcnn/modeling/rpn/inference.py:68-172) on plain arrays, pinned against the reference itself by
tests/golden/rpn.npz (tests/test_oracle_cpu.py).  fp32 arithmetic in the reference's operation order:

  decode          BoxCoder.decode                 pet/rcnn/utils/box_coder.py:51-94
  clip            BoxList.clip_to_image           pet/utils/data/structures/bounding_box.py:294-304
  size test       remove_small_boxes              pet/utils/data/structures/boxlist_ops.py:104-118
  per-level NMS   boxlist_nms(max_proposals)      pet/utils/data/structures/boxlist_ops.py:10-32 (torchvision nms)
  selection       select_over_all_levels          inference.py:145-172
"""
import math

import numpy as np

F = np.float32


def decode(deltas, anchors, weights=(1.0, 1.0, 1.0, 1.0), clip=math.log(1000. / 16)):
    d, a = np.asarray(deltas, F).reshape(-1, 4), np.asarray(anchors, F).reshape(-1, 4)
    widths = (a[:, 2] - a[:, 0]) + F(1)
    heights = (a[:, 3] - a[:, 1]) + F(1)
    ctr_x = a[:, 0] + F(0.5) * widths
    ctr_y = a[:, 1] + F(0.5) * heights
    wx, wy, ww, wh = (F(w) for w in weights)
    dx, dy = d[:, 0] / wx, d[:, 1] / wy
    dw, dh = np.minimum(d[:, 2] / ww, F(clip)), np.minimum(d[:, 3] / wh, F(clip))
    pcx, pcy = dx * widths + ctr_x, dy * heights + ctr_y
    pw, ph = np.exp(dw).astype(F) * widths, np.exp(dh).astype(F) * heights
    out = np.empty_like(d)
    out[:, 0] = pcx - F(0.5) * pw
    out[:, 1] = pcy - F(0.5) * ph
    out[:, 2] = (pcx + F(0.5) * pw) - F(1)
    out[:, 3] = (pcy + F(0.5) * ph) - F(1)
    return out


def clip_and_size(boxes, im_w, im_h, min_size):
    b = boxes.copy()
    b[:, 0] = np.clip(b[:, 0], F(0), F(im_w - 1))
    b[:, 1] = np.clip(b[:, 1], F(0), F(im_h - 1))
    b[:, 2] = np.clip(b[:, 2], F(0), F(im_w - 1))
    b[:, 3] = np.clip(b[:, 3], F(0), F(im_h - 1))
    ws, hs = (b[:, 2] - b[:, 0]) + F(1), (b[:, 3] - b[:, 1]) + F(1)
    return b, (ws >= F(min_size)) & (hs >= F(min_size))


def sigmoid(x):
    return (F(1) / (F(1) + np.exp(-np.asarray(x, F)).astype(F))).astype(F)


def select(anchors, objectness, box_regression, image_sizes, pre_nms_top_n, post_nms_top_n, nms_thresh, min_size,
           fpn_post_nms_top_n, training, per_batch=True, weights=(1.0, 1.0, 1.0, 1.0), nms_fn=None, scores_are_probs=False):
    """anchors[l]: (N, A*H*W, 4) in the reference's (H, W, A) flattening; objectness[l]: (N, A, H, W) logits;
    box_regression[l]: (N, 4A, H, W).  Returns per image (boxes, scores).  nms_fn(boxes, scores, thr) -> keep indices."""
    if nms_fn is None:
        import oracle
        nms_fn = lambda b, s, t: oracle.nms(b, s, t)
    N = objectness[0].shape[0]
    per_image = [[] for _ in range(N)]
    for a, o, r in zip(anchors, objectness, box_regression):
        _, A, H, W = o.shape
        flat = o.reshape(N, A, 1, H, W).transpose(0, 3, 4, 1, 2).reshape(N, -1)          # permute_and_flatten
        prob = flat.astype(F) if scores_are_probs else sigmoid(flat)
        reg = r.reshape(N, A, 4, H, W).transpose(0, 3, 4, 1, 2).reshape(N, -1, 4)
        k = min(pre_nms_top_n, A * H * W)
        for i in range(N):
            order = np.argsort(-prob[i], kind="stable")[:k]
            boxes = decode(reg[i][order], a[i][order], weights)
            boxes, ok = clip_and_size(boxes, image_sizes[i][0], image_sizes[i][1], min_size)
            boxes, sc = boxes[ok], prob[i][order][ok]
            keep = nms_fn(boxes, sc, nms_thresh)
            if post_nms_top_n > 0:
                keep = keep[:post_nms_top_n]
            per_image[i].append((boxes[keep], sc[keep]))
    out = [(np.concatenate([b for b, _ in lv], 0), np.concatenate([s for _, s in lv], 0)) for lv in per_image]
    if len(objectness) > 1:
        if training and per_batch:
            allsc = np.concatenate([s for _, s in out])
            top = np.argsort(-allsc, kind="stable")[:min(fpn_post_nms_top_n, len(allsc))]
            mask = np.zeros(len(allsc), bool)
            mask[top] = True
            res, pos = [], 0
            for b, s in out:
                m = mask[pos:pos + len(s)]
                res.append((b[m], s[m]))
                pos += len(s)
            out = res
        else:
            res = []
            for b, s in out:
                top = np.argsort(-s, kind="stable")[:min(fpn_post_nms_top_n, len(s))]
                res.append((b[top], s[top]))
            out = res
    return out
