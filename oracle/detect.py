"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's detection post-processing
(CLSPostProcessor.forward / filter_results, pet/rcnn/modeling/grid_cascade_rcnn/inference.py:59-124) on plain arrays.
The label-gated NMS inside is oracle.nms(labels=...) (= _C.ml_nms, ml_nms.h:16-39), pinned separately by nms.npz."""
import numpy as np

F = np.float32


def softmax(x):
    x = np.asarray(x, F)
    e = np.exp(x - x.max(axis=-1, keepdims=True)).astype(F)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F)).astype(F)


def cls_postprocess(class_prob, boxes_per_image, image_sizes, score_thresh, nms_thresh, flavor=None):
    """class_prob (R_total, C) probabilities; boxes_per_image list of (R_i, 4); returns per image (boxes, scores, labels)
    ordered by decreasing score (ties: candidate order)."""
    import oracle
    flavor = oracle.FLAVOR_ML_CUDA if flavor is None else flavor
    out, pos = [], 0
    C = class_prob.shape[1]
    for b, (w, h) in zip(boxes_per_image, image_sizes):
        R = b.shape[0]
        prob = class_prob[pos:pos + R]
        pos += R
        bb = np.repeat(np.asarray(b, F), C, axis=0)                      # concat_boxes.repeat(1, C).reshape(-1, 4)
        bb[:, 0] = np.clip(bb[:, 0], F(0), F(w - 1))
        bb[:, 1] = np.clip(bb[:, 1], F(0), F(h - 1))
        bb[:, 2] = np.clip(bb[:, 2], F(0), F(w - 1))
        bb[:, 3] = np.clip(bb[:, 3], F(0), F(h - 1))
        sc = prob.reshape(-1)
        labels = np.tile(np.arange(C), R)
        ok = (sc > F(score_thresh)) & (labels != 0)
        bb, sc, labels = bb[ok], sc[ok], labels[ok]
        keep = oracle.nms(bb, sc, nms_thresh, labels=labels, topk=0, flavor=flavor) if nms_thresh > 0 else np.arange(len(sc))
        out.append((bb[keep], sc[keep], labels[keep]))
    return out
