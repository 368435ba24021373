"""TEST INFRASTRUCTURE ONLY -- numpy restatement of boxlist_iou (pet/utils/data/structures/boxlist_ops.py:123-158) and
Matcher (pet/rcnn/utils/matcher.py:52-112), fp32 in the reference's operation order."""
import numpy as np

F = np.float32


def box_iou(b1, b2):
    b1, b2 = np.asarray(b1, F).reshape(-1, 4), np.asarray(b2, F).reshape(-1, 4)
    a1 = (b1[:, 2] - b1[:, 0] + F(1)) * (b1[:, 3] - b1[:, 1] + F(1))
    a2 = (b2[:, 2] - b2[:, 0] + F(1)) * (b2[:, 3] - b2[:, 1] + F(1))
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.maximum(rb - lt + F(1), F(0))
    inter = wh[:, :, 0] * wh[:, :, 1]
    return (inter / (a1[:, None] + a2[None, :] - inter)).astype(F)


def match(q, high, low, allow_low_quality=False):
    q = np.asarray(q, F)
    vals, idx = q.max(axis=0), q.argmax(axis=0).astype(np.int64)          # argmax: first maximum, like torch.max(dim=0)
    matches = idx.copy()
    matches[vals < F(low)] = -1
    matches[(vals >= F(low)) & (vals < F(high))] = -2
    if allow_low_quality:
        best = q.max(axis=1)
        upd = np.nonzero(q == best[:, None])[1]
        matches[upd] = idx[upd]
    return matches
