"""Box IoU matrix + Matcher on the device, with the reference's names and signatures:

    boxlist_iou(boxlist1, boxlist2) -> (N, M)      pet/utils/data/structures/boxlist_ops.py:123-158
    Matcher(high_threshold, low_threshold, allow_low_quality_matches=False)(match_quality_matrix) -> int64 (N,)
                                                    pet/rcnn/utils/matcher.py:4-112

The fg/bg sampler that follows them (balanced_positive_negative_sampler.py) draws torch.randperm on the device and is left
as it is: its output is random by construction, there is nothing to pin.
"""
import torch

from . import _lib


def boxlist_iou(boxlist1, boxlist2):
    if boxlist1.size != boxlist2.size:
        raise RuntimeError("boxlists should have same image size, got {}, {}".format(boxlist1, boxlist2))
    b1, b2 = boxlist1.bbox, boxlist2.bbox
    _lib.require_cuda(b1, "boxlist1")
    _lib.require_cuda(b2, "boxlist2")
    b1, b2 = b1.float().contiguous(), b2.float().contiguous()
    N, M = b1.shape[0], b2.shape[0]
    out = torch.empty((N, M), dtype=torch.float32, device=b1.device)
    if N and M:
        with _lib.device_of(b1):
            _lib.check(_lib.lib().cpm_box_iou(_lib.ptr(b1), _lib.ptr(b2), N, M, _lib.ptr(out), _lib.stream_ptr(b1.device)))
    return out


class Matcher(object):
    BELOW_LOW_THRESHOLD = -1
    BETWEEN_THRESHOLDS = -2

    def __init__(self, high_threshold, low_threshold, allow_low_quality_matches=False):
        assert low_threshold <= high_threshold
        self.high_threshold = high_threshold
        self.low_threshold = low_threshold
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix):
        """(M ground truth, N predictions) qualities -> int64 (N,): matched ground-truth index, -1 or -2."""
        if match_quality_matrix.numel() == 0:
            if match_quality_matrix.shape[0] == 0:
                raise ValueError("No ground-truth boxes available for one of the images during training")
            raise ValueError("No proposal boxes available for one of the images during training")
        _lib.require_cuda(match_quality_matrix, "match_quality_matrix")
        q = match_quality_matrix.float().contiguous()
        M, N = q.shape
        matches = torch.empty((N,), dtype=torch.int64, device=q.device)
        with _lib.device_of(q):
            L = _lib.lib()
            nbytes = int(L.cpm_matcher_workspace_bytes(M, N))
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=q.device)
            _lib.check(L.cpm_matcher(_lib.ptr(q), M, N, float(self.high_threshold), float(self.low_threshold),
                                     int(bool(self.allow_low_quality_matches)), _lib.ptr(matches), _lib.ptr(ws), nbytes,
                                     _lib.stream_ptr(q.device)))
        return matches
