"""Drop-in for the three entry points of the reference's pybind module `pet.lib.ops._C` that sit on the hot path
(csrc/vision.cpp:21-22,32): the same names, positional signatures and return conventions, bound to libcpm_ops.so.

    # pet/lib/ops/roi_align.py, pet/lib/ops/nms.py:   from cpm_r_cnn_b200 import compat_C as _C

INTEGRATION.md level 2: the reference keeps ALL of its Python (its autograd Function, its Pooler loop, its BoxList code)
and only `_C` changes.  Executed by tests/test_gpu_parity.py::test_level2_C_module_against_the_reference_build.
"""
import torch

from .nms import ml_nms as _ml_nms
from .roi_align import pooler_backward, pooler_forward


def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio, aligned, interpolation_method):
    """ROIAlign.h:57-65 -> (K, C, PH, PW) in input's dtype.  An NCHW map is staged to NHWC once per tensor (cached), a
    channels_last map is read in place."""
    return pooler_forward([input], [spatial_scale], rois, (pooled_height, pooled_width), sampling_ratio, bool(aligned),
                          int(interpolation_method), None)


def roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width,
                       sampling_ratio, aligned, interpolation_method):
    """ROIAlign.h:98-110 -> dense (B, C, H, W) gradient, NCHW-contiguous as the reference returns it
    (ROIAlign_cuda.cu:451-452); deterministic where the reference's atomicAdd scatter is not."""
    return pooler_backward(grad, [(batch_size, channels, height, width)], [spatial_scale], rois,
                           (pooled_height, pooled_width), sampling_ratio, bool(aligned), int(interpolation_method), None,
                           nchw_grad=True)[0]


def ml_nms(dets, scores, labels, iou_threshold, topk):
    """ml_nms.h:16-21 -> int64 keep indices by decreasing score."""
    return _ml_nms(dets, scores, labels, iou_threshold, topk)
