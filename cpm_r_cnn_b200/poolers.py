"""LevelMapper + Pooler with the reference's interface (pet/rcnn/utils/poolers.py:9-40, :43-132).

The reference walks the FPN levels in Python: nonzero() (host sync) -> gather rois -> ROIAlign -> index_put, once per
level (poolers.py:127-130).  Here the whole pyramid is ONE kernel launch: the FPN level of every RoI (Eqn. 1 of the FPN
paper, poolers.py:35-40) is evaluated inside the RoIAlign kernel and each RoI's block lands at its own row of the
output; the backward is one launch that produces the dense gradient of every level.
"""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import _lib
from .roi_align import INTERPOLATION_METHOD, ROIAlign, _float_function, _is_nhwc, pooler_backward, pooler_forward


class LevelMapper(object):
    """poolers.py:9-40.  __call__(boxlists) -> int64 level index per box, computed on the device."""

    def __init__(self, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
        self.k_min = k_min
        self.k_max = k_max
        self.s0 = canonical_scale
        self.lvl0 = canonical_level
        self.eps = eps

    def c_struct(self):
        return _lib.make_mapper(self.k_min, self.k_max, self.s0, self.lvl0, self.eps)

    def __call__(self, boxlists):
        boxes = torch.cat([b.bbox for b in boxlists], dim=0)
        return self.map_boxes(boxes)

    def map_boxes(self, boxes):
        """(K,4) xyxy boxes -> (K,) int64 levels (area with the +1 convention, bounding_box.py:306-310)."""
        _lib.require_cuda(boxes, "boxes")
        K = boxes.shape[0]
        rois = torch.cat([boxes.new_zeros((K, 1)), boxes.float()], dim=1).contiguous()
        out = torch.empty((K,), dtype=torch.int64, device=boxes.device)
        if K:
            m = self.c_struct()
            with _lib.device_of(rois):
                _lib.check(_lib.lib().cpm_level_map(_lib.ptr(rois), K, ctypes.byref(m), _lib.ptr(out),
                                                    _lib.stream_ptr(boxes.device)))
        return out


class _PyramidROIAlign(Function):
    """Multi-level RoIAlign as one autograd node: forward(rois, cfg, *levels) -> (K,C,PH,PW)."""

    @staticmethod
    def forward(ctx, rois, cfg, *levels):
        ctx.save_for_backward(rois)
        ctx.cfg = cfg
        ctx.shapes = [tuple(t.shape) for t in levels]
        ctx.nchw_input = not all(_is_nhwc(t) for t in levels)
        output_size, scales, sampling_ratio, aligned, interp, mapper, channels_last = cfg
        return pooler_forward(list(levels), scales, rois, output_size, sampling_ratio, aligned, interp, mapper,
                              channels_last=channels_last)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, = ctx.saved_tensors
        output_size, scales, sampling_ratio, aligned, interp, mapper, _ = ctx.cfg
        grads = pooler_backward(grad_output, ctx.shapes, scales, rois, output_size, sampling_ratio, aligned, interp,
                                mapper, nchw_grad=ctx.nchw_input)
        return (None, None) + tuple(grads)


class Pooler(nn.Module):
    """poolers.py:43-132: Pooler(method, output_size, scales, sampling_ratio, rotated=False, interpolation)."""

    def __init__(self, method, output_size, scales, sampling_ratio, rotated=False, interpolation="bilinear"):
        assert method in {'ROIPool', 'ROIAlign', 'ROIAlignV2', 'ROIAlignRotated'}, \
            'Unknown pooling method: {}'.format(method)
        if method in ('ROIPool', 'ROIAlignRotated') or rotated:
            # out of the hot path: no CPM config selects them (config.py:877)
            raise NotImplementedError("cpm_ops implements the ROIAlign / ROIAlignV2 poolers only")
        super(Pooler, self).__init__()
        self.output_size = _pair(output_size)
        self.scales = [float(s) for s in scales]
        self.sampling_ratio = sampling_ratio
        self.aligned = "V2" in method
        self.interpolation = interpolation
        self.poolers = nn.ModuleList([
            ROIAlign(self.output_size, spatial_scale=s, sampling_ratio=sampling_ratio, aligned=self.aligned,
                     interpolation=interpolation) for s in scales])
        lvl_min = -torch.log2(torch.tensor(scales[0], dtype=torch.float32)).item()
        lvl_max = -torch.log2(torch.tensor(scales[-1], dtype=torch.float32)).item()
        self.map_levels = LevelMapper(lvl_min, lvl_max)
        # Extension (not in the reference): set to torch.channels_last for heads whose convolutions run in channels_last
        # (the 14x14 grid head); the pooled tensor then has channels_last strides and its gradient is read in place.
        self.pooled_memory_format = torch.contiguous_format
        self.native_bf16 = False      # extension: feed bf16 feature maps to the bf16x8 kernel instead of casting to fp32

    def convert_to_roi_format(self, boxes):
        """poolers.py:90-101: (K,5) [image index, x1, y1, x2, y2] in the boxes' dtype."""
        # built on the boxes' device with fill kernels only: a host->device copy here, however small, queues behind
        # whatever large upload the caller has in flight on the copy engine and stalls the pooler
        return torch.cat([torch.cat([b.bbox.new_full((len(b), 1), float(i)), b.bbox], dim=1) for i, b in enumerate(boxes)],
                         dim=0)

    def forward(self, x, boxes):
        """x: list[Tensor] feature maps (one per level, extra levels are ignored like zip() does at poolers.py:127);
        boxes: list[BoxList].  Returns (K, C, PH, PW) in x[0].dtype."""
        num_levels = len(self.poolers)
        rois = self.convert_to_roi_format(boxes)
        if num_levels == 1:
            self.poolers[0].pooled_memory_format = self.pooled_memory_format
            self.poolers[0].native_bf16 = self.native_bf16
            return self.poolers[0](x[0], rois)
        result_dtype = x[0].dtype             # poolers.py:119-131: the result is allocated in, and cast back to, x[0].dtype
        levels = [_float_function(t, self.native_bf16) for t in list(x)[:num_levels]]
        rois = _float_function(rois)
        rois = rois.to(torch.float32 if levels[0].dtype == torch.bfloat16 else levels[0].dtype)
        cfg = (self.output_size, self.scales[:len(levels)], self.sampling_ratio, self.aligned,
               INTERPOLATION_METHOD[self.interpolation], self.map_levels.c_struct(),
               self.pooled_memory_format == torch.channels_last)
        out = _PyramidROIAlign.apply(rois, cfg, *levels)
        return out if out.dtype == result_dtype else out.to(result_dtype)
