"""RoIAlign op layer: same names and signatures as the reference's pet/lib/ops/roi_align.py
(_ROIAlign :14-60, roi_align :63, ROIAlign :66-95), bound to libcpm_ops.so instead of pet.lib.ops._C.

Layout: the hot kernels read NHWC.  A torch.channels_last (B,C,H,W) tensor is passed through zero-copy; an
NCHW-contiguous fp32 map is staged once per call with cpm_layout_convert (pure streaming; see stage_nhwc to do it
once for several poolers); fp64 / nearest interpolation use the reference-shaped generic kernel on NCHW directly.
The dense input gradient is produced in NHWC and returned as a channels_last-strided (B,C,H,W) tensor.
"""
import ctypes
import os

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import _lib

INTERPOLATION_METHOD = {"bilinear": 0, "nearest": 1}

# "deterministic" (atomic-free tile-owner gather) or "atomic" (red.global.add scatter, the reference's semantics)
BACKWARD_MODE = os.environ.get("CPM_ROI_ALIGN_BACKWARD", "deterministic")


def _is_nhwc(t):
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


class _StagingCache(object):
    """NHWC copies of NCHW feature maps, shared by the poolers of one iteration.

    The reference's FPN emits NCHW maps (FPN.py:96-121) and a CPM iteration runs five Poolers over the same maps (box
    head, three grid stages, re-score head): staging the pyramid once instead of once per Pooler removes ~366 MB of
    traffic per extra call.  An entry is valid only for the very same tensor OBJECT (weak reference) at the same
    in-place version, so a new iteration's maps -- even at the same address -- are never served a stale copy."""

    def __init__(self, capacity=16):
        self.capacity = capacity
        self.entries = []                # (weakref to source, version, staged), most recent last
        self.hits = self.misses = 0

    def lookup(self, x):
        """The cached NHWC copy of tensor object `x` at its current version, or None (dead / stale entries are dropped)."""
        alive = []
        found = None
        for ref, ver, staged in self.entries:
            src = ref()
            if src is None:
                continue
            if src is x and ver == x._version and found is None:
                found = staged
            elif src is x:
                continue                 # written in place since it was staged: drop
            alive.append((ref, ver, staged))
        self.entries = alive
        if found is not None:
            self.hits += 1
        return found

    def put(self, x, staged):
        import weakref
        self.misses += 1
        self.entries.append((weakref.ref(x), x._version, staged))
        if len(self.entries) > self.capacity:
            self.entries = self.entries[-self.capacity:]

    def get(self, x):
        staged = self.lookup(x)
        if staged is None:
            staged = _stage_uncached(x)
            self.put(x, staged)
        return staged

    def clear(self):
        self.entries = []


STAGING_CACHE = _StagingCache()


def _stage_uncached(x):
    if x.dtype not in _lib.DTYPES:
        return x.contiguous(memory_format=torch.channels_last)
    src = x.contiguous()
    B, C, H, W = src.shape
    dst = torch.empty((B, C, H, W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    with _lib.device_of(src):
        _lib.check(_lib.lib().cpm_layout_convert(_lib.ptr(src), _lib.ptr(dst), B, C, H, W, _lib.DTYPES[x.dtype],
                                                 _lib.NHWC, _lib.stream_ptr(x.device)))
    return dst


def stage_nhwc(x, cache=True):
    """(B,C,H,W) any strides -> the same values as a channels_last tensor (zero-copy when it already is; otherwise one
    streaming transpose, remembered for the other poolers that read the same tensor in this iteration)."""
    _lib.require_cuda(x, "input")
    if _is_nhwc(x):
        return x
    if cache:
        return STAGING_CACHE.get(x)
    return _stage_uncached(x)


def make_pyramid(levels, scales, layout):
    p = _lib.Pyramid()
    B, C = levels[0].shape[0], levels[0].shape[1]
    p.num_levels, p.batch, p.channels = len(levels), B, C
    p.dtype, p.layout = _lib.DTYPES[levels[0].dtype], layout
    for i, (t, s) in enumerate(zip(levels, scales)):
        assert t.shape[0] == B and t.shape[1] == C and t.dtype == levels[0].dtype and t.device == levels[0].device
        p.d_level[i] = t.data_ptr()
        p.height[i], p.width[i] = t.shape[2], t.shape[3]
        p.spatial_scale[i] = float(s)
    return p


def stage_pyramid_nhwc(levels, cache=True):
    """NHWC copies of all levels: cached copies are reused, the missing fp32 levels are staged in ONE launch
    (cpm_layout_convert_pyramid)."""
    out = [None] * len(levels)
    todo = []
    for i, t in enumerate(levels):
        _lib.require_cuda(t, "input")
        if _is_nhwc(t):
            out[i] = t
        else:
            hit = STAGING_CACHE.lookup(t) if cache else None
            if hit is not None:
                out[i] = hit
            else:
                todo.append(i)
    fused = [i for i in todo if levels[i].dtype == torch.float32 and levels[i].is_contiguous()
             and levels[i].shape[:2] == levels[todo[0]].shape[:2]]
    if len(fused) > 1:
        srcs = [levels[i] for i in fused]
        dsts = [torch.empty(t.shape, dtype=t.dtype, device=t.device, memory_format=torch.channels_last) for t in srcs]
        ones = [1.0] * len(srcs)
        ps, pd = make_pyramid(srcs, ones, _lib.NCHW), make_pyramid(dsts, ones, _lib.NHWC)
        with _lib.device_of(srcs[0]):
            _lib.check(_lib.lib().cpm_layout_convert_pyramid(ctypes.byref(ps), ctypes.byref(pd),
                                                             _lib.stream_ptr(srcs[0].device)))
        for i, d in zip(fused, dsts):
            out[i] = d
            if cache:
                STAGING_CACHE.put(levels[i], d)
    for i in todo:
        if out[i] is None:
            out[i] = stage_nhwc(levels[i], cache)
    return out


def _prepare_levels(levels, interpolation):
    """-> (tensors kept alive, layout).  fp32 bilinear goes to NHWC (staging NCHW inputs); the rest stays NCHW."""
    dt = levels[0].dtype
    if (dt == torch.float32 and interpolation == 0 and levels[0].shape[1] % 4 == 0) or dt == torch.bfloat16:
        return stage_pyramid_nhwc(levels), _lib.NHWC
    return [t.contiguous() for t in levels], _lib.NCHW


def pooler_forward(levels, scales, rois, output_size, sampling_ratio, aligned, interpolation, mapper, impl=_lib.FWD_AUTO,
                   channels_last=False):
    """One launch for the whole pyramid: list[(B,C,H_l,W_l)] + (K,5) rois -> (K,C,PH,PW).

    channels_last=True returns the same logical tensor with torch.channels_last strides (memory (K,PH,PW,C)): what a
    channels_last conv head consumes directly; the kernel then stores channel vectors straight from registers."""
    for t in levels:
        _lib.require_cuda(t, "input")
    _lib.require_cuda(rois, "rois")
    x0 = levels[0]
    bf16 = x0.dtype == torch.bfloat16      # native bf16 storage (extension): rois stay fp32, arithmetic is fp32
    if (rois.dtype != (torch.float32 if bf16 else x0.dtype)) or rois.device != x0.device:
        # at::checkAllSameGPU / checkAllSameType, ROIAlign_cuda.cu:381-382
        raise RuntimeError("cpm_ops: input and rois must have the same dtype and device (got %s/%s, %s/%s)"
                           % (x0.dtype, x0.device, rois.dtype, rois.device))
    if x0.dtype not in (torch.float32, torch.float64, torch.bfloat16):
        raise RuntimeError("cpm_ops: roi_align supports float32, float64 and (native, opt-in) bfloat16 inputs (got %s)"
                           % x0.dtype)
    if bf16 and (interpolation != 0 or x0.shape[1] % 8 != 0):
        raise RuntimeError("cpm_ops: the bf16 path needs bilinear interpolation and C % 8 == 0")
    ph, pw = output_size
    K, C = rois.shape[0], x0.shape[1]
    out = torch.empty((K, C, ph, pw), dtype=x0.dtype, device=x0.device,
                      memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    if K == 0:
        return out
    staged, layout = _prepare_levels(levels, interpolation)
    pyr = make_pyramid(staged, scales, layout)
    rois = rois.contiguous()
    args = (ctypes.byref(pyr), _lib.ptr(rois), K, ph, pw, int(sampling_ratio), int(bool(aligned)), interpolation,
            ctypes.byref(mapper) if mapper is not None else None, None, impl)
    with _lib.device_of(x0):
        L = _lib.lib()
        if channels_last:
            rc = L.cpm_roi_align_forward_ex(*args, _lib.POOLED_KHWC, _lib.ptr(out), _lib.stream_ptr(x0.device))
            if rc == _lib.ERR_UNSUPPORTED:
                # parameters outside the column-table kernel: pool into the reference layout, let torch restride
                tmp = torch.empty((K, C, ph, pw), dtype=x0.dtype, device=x0.device)
                _lib.check(L.cpm_roi_align_forward_ex(*args, _lib.POOLED_KCHW, _lib.ptr(tmp), _lib.stream_ptr(x0.device)))
                return tmp.contiguous(memory_format=torch.channels_last)
            _lib.check(rc)
        else:
            _lib.check(L.cpm_roi_align_forward_ex(*args, _lib.POOLED_KCHW, _lib.ptr(out), _lib.stream_ptr(x0.device)))
    return out


_WARNED = set()


def _warn_once(key, msg):
    if key not in _WARNED:
        _WARNED.add(key)
        import warnings
        warnings.warn(msg, RuntimeWarning, stacklevel=3)


def pooler_backward(grad_output, shapes, scales, rois, output_size, sampling_ratio, aligned, interpolation, mapper,
                    mode=None, nchw_grad=False):
    """Dense gradients of every level: list[(B,C,H_l,W_l)].

    nchw_grad=True (the forward was fed NCHW-contiguous maps, what the reference's FPN emits) asks for NCHW-contiguous
    gradients, the layout _C.roi_align_backward returns (ROIAlign_cuda.cu:451-452): the deterministic tile kernels write
    them directly.  Otherwise the fast kernels write NHWC and the tensors come back with channels_last strides."""
    _lib.require_cuda(grad_output, "grad_output")
    mode = BACKWARD_MODE if mode is None else mode
    if mode not in ("deterministic", "atomic"):
        raise ValueError("cpm_ops: CPM_ROI_ALIGN_BACKWARD / mode must be 'deterministic' or 'atomic' (got %r)" % (mode,))
    if grad_output.dtype == torch.bfloat16:
        # bf16 pooled gradient (native bf16 forward): accumulate the dense gradient in fp32, round once at the end
        grads = pooler_backward(grad_output.float(), shapes, scales, rois, output_size, sampling_ratio, aligned,
                                interpolation, mapper, mode, nchw_grad)
        return [g.to(torch.bfloat16) for g in grads]
    ph, pw = output_size
    K = rois.shape[0]
    dt, dev = grad_output.dtype, grad_output.device
    B, C = shapes[0][0], shapes[0][1]
    nhwc = dt == torch.float32 and interpolation == 0 and C % 4 == 0
    staged = nhwc and sampling_ratio >= 1 and ph * sampling_ratio <= 32 and pw * sampling_ratio <= 32
    if mode == "deterministic" and not staged:
        _warn_once(("det", ph, pw, int(sampling_ratio), str(dt), interpolation),
                   "cpm_ops: the deterministic RoIAlign backward needs fp32, bilinear interpolation, C % 4 == 0 and "
                   "1 <= sampling_ratio with pooled size * sampling_ratio <= 32; this call uses the atomic "
                   "(order-non-deterministic) scatter instead")
    use = _lib.BWD_DETERMINISTIC if (mode == "deterministic" and staged) else _lib.BWD_ATOMIC
    # a channels_last-strided pooled gradient (what a channels_last conv head hands back) is read in place by the staged
    # kernel; anything else is taken in the reference's (K,C,PH,PW) order
    khwc = (use == _lib.BWD_DETERMINISTIC and K > 0 and ph * pw > 1 and C > 1 and not grad_output.is_contiguous()
            and grad_output.is_contiguous(memory_format=torch.channels_last) and grad_output.data_ptr() % 16 == 0)
    if not khwc:
        grad_output = grad_output.contiguous()
    out_nchw = (not nhwc) or (nchw_grad and use == _lib.BWD_DETERMINISTIC)
    fmt = torch.contiguous_format if out_nchw else torch.channels_last
    grads = [torch.empty(tuple(s), dtype=dt, device=dev, memory_format=fmt) for s in shapes]
    if B == 0 or any(g.numel() == 0 for g in grads):
        return grads
    pyr = make_pyramid(grads, scales, _lib.NCHW if out_nchw else _lib.NHWC)
    rois = rois.contiguous()
    layout = _lib.POOLED_KHWC if khwc else _lib.POOLED_KCHW
    ws, ws_bytes = None, 0
    with _lib.device_of(grad_output):
        if use == _lib.BWD_DETERMINISTIC:
            ws_bytes = int(_lib.lib().cpm_roi_align_backward_workspace_bytes_pyr(ctypes.byref(pyr), K, ph, pw,
                                                                                 int(sampling_ratio)))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(_lib.lib().cpm_roi_align_backward_ex(ctypes.byref(pyr), _lib.ptr(grad_output), _lib.ptr(rois), K, ph, pw,
                                                        int(sampling_ratio), int(bool(aligned)), interpolation,
                                                        ctypes.byref(mapper) if mapper is not None else None, None, use,
                                                        layout, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)))
    return grads


class _ROIAlign(Function):
    """pet/lib/ops/roi_align.py:14-60: the reference's argument list and its `(grad_input, None x 6)` backward, plus one
    optional trailing argument (extension): channels_last=True returns the pooled tensor with channels_last strides."""

    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, sampling_ratio, aligned, interpolation="bilinear",
                channels_last=False):
        ctx.save_for_backward(roi)
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.sampling_ratio = sampling_ratio
        ctx.input_shape = input.size()
        ctx.aligned = aligned
        ctx.interpolation_method = INTERPOLATION_METHOD[interpolation]
        ctx.nchw_input = not _is_nhwc(input)
        return pooler_forward([input], [spatial_scale], roi, ctx.output_size, sampling_ratio, aligned,
                              ctx.interpolation_method, None, channels_last=bool(channels_last))

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, = ctx.saved_tensors
        grad_input = pooler_backward(grad_output, [ctx.input_shape], [ctx.spatial_scale], rois, ctx.output_size,
                                     ctx.sampling_ratio, ctx.aligned, ctx.interpolation_method, None,
                                     nchw_grad=ctx.nchw_input)[0]
        return grad_input, None, None, None, None, None, None, None


roi_align = _ROIAlign.apply


def _float_function(x, keep_bf16=False):
    """apex.amp.float_function (roi_align.py:76): half inputs are computed in fp32.  keep_bf16 (ROIAlign.native_bf16 /
    Pooler.native_bf16, an extension) hands bf16 feature maps to the bf16x8 kernel instead of casting them."""
    if keep_bf16 and x.dtype == torch.bfloat16:
        return x
    return x.float() if x.dtype in (torch.float16, torch.bfloat16) else x


class ROIAlign(nn.Module):
    """pet/lib/ops/roi_align.py:66-95."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned, interpolation="bilinear"):
        assert interpolation in INTERPOLATION_METHOD, "Unknown interpolation method: {}".format(interpolation)
        super(ROIAlign, self).__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned
        self.interpolation_method = interpolation
        # Extension (not in the reference): torch.channels_last makes forward return the pooled tensor with channels_last
        # strides -- for heads that run their convolutions in channels_last.  The default is the reference's layout
        # (heads that flatten with x.view(K, -1), cls_heads.py:43, need it).
        self.pooled_memory_format = torch.contiguous_format
        # Extension: consume bf16 feature maps as stored (bf16x8 loads, fp32 arithmetic, bf16 output) instead of the
        # reference's cast-to-fp32 (apex.amp.float_function)
        self.native_bf16 = False

    def forward(self, input, rois):
        """input: NCHW images (any strides); rois: Bx5 boxes, first column = index into N, then xyxy."""
        assert rois.dim() == 2 and rois.size(1) == 5
        return roi_align(_float_function(input, self.native_bf16), _float_function(rois), self.output_size,
                         self.spatial_scale, self.sampling_ratio, self.aligned, self.interpolation_method,
                         self.pooled_memory_format == torch.channels_last)

    def __repr__(self):
        tmpstr = self.__class__.__name__ + "("
        tmpstr += "output_size=" + str(self.output_size)
        tmpstr += ", spatial_scale=" + str(self.spatial_scale)
        tmpstr += ", sampling_ratio=" + str(self.sampling_ratio)
        tmpstr += ", aligned=" + str(self.aligned)
        tmpstr += ")"
        return tmpstr
