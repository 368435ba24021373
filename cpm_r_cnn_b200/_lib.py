"""ctypes binding of libcpm_ops.so (include/cpm_ops.h).  There is no fallback: if the CUDA library is missing or a
tensor is not on a CUDA device the ops raise RuntimeError (the reference raises the same way for CPU tensors on its
CUDA-only ops, e.g. ml_nms.h:38 "CPU version not implemented")."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPM_OPS_LIB") or os.path.join(_HERE, "libcpm_ops.so")

CPM_MAX_LEVELS = 8
F32, F64, BF16 = 0, 1, 2
NCHW, NHWC = 0, 1
INTERP = {"bilinear": 0, "nearest": 1}
IOU_PLAIN, IOU_TV_CUDA, IOU_ML_CUDA = 0, 1, 2
BWD_DETERMINISTIC, BWD_ATOMIC = 0, 1
FWD_AUTO, FWD_GENERIC, FWD_NHWC, FWD_COLS, FWD_ROWS = 0, 1, 2, 4, 8
POOLED_KCHW, POOLED_KHWC = 0, 1
ERR_UNSUPPORTED = -2


class Pyramid(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("batch", ctypes.c_int32), ("channels", ctypes.c_int32),
                ("dtype", ctypes.c_int32), ("layout", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("d_level", ctypes.c_void_p * CPM_MAX_LEVELS), ("height", ctypes.c_int32 * CPM_MAX_LEVELS),
                ("width", ctypes.c_int32 * CPM_MAX_LEVELS), ("spatial_scale", ctypes.c_float * CPM_MAX_LEVELS)]


class RpnLevels(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("num_images", ctypes.c_int32), ("row", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("d_objectness", ctypes.c_void_p * CPM_MAX_LEVELS),
                ("d_regression", ctypes.c_void_p * CPM_MAX_LEVELS), ("d_anchors", ctypes.c_void_p * CPM_MAX_LEVELS),
                ("anchors_per_image", ctypes.c_int32 * CPM_MAX_LEVELS), ("A", ctypes.c_int32 * CPM_MAX_LEVELS),
                ("HW", ctypes.c_int32 * CPM_MAX_LEVELS), ("k", ctypes.c_int32 * CPM_MAX_LEVELS)]


class LevelMapperC(ctypes.Structure):
    _fields_ = [("k_min", ctypes.c_float), ("k_max", ctypes.c_float), ("canonical_scale", ctypes.c_float),
                ("canonical_level", ctypes.c_float), ("eps", ctypes.c_float)]


_SIGNATURES = {
    "cpm_last_error": (ctypes.c_char_p, []),
    "cpm_version": (ctypes.c_int, []),
    "cpm_launch_count": (ctypes.c_uint64, []),
    "cpm_set_device": (ctypes.c_int, [ctypes.c_int]),
    "cpm_roi_align_forward": (ctypes.c_int, [ctypes.POINTER(Pyramid), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(LevelMapperC), ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "cpm_roi_align_forward_ex": (ctypes.c_int, [ctypes.POINTER(Pyramid), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.POINTER(LevelMapperC), ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_void_p]),
    "cpm_roi_align_backward_ex": (ctypes.c_int, [ctypes.POINTER(Pyramid), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.POINTER(LevelMapperC), ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "cpm_roi_align_backward_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "cpm_roi_align_backward_workspace_bytes_pyr": (ctypes.c_size_t, [ctypes.POINTER(Pyramid), ctypes.c_int64, ctypes.c_int,
                                                                     ctypes.c_int, ctypes.c_int]),
    "cpm_roi_align_backward": (ctypes.c_int, [ctypes.POINTER(Pyramid), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(LevelMapperC), ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "cpm_level_map": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(LevelMapperC), ctypes.c_void_p,
                                     ctypes.c_void_p]),
    "cpm_layout_convert": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "cpm_layout_convert_pyramid": (ctypes.c_int, [ctypes.POINTER(Pyramid), ctypes.POINTER(Pyramid), ctypes.c_void_p]),
    "cpm_nms_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64]),
    "cpm_nms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
                               ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_size_t, ctypes.c_void_p]),
    "cpm_nms_batched_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int64]),
    "cpm_nms_batched": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_float, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                       ctypes.c_void_p]),
    "cpm_grid_decode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_float, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    "cpm_rpn_decode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.c_float, ctypes.c_float,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "cpm_rpn_flatten_objectness": (ctypes.c_int, [ctypes.POINTER(RpnLevels), ctypes.c_void_p, ctypes.c_void_p]),
    "cpm_rpn_select_decode": (ctypes.c_int, [ctypes.POINTER(RpnLevels), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.c_float,
                                             ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "cpm_grid_targets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_int32), ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "cpm_box_iou": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                   ctypes.c_void_p]),
    "cpm_matcher_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int64]),
    "cpm_matcher": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_float, ctypes.c_float,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises RuntimeError when libcpm_ops.so has not been built (python -m cpm_r_cnn_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("cpm_ops: %s is missing -- build it with `python -m cpm_r_cnn_b200.build` "
                               "(there is no CPU/PyTorch fallback)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc):
    if rc != 0:
        raise RuntimeError("cpm_ops: " + lib().cpm_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(lib().cpm_launch_count())


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        # ROIAlign_cuda.cu:376 "input must be a CUDA tensor"; ml_nms.h:38 "CPU version not implemented"
        raise RuntimeError("cpm_ops: %s must be a CUDA tensor (no CPU implementation)" % name)


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


class device_of(object):
    """CUDAGuard of ROIAlign_cuda.cu:383 / ml_nms.cu:90: makes the tensor's device current for torch and the library."""

    def __init__(self, t):
        self.guard = torch.cuda.device(t.device)
        self.index = t.device.index if t.device.index is not None else torch.cuda.current_device()

    def __enter__(self):
        self.guard.__enter__()
        check(lib().cpm_set_device(self.index))
        return self

    def __exit__(self, *a):
        return self.guard.__exit__(*a)


DTYPES = {torch.float32: F32, torch.float64: F64, torch.bfloat16: BF16}


def make_mapper(k_min, k_max, canonical_scale=224.0, canonical_level=4.0, eps=1e-6):
    return LevelMapperC(float(k_min), float(k_max), float(canonical_scale), float(canonical_level), float(eps))
