"""TEST INFRASTRUCTURE ONLY -- numpy front-end of the CPU oracle (oracle/cpm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package cpm_r_cnn_b200 never does (tests/test_boundary_cpu.py greps for it).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcpm_oracle.so")
_lib = None

FLAVOR_PLAIN, FLAVOR_TV_CUDA, FLAVOR_ML_CUDA = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "cpm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libcpm_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_nms.restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def roi_align_forward(feat, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, aligned,
                      interpolation_method=0, strict_aligned_check=False):
    """ROIAlign.h:57-65 semantics on numpy arrays (float32 or float64)."""
    dt = np.float64 if feat.dtype == np.float64 else np.float32
    suf = "f64" if dt == np.float64 else "f32"
    feat, rois = _c(feat, dt), _c(rois, dt).reshape(-1, 5)
    B, C, H, W = feat.shape
    K = rois.shape[0]
    out = np.empty((K, C, pooled_h, pooled_w), dt)
    fn = getattr(lib(), "orc_roi_align_fwd_" + suf)
    sc = ctypes.c_double(spatial_scale) if dt == np.float64 else ctypes.c_float(spatial_scale)
    rc = fn(_p(feat), B, C, H, W, _p(rois), K, sc, pooled_h, pooled_w, int(sampling_ratio),
            int(bool(aligned)), int(interpolation_method), int(strict_aligned_check), _p(out))
    if rc != 0:
        raise RuntimeError("ROIs in ROIAlign cannot have non-negative size!")
    return out


def roi_align_backward(grad, rois, spatial_scale, pooled_h, pooled_w, B, C, H, W, sampling_ratio, aligned,
                       interpolation_method=0, strict_aligned_check=False):
    """ROIAlign.h:98-110 semantics on numpy arrays."""
    dt = np.float64 if grad.dtype == np.float64 else np.float32
    suf = "f64" if dt == np.float64 else "f32"
    grad, rois = _c(grad, dt), _c(rois, dt).reshape(-1, 5)
    K = rois.shape[0]
    gin = np.empty((B, C, H, W), dt)
    fn = getattr(lib(), "orc_roi_align_bwd_" + suf)
    sc = ctypes.c_double(spatial_scale) if dt == np.float64 else ctypes.c_float(spatial_scale)
    rc = fn(_p(grad), _p(rois), K, sc, pooled_h, pooled_w, B, C, H, W, int(sampling_ratio),
            int(bool(aligned)), int(interpolation_method), int(strict_aligned_check), _p(gin))
    if rc != 0:
        raise RuntimeError("ROIs in ROIAlign do not have non-negative size!")
    return gin


def touched_pixels(rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, aligned, B, H, W):
    """Count of distinct (img, y, x) pixels read by any bilinear tap (SURVEY.md 8(d)'s U, one level)."""
    rois = _c(rois, np.float32).reshape(-1, 5)
    mask = np.zeros((B, H, W), np.uint8)
    lib().orc_roi_align_touch_f32(_p(rois), rois.shape[0], ctypes.c_float(spatial_scale), pooled_h, pooled_w,
                                  int(sampling_ratio), int(bool(aligned)), B, H, W, _p(mask))
    return int(mask.sum())


def level_map(rois, k_min, k_max, canonical_scale=224.0, canonical_level=4.0, eps=1e-6, recip_div=False):
    """poolers.py:29-40 on (K,5) rois."""
    rois = _c(rois, np.float32).reshape(-1, 5)
    out = np.empty((rois.shape[0],), np.int64)
    lib().orc_level_map(_p(rois), rois.shape[0], ctypes.c_float(k_min), ctypes.c_float(k_max),
                        ctypes.c_float(canonical_scale), ctypes.c_float(canonical_level), ctypes.c_float(eps),
                        int(recip_div), _p(out))
    return out


def nms(boxes, scores, iou_threshold, labels=None, topk=0, flavor=FLAVOR_PLAIN):
    """keep indices (int64, descending score). labels=None -> torchvision.ops.nms; else ml_nms.h:16-21."""
    boxes, scores = _c(boxes, np.float32).reshape(-1, 4), _c(scores, np.float32).reshape(-1)
    N = boxes.shape[0]
    keep = np.empty((max(N, 1),), np.int64)
    lab = None
    if labels is not None:
        lab = _c(labels, np.int64).reshape(-1)
    n = lib().orc_nms(_p(boxes), _p(scores), _p(lab) if lab is not None else None, ctypes.c_int64(N),
                      ctypes.c_float(iou_threshold), ctypes.c_int64(topk), int(flavor), _p(keep))
    return keep[:n].copy()


def calc_sub_regions(grid_points, grid_size, whole_map_size):
    """pet/rcnn/modeling/grid_rcnn/loss.py:244-273 (only the (x1, y1) offsets are used by decode)."""
    half = whole_map_size // 4 * 2
    res = []
    for i in range(grid_points):
        xi, yi = i // grid_size, i % grid_size
        sx = 0 if xi == 0 else half if xi == grid_size - 1 else max(int((xi / (grid_size - 1) - 0.25) * whole_map_size), 0)
        sy = 0 if yi == 0 else half if yi == grid_size - 1 else max(int((yi / (grid_size - 1) - 0.25) * whole_map_size), 0)
        res.append((sx, sy, sx + half, sy + half))
    return res


def grid_decode(logits, boxes, sub_regions, mapping_ratio, return_aux=False):
    """grid_cascade_rcnn/inference.py:189-279 on numpy arrays; returns (R,4) boxes."""
    logits, boxes = _c(logits, np.float32), _c(boxes, np.float32).reshape(-1, 4)
    R, P, h, w = logits.shape
    sub = np.ascontiguousarray([[s[0], s[1]] for s in sub_regions], dtype=np.int32)
    out = np.empty((R, 4), np.float32)
    sc = np.empty((R, P), np.float32)
    pos = np.empty((R, P), np.int32)
    lib().orc_grid_decode(_p(logits), _p(boxes), R, P, h, w, _p(sub), ctypes.c_float(mapping_ratio), _p(out),
                          _p(sc), _p(pos))
    return (out, sc, pos) if return_aux else out
