"""Backward self-check on the GPU box: deterministic (TMA / staged) vs the red.global.add scatter on the bench workload,
NHWC and NCHW gradient pyramids, run-to-run bit identity.   python tools/bwd_check.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cpm_r_cnn_b200 import _lib, synthetic as sy  # noqa: E402
from cpm_r_cnn_b200.roi_align import pooler_backward  # noqa: E402

dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
shapes = [tuple(f.shape) for f in feats_h]
rois = rois_h.to(dev)
mapper = _lib.make_mapper(2, 5)
ok = True
for p, go_h in zip(bench.POOLERS, gouts_h):
    go = go_h.to(dev)
    at = pooler_backward(go, shapes, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, mode="atomic")
    ab = pooler_backward(go.abs(), shapes, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, mode="atomic")
    for nchw in (False, True):
        d1 = pooler_backward(go, shapes, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, mode="deterministic", nchw_grad=nchw)
        d2 = pooler_backward(go, shapes, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, mode="deterministic", nchw_grad=nchw)
        torch.cuda.synchronize()
        for l, (a, b, r, s) in enumerate(zip(d1, d2, at, ab)):
            same = torch.equal(a, b)
            rms = float(r.pow(2).mean().sqrt())
            exc = ((a - r).abs() - 1e-5 * (s + rms)).max().item()
            fmt = "nchw" if a.is_contiguous() else "nhwc"
            print("P=%d %s level %d: run-to-run identical %s, max err %.3e, excess over bound %.3e" % (
                p[0], fmt, l, same, (a - r).abs().max().item(), exc))
            ok = ok and same and exc <= 0
print("BWD_CHECK", "OK" if ok else "FAILED")
