"""Per-CTA timeline of the staged backward tile kernel (globaltimer stamps written by a build with -DBWD_TRACE=1):
    touch cpm_r_cnn_b200/csrc/roi_align_bwd.cu; CPM_NVCC_EXTRA=-DBWD_TRACE=1 python -m cpm_r_cnn_b200.build
    python tools/bwd_trace.py          # on the GPU box; rebuild without the flag afterwards
Stamps: 0 kernel entry, 1 candidate list ready, 2 last item accumulated, 3 tile written; 4 = items, 5 = candidates."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from cpm_r_cnn_b200 import _lib, synthetic as sy
from cpm_r_cnn_b200.roi_align import pooler_backward
dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
shapes = [tuple(f.shape) for f in feats_h]
rois = rois_h.to(dev); mapper = _lib.make_mapper(2, 5)
L = ctypes.CDLL(os.path.join(ROOT, "cpm_r_cnn_b200", "libcpm_ops.so"))
for p, go_h in zip(bench.POOLERS, gouts_h):
    go = go_h.to(dev)
    for nchw in (False, True):
        for _ in range(3):
            pooler_backward(go, shapes, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, nchw_grad=nchw)
        torch.cuda.synchronize()
        n = 5696
        buf = np.zeros(8192 * 8, dtype=np.uint64)
        L.cpm_debug_bwd_trace(buf.ctypes.data_as(ctypes.c_void_p), 8192 * 8)
        t = buf.reshape(8192, 8)[:n].astype(np.int64)
        d = lambda a, b: (t[:, b] - t[:, a]) / 1e3
        span = (t[:, 3].max() - t[:, 0].min()) / 1e3
        items, pc = t[:, 4], t[:, 5]
        busy = items > 0
        print("P=%d %s: span %.1f us, %d CTAs (%d with work), concurrency %.0f" % (p[0], "nchw" if nchw else "nhwc", span, n, busy.sum(), d(0, 3).sum() / span))
        print("   with work (mean us): list %.2f | items %.2f (%.1f items, %.2f us/item, %.1f candidates) | write-out %.2f | total %.2f; empty tiles: total %.2f" % (
            d(0, 1)[busy].mean(), d(1, 2)[busy].mean(), items[busy].mean(), (d(1, 2)[busy] / items[busy]).mean(), pc[busy].mean(),
            d(2, 3)[busy].mean(), d(0, 3)[busy].mean(), d(0, 3)[~busy].mean() if (~busy).any() else 0.0))
        tot = d(0, 3).sum()
        print("   share of CTA-time: list %.0f%%, items %.0f%%, write-out %.0f%%, empty tiles %.0f%%" % (
            100 * d(0, 1)[busy].sum() / tot, 100 * d(1, 2)[busy].sum() / tot, 100 * d(2, 3)[busy].sum() / tot, 100 * d(0, 3)[~busy].sum() / tot))
