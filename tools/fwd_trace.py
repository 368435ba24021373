"""Per-CTA timeline of the row-streaming forward kernel (globaltimer stamps written by a build with -DFWDR_TRACE=1):
    touch cpm_r_cnn_b200/csrc/roi_align_fwd_rows.cu; CPM_NVCC_EXTRA=-DFWDR_TRACE=1 python -m cpm_r_cnn_b200.build
    python tools/fwd_trace.py          # on the GPU box; rebuild without the flag afterwards
Stamps: 0 kernel entry, 1 geometry done (last warp), 2 first row arrived, 3 last row of chunk 0 pooled, 5 exit, 6 producer done."""
import ctypes, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from cpm_r_cnn_b200 import _lib, synthetic as sy
from cpm_r_cnn_b200.roi_align import pooler_forward
dev = torch.device("cuda", 0)
rois_h, feats_h, _ = bench.make_workload(0)
feats = [f.to(dev).contiguous(memory_format=torch.channels_last) for f in feats_h]
rois = rois_h.to(dev); mapper = _lib.make_mapper(2, 5)
L = ctypes.CDLL(os.path.join(ROOT, "cpm_r_cnn_b200", "libcpm_ops.so"))
for P in (7, 14):
    for _ in range(3):
        pooler_forward(feats, list(sy.FPN_SCALES), rois, (P, P), 2, False, 0, mapper, impl=_lib.FWD_ROWS)
    torch.cuda.synchronize()
    n = (4096 if P == 14 else 2048) // int(os.environ.get("CPM_FWD_CPC", "1" if P == 7 else "2"))
    buf = np.zeros(8192 * 8, dtype=np.uint64)
    L.cpm_debug_fwd_trace(buf.ctypes.data_as(ctypes.c_void_p), 8192 * 8)
    t = buf.reshape(8192, 8)[:n].astype(np.int64)
    d = lambda a, b: (t[:, b] - t[:, a]) / 1e3
    nr = t[:, 7] >> 32
    span = (t[:, 5].max() - t[:, 0].min()) / 1e3
    print("   producer done at %.2f after start (consumers' last row of chunk 0 at %.2f, end %.2f)" % (np.median(d(0, 6)), np.median(d(0, 3)), np.median(d(0, 5))))
    print("P=%d: span %.1f us | start->tables %.2f | ->first row %.2f | chunk-0 rows %.2f (%.3f/row) | ->end %.2f | total %.2f (medians, us); concurrency %.0f" % (
        P, span, np.median(d(0, 1)), np.median(d(1, 2)), np.median(d(2, 3)), np.median(d(2, 3) / nr), np.median(d(3, 5)), np.median(d(0, 5)), d(0, 5).sum() / span))
