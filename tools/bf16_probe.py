"""Native bf16 forward (bf16 pyramid, bf16x8 taps, fp32 arithmetic, bf16 pooled output) beside the fp32 forward."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cpm_r_cnn_b200 import _lib, synthetic as sy
from cpm_r_cnn_b200.roi_align import pooler_forward
dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
feats = [f.to(dev).contiguous(memory_format=torch.channels_last) for f in feats_h]
feats16 = [f.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in feats]
mapper = _lib.make_mapper(2, 5)
r = rois_h.to(dev)
def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): keep = fn()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    for i in range(n):
        e[i].record(); g.replay()
    e[n].record(); torch.cuda.synchronize()
    t = sorted(e[i].elapsed_time(e[i + 1]) for i in range(n))
    return t[n // 2]
for P in (7, 14):
    t32 = timeit(lambda: pooler_forward(feats, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper))
    t16 = timeit(lambda: pooler_forward(feats16, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper))
    o = pooler_forward(feats16, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper)
    print("P=%2d fwd fp32 %.4f ms   bf16 %.4f ms  (out %s %s)" % (P, t32, t16, o.dtype, tuple(o.shape)))
