"""Where the backward's time goes: the bench workload's RoIs restricted to one FPN level at a time (all levels are still
written), plus no RoIs at all (pure zero-fill of the pyramid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cpm_r_cnn_b200 import _lib, synthetic as sy
from cpm_r_cnn_b200.roi_align import pooler_backward, pooler_forward
dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
shapes = [tuple(f.shape) for f in feats_h]
feats = [f.to(dev).contiguous(memory_format=torch.channels_last) for f in feats_h]
mapper = _lib.make_mapper(2, 5)
lv = sy.fpn_levels_host(rois_h)
def timeit(fn, n=20):
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
for P, go_h in zip((7, 14), gouts_h):
    for sel_name, sel in [("all", torch.ones_like(lv, dtype=torch.bool)), ("none", torch.zeros_like(lv, dtype=torch.bool))] + [("P%d" % (l + 2), lv == l) for l in range(4)]:
        r = rois_h[sel].to(dev)
        go = go_h[sel].to(dev)
        tb = timeit(lambda: pooler_backward(go, shapes, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper))
        tf = timeit(lambda: pooler_forward(feats, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper)) if r.shape[0] else 0.0
        print("P=%2d rois=%-4s K=%4d  bwd %.4f ms  fwd %.4f ms" % (P, sel_name, r.shape[0], tb, tf))
# scan cost: the same RoIs moved far outside the image (same areas -> same levels and (level, image) lists, but no tile
# is reached): every CTA scans its list, finds no candidate and stores zeros
for P, go_h in zip((7, 14), gouts_h):
    r = rois_h.clone()
    r[:, 1:] += 1.0e5
    r = r.to(dev)
    go = go_h.to(dev)
    tb = timeit(lambda: pooler_backward(go, shapes, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper))
    print("P=%2d rois=outside K=%4d  bwd %.4f ms" % (P, r.shape[0], tb))
# half / double the RoIs of the workload (same distribution)
import torch as _t
for P, go_h in zip((7, 14), gouts_h):
    for frac in (0.25, 0.5):
        n = int(rois_h.shape[0] * frac)
        idx = _t.arange(0, rois_h.shape[0], int(1 / frac))
        r = rois_h[idx].to(dev); go = go_h[idx].to(dev)
        tb = timeit(lambda: pooler_backward(go, shapes, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper))
        tf = timeit(lambda: pooler_forward(feats, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper))
        print("P=%2d rois=%.2f K=%4d  bwd %.4f ms  fwd %.4f ms" % (P, frac, r.shape[0], tb, tf))
