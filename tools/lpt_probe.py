"""Does the launch order of the forward's CTAs matter?  The bench workload's RoIs in their own order, sorted by decreasing
/ increasing footprint (what a longest-first schedule would give), forward only."""
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
scale = torch.tensor(sy.FPN_SCALES)[lv]
fp = ((rois_h[:, 3] - rois_h[:, 1]) * scale + 2) * ((rois_h[:, 4] - rois_h[:, 2]) * scale + 2)
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
orders = {"as generated": torch.arange(rois_h.shape[0]), "largest first": torch.argsort(fp, descending=True),
          "smallest first": torch.argsort(fp), "by level, coarse first": torch.argsort(lv, descending=True, stable=True)}
for P, go_h in zip((7, 14), gouts_h):
    for name, o in orders.items():
        r = rois_h[o].to(dev); go = go_h[o].to(dev)
        tf = timeit(lambda: pooler_forward(feats, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper))
        tb = timeit(lambda: pooler_backward(go, shapes, list(sy.FPN_SCALES), r, (P, P), 2, False, 0, mapper))
        print("P=%2d  %-24s fwd %.4f ms   bwd %.4f ms" % (P, name, tf, tb))
