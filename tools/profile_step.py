"""The bench step for profilers and quick timing.

    python tools/profile_step.py [--steps N] [--graph] [--resident] [--nms]

Default: the headline step of bench.py -- NCHW maps in, one NHWC staging of the pyramid, 7x7 and 14x14 Pooler forward +
deterministic backward writing NCHW gradients -- run eagerly `--steps` times (what ncu attaches to).  --resident: the four
ops on a pyramid that is already channels_last (NHWC gradients), the round-1 configuration.  --graph: per-op CUDA-graph
replay times (median) instead of eager runs.
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cpm_r_cnn_b200 import _lib, synthetic as sy  # noqa: E402
from cpm_r_cnn_b200.roi_align import pooler_backward, pooler_forward  # noqa: E402
import cpm_r_cnn_b200 as ops  # noqa: E402

ra = importlib.import_module("cpm_r_cnn_b200.roi_align")
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--nms", action="store_true")
ap.add_argument("--graph", action="store_true", help="time CUDA-graph replays (no host enqueue cost in the numbers)")
ap.add_argument("--resident", action="store_true", help="channels_last pyramid, NHWC gradients (round-1 configuration)")
ap.add_argument("--mode", default="deterministic")
args = ap.parse_args()
dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
feats = [f.to(dev) for f in feats_h]
if args.resident:
    feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
rois = rois_h.to(dev)
gouts = [g.to(dev) for g in gouts_h]
shapes = [tuple(f.shape) for f in feats_h]
scales = list(sy.FPN_SCALES)
mapper = _lib.make_mapper(2, 5)
nchw = not args.resident
names, fns = [], []
if nchw:
    names.append("stage")
    fns.append(lambda: ra.stage_pyramid_nhwc(feats, cache=True))
for p, go in zip(bench.POOLERS, gouts):
    names += ["fwd%d" % p[0], "bwd%d" % p[0]]
    fns.append(lambda p=p: pooler_forward(feats, scales, rois, p, 2, False, 0, mapper))
    fns.append(lambda p=p, go=go: pooler_backward(go, shapes, scales, rois, p, 2, False, 0, mapper, mode=args.mode, nchw_grad=nchw))


def run_step():
    ra.STAGING_CACHE.clear()
    return [f() for f in fns]


tot = {}
if args.graph:
    for _ in range(2):
        run_step()
    torch.cuda.synchronize()
    ra.STAGING_CACHE.clear()
    caps = [bench.capture(f) for f in fns]          # the staging graph fills the cache the forward graphs then hit
    for n, t in zip(names, bench.time_graphs([c[0] for c in caps], max(args.steps, 5))):
        tot[n] = [t]
else:
    for i in range(args.steps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(fns) + 1)]
        ra.STAGING_CACHE.clear()
        for j, f in enumerate(fns):
            evs[j].record()
            f()
        evs[len(fns)].record()
        torch.cuda.synchronize()
        if i >= 2:
            for j, n in enumerate(names):
                tot.setdefault(n, []).append(evs[j].elapsed_time(evs[j + 1]))
if tot:
    ab = bench.algorithmic_bytes(rois_h)
    total = 0.0
    for n, v in tot.items():
        ms = sorted(v)[len(v) // 2]
        total += ms
        if n in ab:
            print("%-6s median %.4f ms  min %.4f ms   %.0f GB/s  frac %.3f" % (n, ms, min(v), ab[n] / ms / 1e6, ab[n] / ms / 1e6 / 6540.8))
        else:
            print("%-6s median %.4f ms  min %.4f ms" % (n, ms, min(v)))
    print("step   %.4f ms" % total)
if args.nms:
    gen = torch.Generator().manual_seed(1000)
    b, s, seg = sy.rpn_like_candidates(gen, 16, 5, 1000)
    for i in range(2):
        ops.batched_nms(b.to(dev), s.to(dev), seg.to(dev), 80, 0.7, sync=False)
    b, s, seg, lab, img = sy.detection_candidates(gen, 16, 1000, 80, -1.0)
    for i in range(2):
        ops.batched_nms(b.to(dev), s.to(dev), seg.to(dev), 1280, 0.3, sync=False)
    logits = torch.randn(16000, 9, 28, 28, generator=gen).to(dev)     # the bench's decode size (streaming kernel)
    boxes = sy.coco_like_boxes(gen, 16000).to(dev)
    sub = ops.calc_sub_regions(9, 3, 56)
    for i in range(2):
        ops.grid_decode(logits, boxes, sub, 0.5)
    torch.cuda.synchronize()
print("ok")
