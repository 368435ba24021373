"""One warm-up-then-measure pass of the bench step (7x7 + 14x14 Pooler fwd+bwd on the configs[1] workload) for ncu.
    python tools/profile_step.py [--steps N] [--nms]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cpm_r_cnn_b200 import _lib, synthetic as sy  # noqa: E402
from cpm_r_cnn_b200.roi_align import pooler_backward, pooler_forward  # noqa: E402
import cpm_r_cnn_b200 as ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--nms", action="store_true")
ap.add_argument("--graph", action="store_true", help="time CUDA-graph replays (no host enqueue cost in the numbers)")
ap.add_argument("--mode", default="deterministic")
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--cl", action="store_true", help="channels_last pooled tensors (forward output and grad_out)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
feats = [f.to(dev).contiguous(memory_format=torch.channels_last) for f in feats_h]
rois = rois_h.to(dev)
gouts = [g.to(dev) for g in gouts_h]
if args.cl:
    gouts = [g.contiguous(memory_format=torch.channels_last) for g in gouts]
shapes = [tuple(f.shape) for f in feats_h]
mapper = _lib.make_mapper(2, 5)
tot = {}
if args.graph:
    ops_ = []
    for p, go in zip(bench.POOLERS, gouts):
        ops_.append(lambda p=p: pooler_forward(feats, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, impl=args.impl, channels_last=args.cl))
        ops_.append(lambda p=p, go=go: pooler_backward(go, shapes, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, mode=args.mode))
    for f in ops_:
        f(); f()
    torch.cuda.synchronize()
    graphs, keep = [], []
    for f in ops_:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep.append(f())
        graphs.append(g)
    for i in range(args.steps + 3):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        for j, g in enumerate(graphs):
            evs[j].record(); g.replay()
        evs[4].record()
        torch.cuda.synchronize()
        if i >= 3:
            for j, n in enumerate(["fwd7", "bwd7", "fwd14", "bwd14"]):
                tot.setdefault(n, []).append(evs[j].elapsed_time(evs[j + 1]))
    args.steps = 0
for i in range(args.steps):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    k = 0
    for p, go in zip(bench.POOLERS, gouts):
        evs[k].record()
        out = pooler_forward(feats, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, impl=args.impl, channels_last=args.cl)
        evs[k + 1].record()
        grads = pooler_backward(go, shapes, list(sy.FPN_SCALES), rois, p, 2, False, 0, mapper, mode=args.mode)
        k += 2
    evs[k].record()
    torch.cuda.synchronize()
    if i >= 2:
        for j, n in enumerate(["fwd7", "bwd7", "fwd14", "bwd14"]):
            tot.setdefault(n, []).append(evs[j].elapsed_time(evs[j + 1]))
if tot:
    ab = bench.algorithmic_bytes(rois_h)
    for n, v in tot.items():
        ms = sorted(v)[len(v) // 2]
        print("%-6s median %.4f ms  min %.4f ms   %.0f GB/s  frac %.3f" % (n, ms, min(v), ab[n] / ms / 1e6, ab[n] / ms / 1e6 / 6540.8))
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
