"""Forward self-check on the GPU box: the row-streaming TMA kernel (CPM_FWD_ROWS) against the column-table kernel
(CPM_FWD_COLS) and the general gather (CPM_FWD_NHWC) on the bench workload, on small maps with adversarial RoIs, for
sampling_ratio 1 / 2 and aligned on / off; then CUDA-graph replay times of both kernels.   python tools/fwd_check.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cpm_r_cnn_b200 import _lib, synthetic as sy  # noqa: E402
from cpm_r_cnn_b200.roi_align import pooler_forward  # noqa: E402

dev = torch.device("cuda", 0)
mapper = _lib.make_mapper(2, 5)
ok = True


def compare(tag, feats, scales, rois, P, sr, aligned, mp):
    global ok
    a = pooler_forward(feats, scales, rois, (P, P), sr, aligned, 0, mp, impl=_lib.FWD_ROWS)
    b = pooler_forward(feats, scales, rois, (P, P), sr, aligned, 0, mp, impl=_lib.FWD_NHWC)
    c = pooler_forward(feats, scales, rois, (P, P), sr, aligned, 0, mp, impl=_lib.FWD_COLS)
    torch.cuda.synchronize()
    rms = float(b.pow(2).mean().sqrt())
    exc = ((a - b).abs() - 1e-5 * (b.abs() + rms)).max().item()
    excc = ((c - b).abs() - 1e-5 * (b.abs() + rms)).max().item()
    bad = (((a - b).abs() - 1e-5 * (b.abs() + rms)) > 0).flatten(1).any(1).nonzero().flatten().tolist()[:8]
    print("%-28s P=%2d sr=%d aligned=%d: rows max err %.3e (excess %.3e), cols excess %.3e, finite %s %s" % (
        tag, P, sr, aligned, (a - b).abs().max().item(), exc, excc, bool(torch.isfinite(a).all()), bad if bad else ""))
    ok = ok and exc <= 0 and bool(torch.isfinite(a).all())


rois_h, feats_h, gouts_h = bench.make_workload(0)
feats = [f.to(dev).contiguous(memory_format=torch.channels_last) for f in feats_h]
rois = rois_h.to(dev)
scales = list(sy.FPN_SCALES)
for P in (7, 14):
    for sr in (2, 1):
        for aligned in (False, True):
            compare("bench workload", feats, scales, rois, P, sr, aligned, mapper)
gen = torch.Generator().manual_seed(5)
small = [f.to(dev).contiguous(memory_format=torch.channels_last) for f in sy.pyramid(gen, 2, 128, 200, 336)]
r2 = torch.cat([sy.coco_like_rois(gen, 64, 2, 200, 336), sy.adversarial_rois(200, 336, 2)], 0).to(dev)
for P in (7, 14):
    for aligned in (False, True):
        compare("small maps + adversarial", small, scales, r2, P, 2, aligned, mapper)
# one level, a map wider than 48 pixels pooled whole (direct-gather path), a tall map (> 128 rows)
big = [torch.randn(1, 128, 300, 90, generator=gen).to(dev).contiguous(memory_format=torch.channels_last)]
r3 = torch.tensor([[0, 0, 0, 89, 299], [0, 3, 5, 60, 280], [0, 10, 10, 40, 250], [0, 20.5, 30.25, 33, 47], [0, 0, 0, 47, 20]],
                  dtype=torch.float32, device=dev)
for P in (7, 14):
    compare("single level, large RoIs", big, [1.0], r3, P, 2, False, None)

if ok and "--time" in sys.argv:
    for P in (7, 14):
        for name, impl in (("cols", _lib.FWD_COLS), ("rows", _lib.FWD_ROWS)):
            fn = lambda: pooler_forward(feats, scales, rois, (P, P), 2, False, 0, mapper, impl=impl)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            g, _ = bench.capture(fn)
            print("P=%2d %s: %.4f ms" % (P, name, bench.time_graphs([g], 30)[0]))
print("FWD_CHECK", "OK" if ok else "FAILED")
