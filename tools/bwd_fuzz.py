"""Randomised parity sweep of the deterministic backward against the red.global.add scatter (the reference's own scheme):
map sizes from 1x1 to ~90x130, 1-4 levels, channel counts 4..256, RoIs from sub-pixel to several times the map, both
alignments and sampling grids, both pooled-gradient layouts, NHWC and NCHW gradient pyramids; every result also compared
with a second run (bit-identical).   python tools/bwd_fuzz.py [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpm_r_cnn_b200 import _lib  # noqa: E402
from cpm_r_cnn_b200.roi_align import pooler_backward  # noqa: E402

dev = torch.device("cuda", 0)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
gen = torch.Generator().manual_seed(77)
bad, worst = 0, -1e9
for case in range(n_cases):
    P = (7, 14, 5, 3)[int(torch.randint(0, 4, (1,), generator=gen))]
    PW = P if case % 4 else max(1, P - 2)
    C = int(torch.randint(1, 65, (1,), generator=gen)) * 4
    L = int(torch.randint(1, 5, (1,), generator=gen))
    B = int(torch.randint(1, 4, (1,), generator=gen))
    sr = int(torch.randint(1, 3, (1,), generator=gen))
    aligned = bool(torch.randint(0, 2, (1,), generator=gen))
    H0, W0 = int(torch.randint(1, 90, (1,), generator=gen)), int(torch.randint(1, 130, (1,), generator=gen))
    scales = [1.0 / (1 << l) for l in range(L)]
    shapes = [(B, C, max(1, -(-H0 >> l)), max(1, -(-W0 >> l))) for l in range(L)]
    K = int(torch.randint(1, 60, (1,), generator=gen))
    side = torch.exp(torch.empty(K).uniform_(-1.5, 5.5, generator=gen))
    ar = torch.exp(torch.empty(K).uniform_(-1.2, 1.2, generator=gen))
    w, h = side * ar.sqrt(), side / ar.sqrt()
    cx = (torch.rand(K, generator=gen) * 1.6 - 0.3) * W0
    cy = (torch.rand(K, generator=gen) * 1.6 - 0.3) * H0
    rois = torch.stack([torch.randint(0, B, (K,), generator=gen).float(), cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    if not aligned:
        rois[K // 2:] = rois[:K - K // 2].clone()          # duplicates: many RoIs on the same pixels
    rois = rois.to(dev)
    mapper = _lib.make_mapper(0, L - 1, 32.0, 1.0) if L > 1 else None
    go = torch.randn(K, C, P, PW, generator=gen).to(dev)
    if case % 3 == 0:
        go = go.contiguous(memory_format=torch.channels_last)
    nchw = bool(case % 2)
    ref = pooler_backward(go, shapes, scales, rois, (P, PW), sr, aligned, 0, mapper, mode="atomic")
    mag = pooler_backward(go.abs(), shapes, scales, rois, (P, PW), sr, aligned, 0, mapper, mode="atomic")
    a = pooler_backward(go, shapes, scales, rois, (P, PW), sr, aligned, 0, mapper, mode="deterministic", nchw_grad=nchw)
    b = pooler_backward(go, shapes, scales, rois, (P, PW), sr, aligned, 0, mapper, mode="deterministic", nchw_grad=nchw)
    torch.cuda.synchronize()
    for l in range(L):
        rms = float(ref[l].pow(2).mean().sqrt())
        exc = float(((a[l] - ref[l]).abs() - 1e-5 * (mag[l] + rms)).max())
        same = torch.equal(a[l], b[l])
        worst = max(worst, exc)
        if exc > 0 or not same or not bool(torch.isfinite(a[l]).all()):
            bad += 1
            print("case %d level %d FAILED: P=%dx%d C=%d L=%d B=%d sr=%d aligned=%d map %dx%d K=%d nchw=%d excess %.3e identical %s" % (
                case, l, P, PW, C, L, B, sr, aligned, H0, W0, K, nchw, exc, same))
print("BWD_FUZZ %s: %d cases, worst excess over the 1e-5 bound %.3e" % ("OK" if bad == 0 else "FAILED", n_cases, worst))
