"""Randomised parity sweep of the row-streaming forward against the reference-shaped generic kernel: map sizes from 1x1 to
~90x130 (row segments beyond 48 pixels, footprints beyond 128 rows included), 1-4 levels, every supported channel count, RoIs
from sub-pixel to several times the map, inside and outside, both alignments and sampling grids.   python tools/fwd_fuzz.py [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpm_r_cnn_b200 import _lib  # noqa: E402
from cpm_r_cnn_b200.roi_align import pooler_forward  # noqa: E402

dev = torch.device("cuda", 0)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
gen = torch.Generator().manual_seed(2024)
worst, bad = -1e9, 0
for case in range(n_cases):
    P = 7 if torch.rand(1, generator=gen).item() < 0.5 else 14
    C = (128 if P == 7 else 64) * int(torch.randint(1, 4 if P == 7 else 5, (1,), generator=gen))
    L = int(torch.randint(1, 5, (1,), generator=gen))
    B = int(torch.randint(1, 4, (1,), generator=gen))
    sr = int(torch.randint(1, 3, (1,), generator=gen))
    aligned = bool(torch.randint(0, 2, (1,), generator=gen))
    H0 = int(torch.randint(1, 180 if case % 7 == 0 else 90, (1,), generator=gen))
    W0 = int(torch.randint(1, 130, (1,), generator=gen))
    scales = [1.0 / (1 << l) for l in range(L)]
    feats = [torch.randn(B, C, max(1, -(-H0 >> l)), max(1, -(-W0 >> l)), generator=gen).to(dev).contiguous(memory_format=torch.channels_last)
             for l in range(L)]
    K = int(torch.randint(1, 40, (1,), generator=gen))
    side = torch.exp(torch.empty(K).uniform_(-1.5, 5.5, generator=gen))
    ar = torch.exp(torch.empty(K).uniform_(-1.2, 1.2, generator=gen))
    w, h = side * ar.sqrt(), side / ar.sqrt()
    cx = (torch.rand(K, generator=gen) * 1.6 - 0.3) * W0
    cy = (torch.rand(K, generator=gen) * 1.6 - 0.3) * H0
    rois = torch.stack([torch.randint(0, B, (K,), generator=gen).float(), cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    if case % 5 == 0:
        rois[0, 0] = B + 2                      # bad image index: zeros
    if aligned and case % 3 == 0:
        rois[-1, 3], rois[-1, 1] = rois[-1, 1].clone(), rois[-1, 3].clone()      # negative width
        rois[-1, 4], rois[-1, 2] = rois[-1, 2].clone(), rois[-1, 4].clone()      # negative height
    rois = rois.to(dev)
    mapper = _lib.make_mapper(0, L - 1, 32.0, 1.0) if L > 1 else None
    a = pooler_forward(feats, scales, rois, (P, P), sr, aligned, 0, mapper, impl=_lib.FWD_ROWS)
    b = pooler_forward(feats, scales, rois, (P, P), sr, aligned, 0, mapper, impl=_lib.FWD_NHWC)
    torch.cuda.synchronize()
    rms = float(b.pow(2).mean().sqrt())
    exc = float(((a - b).abs() - 1e-5 * (b.abs() + rms)).max())
    fin = bool(torch.isfinite(a).all())
    worst = max(worst, exc)
    if exc > 0 or not fin:
        bad += 1
        print("case %d FAILED: P=%d C=%d L=%d B=%d sr=%d aligned=%d map %dx%d K=%d excess %.3e finite %s" % (
            case, P, C, L, B, sr, aligned, H0, W0, K, exc, fin))
print("FWD_FUZZ %s: %d cases, worst excess over the 1e-5 bound %.3e" % ("OK" if bad == 0 else "FAILED", n_cases, worst))
