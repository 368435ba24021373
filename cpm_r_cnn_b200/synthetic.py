"""Seeded synthetic inputs of SURVEY.md section 8(d): COCO-shaped RoIs, FPN pyramids, NMS candidate sets, grid
heat-maps.  Used by tests/ and bench.py (CPU generators, deterministic for a given seed)."""
import math

import torch

IMG_H, IMG_W = 800, 1344            # 800x1333 padded to SIZE_DIVISIBILITY=32 (config.py:257,294)
FPN_SCALES = (1 / 4., 1 / 8., 1 / 16., 1 / 32.)
FPN_CHANNELS = 256


def level_shapes(img_h=IMG_H, img_w=IMG_W, scales=FPN_SCALES):
    return [(int(math.ceil(img_h * s)), int(math.ceil(img_w * s))) for s in scales]


def coco_like_boxes(gen, n, img_h=IMG_H, img_w=IMG_W, s_min=16.0, s_max=600.0):
    """sqrt(area) log-uniform in [s_min, s_max], aspect ratio log-uniform in [0.5, 2], box inside the image."""
    s = torch.exp(torch.empty(n).uniform_(math.log(s_min), math.log(s_max), generator=gen))
    ar = torch.exp(torch.empty(n).uniform_(math.log(0.5), math.log(2.0), generator=gen))
    w = (s * torch.sqrt(ar)).clamp(max=img_w - 1.0)
    h = (s / torch.sqrt(ar)).clamp(max=img_h - 1.0)
    x1 = torch.rand(n, generator=gen) * (img_w - 1 - w)
    y1 = torch.rand(n, generator=gen) * (img_h - 1 - h)
    return torch.stack([x1, y1, x1 + w, y1 + h], 1)


def coco_like_rois(gen, n_per_image, n_images, img_h=IMG_H, img_w=IMG_W):
    """(K,5) [image index, x1, y1, x2, y2], images in order (what Pooler.convert_to_roi_format builds)."""
    parts = []
    for i in range(n_images):
        b = coco_like_boxes(gen, n_per_image, img_h, img_w)
        parts.append(torch.cat([torch.full((n_per_image, 1), float(i)), b], 1))
    return torch.cat(parts, 0)


def adversarial_rois(img_h, img_w, n_images):
    last = float(n_images - 1)
    return torch.tensor([
        [0, 0, 0, img_w - 1, img_h - 1],                          # whole image
        [last, -20, -30, 40, 50],                                 # sticks out top-left
        [0, img_w - 30, img_h - 20, img_w + 60, img_h + 45],      # sticks out bottom-right
        [last, 100, 100, 100, 100],                               # zero area
        [0, 50.25, 60.5, 50.75, 61.0],                            # sub-pixel
        [last, 10, 10, 12, 150],                                  # thin, tall
        [0, 5, 90, 330, 93],                                      # thin, wide
        [last, img_w + 200, img_h + 200, img_w + 300, img_h + 300],   # fully outside
    ], dtype=torch.float32)


def pyramid(gen, n_images, channels=FPN_CHANNELS, img_h=IMG_H, img_w=IMG_W, scales=FPN_SCALES, dtype=torch.float32):
    return [torch.randn(n_images, channels, h, w, generator=gen, dtype=dtype) for (h, w) in
            level_shapes(img_h, img_w, scales)]


def tie_free_scores(gen, n):
    perm = torch.randperm(n, generator=gen)
    return (perm.float() + 0.5) / n


def rpn_like_candidates(gen, n_images, n_levels, n_per_segment, img_h=IMG_H, img_w=IMG_W):
    """RPN flavour of config #3: per (image, level) n boxes = n/3 base boxes x 3 jittered copies; tie-free scores.
    Returns boxes (N,4), scores (N,), segments (N,) int32 with N = n_images*n_levels*n_per_segment."""
    boxes, scores, segs = [], [], []
    nb = (n_per_segment + 2) // 3
    for s in range(n_images * n_levels):
        base = coco_like_boxes(gen, nb, img_h, img_w)
        b = torch.cat([base + torch.randn(base.shape, generator=gen) * j for j in (0.0, 4.0, 10.0)], 0)[:n_per_segment]
        boxes.append(b)
        scores.append(tie_free_scores(gen, n_per_segment))
        segs.append(torch.full((n_per_segment,), s, dtype=torch.int32))
    return torch.cat(boxes), torch.cat(scores), torch.cat(segs)


def detection_candidates(gen, n_images, n_props, n_classes, score_thresh=0.03, img_h=IMG_H, img_w=IMG_W):
    """Detection flavour (CLSPostProcessor.filter_results, grid_cascade_rcnn/inference.py:107-124): per image
    n_props proposals x n_classes foreground classes sharing the box, scores = softmax of N(0,1) logits, gated by
    score > thresh (thresh < 0: un-gated stress variant).  segment = image * n_classes + (label - 1)."""
    boxes, scores, segs, labels, img = [], [], [], [], []
    for i in range(n_images):
        b = coco_like_boxes(gen, n_props, img_h, img_w)
        logit = torch.randn(n_props, n_classes + 1, generator=gen)
        sc = torch.softmax(logit, 1)[:, 1:]
        lab = torch.arange(1, n_classes + 1).repeat(n_props, 1)
        m = sc > score_thresh
        idx = m.nonzero()
        boxes.append(b[idx[:, 0]])
        scores.append(sc[m])
        labels.append(lab[m])
        segs.append((i * n_classes + lab[m] - 1).to(torch.int32))
        img.append(torch.full((int(m.sum()),), i, dtype=torch.int64))
    return torch.cat(boxes), torch.cat(scores), torch.cat(segs), torch.cat(labels), torch.cat(img)


def fpn_levels_host(rois, k_min=2, k_max=5, s0=224.0, lvl0=4.0, eps=1e-6):
    """LevelMapper (poolers.py:29-40) with torch CPU ops; used only to size/describe workloads."""
    area = (rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1)
    lv = torch.floor(lvl0 + torch.log2(torch.sqrt(area) / s0 + eps)).clamp(min=k_min, max=k_max)
    return lv.to(torch.int64) - k_min


def touched_pixels(rois, levels, shapes, scales, pooled, sampling_ratio, aligned=False):
    """U of SURVEY.md 8(d): number of distinct (image, level, y, x) feature pixels read by any bilinear tap of any RoI
    (tap rules of ROIAlign_cuda.cu:36-86).  rois (K,5) float32 CPU, levels (K,) int64, shapes [(H,W)] per level."""
    ph, pw = pooled
    g = int(sampling_ratio)
    assert g > 0
    total = 0
    off = 0.5 if aligned else 0.0
    for l, ((H, W), sc) in enumerate(zip(shapes, scales)):
        r = rois[levels == l]
        if r.numel() == 0:
            continue
        b = r[:, 0].to(torch.int64)
        sw, sh = r[:, 1] * sc - off, r[:, 2] * sc - off
        rw, rh = r[:, 3] * sc - off - sw, r[:, 4] * sc - off - sh
        if not aligned:
            rw, rh = rw.clamp(min=1.0), rh.clamp(min=1.0)
        bw, bh = rw / pw, rh / ph

        def taps(start, bin_, P, size):
            k = torch.arange(P * g)
            v = start[:, None] + (k // g)[None, :] * bin_[:, None] + ((k % g).float() + .5)[None, :] * bin_[:, None] / g
            valid = (v >= -1.0) & (v <= size)
            v = v.clamp(min=0.0)
            lo = v.floor().to(torch.int64).clamp(max=size - 1)
            hi = (lo + 1).clamp(max=size - 1)
            return lo, hi, valid

        ylo, yhi, yv = taps(sh, bh, ph, H)
        xlo, xhi, xv = taps(sw, bw, pw, W)
        ys = torch.stack([ylo, yhi], 2).flatten(1)          # (n, 2*PH*g)
        xs = torch.stack([xlo, xhi], 2).flatten(1)
        yvv = torch.stack([yv, yv], 2).flatten(1)
        xvv = torch.stack([xv, xv], 2).flatten(1)
        idx = (b[:, None, None] * H + ys[:, :, None]) * W + xs[:, None, :]
        ok = yvv[:, :, None] & xvv[:, None, :]
        total += int(torch.unique(idx[ok]).numel())
    return total
