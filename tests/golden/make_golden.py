"""Generates tests/golden/*.npz from the REFERENCE itself (run in the build container only).

Sources of truth used here (nothing from this repo's product or oracle code):
  * oracle/_ref/pet_ref_cpu.so  = unmodified /root/reference/pet/lib/ops/csrc/ROIAlign/ROIAlign_cpu.cpp and
    NMS/soft_nms.cpp compiled in place by oracle/build_ref.py;
  * the reference's own Python (imported from /root/reference with the four shims of SURVEY.md 8c):
    pet/rcnn/utils/poolers.py (LevelMapper, Pooler), pet/utils/data/structures/bounding_box.py (BoxList),
    pet/rcnn/modeling/grid_cascade_rcnn/inference.py (GridPostProcessor.get_boxes);
  * torchvision 0.26 CPU `nms` (the third-party op behind pet/lib/ops/nms.py:2,10).

    python tests/golden/make_golden.py          # rewrites the fixtures (seeded, deterministic)

/root/reference does not exist on the GPU box; the committed .npz files are what travels.
"""
import os
import sys
import types

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import build_ref  # noqa: E402

LEVEL_SHAPES = [(50, 84), (25, 42), (13, 21), (7, 11)]       # a 200x336 image at strides 4..32
SCALES = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
IMG_H, IMG_W = 200, 336


def install_shims(ref_ops):
    apex = types.ModuleType("apex")
    amp = types.ModuleType("apex.amp")
    amp.float_function = lambda f: f
    apex.amp = amp
    sys.modules["apex"], sys.modules["apex.amp"] = apex, amp

    class _FakeC(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            if hasattr(ref_ops, k):
                return getattr(ref_ops, k)
            return lambda *a, **kw: (_ for _ in ()).throw(RuntimeError("not built: _C." + k))

    sys.modules["pet.lib.ops._C"] = _FakeC("pet.lib.ops._C")
    for n, t in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, n):
            setattr(np, n, t)
    import yaml
    _yl = yaml.load
    yaml.load = lambda s, Loader=None: _yl(s, Loader=Loader or yaml.FullLoader)
    pc = types.ModuleType("pycocotools")
    sys.modules["pycocotools"] = pc
    for sub in ("mask", "coco", "cocoeval"):
        m = types.ModuleType("pycocotools." + sub)
        sys.modules["pycocotools." + sub] = m
        setattr(pc, sub, m)
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools.cocoeval"].COCOeval = object
    # decode calls .cuda()/.get_device() unconditionally (inference.py:193,278)
    torch.Tensor.cuda = lambda self, *a, **k: self


def coco_like_rois(gen, n, img_h, img_w, n_img):
    """SURVEY.md 8(d) RoI generator (sqrt(area) log-uniform, aspect log-uniform) + adversarial rows."""
    s = torch.exp(torch.empty(n).uniform_(np.log(8.0), np.log(0.9 * min(img_h, img_w)), generator=gen))
    ar = torch.exp(torch.empty(n).uniform_(np.log(0.5), np.log(2.0), generator=gen))
    w = (s * torch.sqrt(ar)).clamp(max=img_w - 1)
    h = (s / torch.sqrt(ar)).clamp(max=img_h - 1)
    x1 = torch.rand(n, generator=gen) * (img_w - 1 - w)
    y1 = torch.rand(n, generator=gen) * (img_h - 1 - h)
    img = torch.randint(0, n_img, (n,), generator=gen).float()
    rois = torch.stack([img, x1, y1, x1 + w, y1 + h], 1)
    adv = torch.tensor([
        [0, 0, 0, img_w - 1, img_h - 1],           # whole image
        [1, -20, -30, 40, 50],                     # sticks out top-left
        [0, img_w - 30, img_h - 20, img_w + 60, img_h + 45],   # sticks out bottom-right
        [1, 100, 100, 100, 100],                   # zero area
        [0, 50.25, 60.5, 50.75, 61.0],             # sub-pixel
        [1, 10, 10, 12, 150],                      # thin, tall
        [0, 5, 90, 330, 93],                       # thin, wide
        [1, img_w + 200, img_h + 200, img_w + 300, img_h + 300],   # fully outside
    ], dtype=torch.float32)
    return torch.cat([rois, adv], 0)


def gen_roi_align(ref):
    g = torch.Generator().manual_seed(1234)
    B, C = 2, 3
    feats = [torch.randn(B, C, h, w, generator=g) for (h, w) in LEVEL_SHAPES]
    rois = coco_like_rois(g, 32, IMG_H, IMG_W, B)
    out = {"rois": rois.numpy()}
    for l, f in enumerate(feats):
        out["feat%d" % l] = f.numpy()
    # (tag, PH, PW, sampling_ratio, aligned, levels)
    cases = [("p7s2", 7, 7, 2, False, [0, 1, 2, 3]), ("p14s2", 14, 14, 2, False, [0, 2]),
             ("p7s0", 7, 7, 0, False, [1]), ("p7s2a", 7, 7, 2, True, [1, 3]), ("p5x3s3", 5, 3, 3, False, [2])]
    valid_aligned = (rois[:, 3] >= rois[:, 1]) & (rois[:, 4] >= rois[:, 2])
    for tag, ph, pw, sr, al, levels in cases:
        r = rois[valid_aligned] if al else rois
        out["%s_sel" % tag] = np.nonzero(valid_aligned.numpy() if al else np.ones(len(rois), bool))[0]
        for l in levels:
            f = feats[l]
            o = ref.roi_align_forward(f, r, SCALES[l], ph, pw, sr, al, 0)
            go = torch.randn(o.shape, generator=g)
            gi = ref.roi_align_backward(go, r, SCALES[l], ph, pw, B, C, f.shape[2], f.shape[3], sr, al, 0)
            out["%s_l%d_out" % (tag, l)] = o.numpy()
            out["%s_l%d_gout" % (tag, l)] = go.numpy()
            out["%s_l%d_gin" % (tag, l)] = gi.numpy()
    np.savez_compressed(os.path.join(HERE, "roi_align.npz"), **out)
    return feats, rois


def gen_pooler(feats, rois):
    """The reference Pooler (poolers.py:43-132) end to end on CPU, its ROIAlign bound to the reference CPU op."""
    from pet.rcnn.utils.poolers import Pooler, LevelMapper
    from pet.utils.data.structures.bounding_box import BoxList
    out = {}
    boxlists = []
    for i in range(feats[0].shape[0]):
        sel = rois[:, 0] == i
        boxlists.append(BoxList(rois[sel][:, 1:].clone(), (IMG_W, IMG_H), mode="xyxy"))
    order = torch.cat([torch.nonzero(rois[:, 0] == i).squeeze(1) for i in range(feats[0].shape[0])])
    out["order"] = order.numpy()
    lm = LevelMapper(2, 5)
    out["levels"] = lm(boxlists).numpy()
    for tag, res in (("p7", (7, 7)), ("p14", (14, 14))):
        pooler = Pooler("ROIAlign", res, SCALES, 2)
        fs = [f.clone().requires_grad_(True) for f in feats]
        y = pooler(fs, boxlists)
        g = torch.Generator().manual_seed(7)
        go = torch.randn(y.shape, generator=g)
        y.backward(go)
        out[tag + "_out"] = y.detach().numpy()
        out[tag + "_gout"] = go.numpy()
        for l, f in enumerate(fs):
            out["%s_gin%d" % (tag, l)] = f.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "pooler.npz"), **out)


def gen_levels():
    from pet.rcnn.utils.poolers import LevelMapper
    from pet.utils.data.structures.bounding_box import BoxList
    g = torch.Generator().manual_seed(99)
    n = 4000
    s = torch.exp(torch.empty(n).uniform_(np.log(2.0), np.log(1500.0), generator=g))
    ar = torch.exp(torch.empty(n).uniform_(np.log(0.2), np.log(5.0), generator=g))
    w, h = s * torch.sqrt(ar), s / torch.sqrt(ar)
    x1, y1 = torch.rand(n, generator=g) * 800, torch.rand(n, generator=g) * 800
    boxes = torch.stack([x1, y1, x1 + w, y1 + h], 1)
    # exact canonical boundaries: sqrt(area) = 112, 224, 448 with the +1 convention
    exact = torch.tensor([[0, 0, 111, 111], [0, 0, 223, 223], [0, 0, 447, 447], [0, 0, 55, 55], [3, 4, 3, 4]],
                         dtype=torch.float32)
    boxes = torch.cat([boxes, exact], 0)
    lv = LevelMapper(2, 5)([BoxList(boxes, (2000, 2000), mode="xyxy")])
    np.savez_compressed(os.path.join(HERE, "levels.npz"), boxes=boxes.numpy(), levels=lv.numpy())


def gen_nms(ref):
    g = torch.Generator().manual_seed(4321)
    out = {}
    n = 700
    base = torch.rand(n // 3 + 1, 2, generator=g) * 500
    wh = torch.rand(n // 3 + 1, 2, generator=g) * 150 + 4
    b0 = torch.cat([base, base + wh], 1)
    boxes = torch.cat([b0 + torch.randn(b0.shape, generator=g) * s for s in (0.0, 4.0, 10.0)], 0)[:n]
    boxes[5] = torch.tensor([10., 10., 10., 10.])      # zero-area pair: IoU = 0/0 = NaN -> never suppressed
    boxes[6] = torch.tensor([10., 10., 10., 10.])
    perm = torch.randperm(n, generator=g)
    scores = (perm.float() + 0.5) / n                  # tie-free
    labels = torch.randint(1, 6, (n,), generator=g)
    out.update(boxes=boxes.numpy(), scores=scores.numpy(), labels=labels.numpy())
    for thr in (0.3, 0.5, 0.7):
        keep = torchvision.ops.nms(boxes, scores, thr)
        out["nms_keep_%02d" % int(thr * 10)] = keep.numpy()
        # reference soft_nms in hard mode (NMS/soft_nms.cpp) is a second CPU oracle for the keep SET
        _, _, sk = ref.soft_nms(boxes.clone(), scores.clone(), 0.5, thr, 1e-4, 0)
        out["softnms_hard_keep_%02d" % int(thr * 10)] = np.sort(sk.numpy())
        # label-gated ml_nms = per-class torchvision nms merged by descending score
        parts = []
        for c in labels.unique():
            idx = torch.nonzero(labels == c).squeeze(1)
            parts.append(idx[torchvision.ops.nms(boxes[idx], scores[idx], thr)])
        allk = torch.cat(parts)
        allk = allk[torch.argsort(scores[allk], descending=True)]
        out["mlnms_keep_%02d" % int(thr * 10)] = allk.numpy()
    # ties: equal scores -> torchvision's stable sort keeps the lower index first
    tscores = torch.round(scores * 20) / 20
    out["tie_scores"] = tscores.numpy()
    out["tie_keep_05"] = torchvision.ops.nms(boxes, tscores, 0.5).numpy()
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **out)


def gen_decode():
    from pet.rcnn.modeling.grid_cascade_rcnn.inference import GridPostProcessor
    from pet.utils.data.structures.bounding_box import BoxList
    g = torch.Generator().manual_seed(2468)
    R = 24
    logits = torch.randn(R, 9, 28, 28, generator=g) * 2
    xy = torch.rand(R, 2, generator=g) * 400
    wh = torch.rand(R, 2, generator=g) * 300 + 8
    boxes = torch.cat([xy, xy + wh], 1)
    out = {"logits": logits.numpy(), "boxes": boxes.numpy()}
    torch.Tensor.get_device = lambda self: "cpu"
    for stage in range(3):
        pp = GridPostProcessor(stage, 9, 14)
        bl = BoxList(boxes.clone(), (1000, 1000), mode="xyxy")
        res = pp.get_boxes(bl, logits.clone(), False)
        out["stage%d" % stage] = res.numpy()
        out["sub_regions"] = np.asarray(pp.sub_regions, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **out)


def rpn_inputs(seed, N=2, A=3, level_shapes=((25, 42), (13, 21), (7, 11)), strides=(8, 16, 32), img=(336, 200)):
    """Seeded RPN head outputs + grid anchors ((H, W, A) order, what the reference's AnchorGenerator emits) for a
    200x336 image; regression scaled so that decoded boxes overlap heavily (NMS has work to do)."""
    g = torch.Generator().manual_seed(seed)
    anchors, obj, reg = [], [], []
    for (h, w), st in zip(level_shapes, strides):
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        ctr = torch.stack([xs, ys], -1).reshape(-1, 1, 2) * st + st / 2
        half = torch.tensor([[2.0 * st, 1.0 * st], [1.4 * st, 1.4 * st], [1.0 * st, 2.0 * st]][:A]).reshape(1, A, 2)
        a = torch.cat([ctr - half, ctr + half - 1], -1).reshape(-1, 4)
        anchors.append(a)
        obj.append(torch.randn(N, A, h, w, generator=g) * 2)
        reg.append(torch.randn(N, 4 * A, h, w, generator=g) * 0.4)
    return anchors, obj, reg, img


def gen_rpn():
    """RPNPostProcessor (pet/rcnn/modeling/rpn/inference.py) end to end on CPU: training (top-k over the batch) and
    testing (per image) behaviour, min_size 0 and 16."""
    from pet.rcnn.modeling.rpn.inference import RPNPostProcessor
    from pet.rcnn.utils.box_coder import BoxCoder
    from pet.utils.data.structures.bounding_box import BoxList
    anchors, obj, reg, img = rpn_inputs(2024)
    N = obj[0].shape[0]
    out = {"img_wh": np.array(img, np.int64), "N": np.array(N)}
    for l in range(len(obj)):
        out["anchors%d" % l], out["obj%d" % l], out["reg%d" % l] = anchors[l].numpy(), obj[l].numpy(), reg[l].numpy()
    cases = {"train": (300, 60, 0.7, 0, 100, True), "test": (200, 50, 0.7, 0, 80, False), "minsize": (300, 60, 0.6, 16, 100, False)}
    for tag, (pre, post, thr, min_size, fpn_post, training) in cases.items():
        pp = RPNPostProcessor(pre, post, thr, min_size, BoxCoder(weights=(1.0, 1.0, 1.0, 1.0)), fpn_post, True)
        pp.train(training)
        alist = [[BoxList(a.clone(), img, mode="xyxy") for a in anchors] for _ in range(N)]
        res = pp(alist, [o.clone() for o in obj], [r.clone() for r in reg])
        out[tag + "_params"] = np.array([pre, post, thr, min_size, fpn_post, int(training)], np.float64)
        for i, bl in enumerate(res):
            out["%s_boxes%d" % (tag, i)] = bl.bbox.numpy()
            out["%s_scores%d" % (tag, i)] = bl.get_field("objectness").numpy()
    np.savez_compressed(os.path.join(HERE, "rpn.npz"), **out)


def gen_detect():
    """CLSPostProcessor.forward (grid_cascade_rcnn/inference.py:59-124) on CPU.  `_C.ml_nms` has no CPU build
    (ml_nms.h:38), so for this fixture only it is bound to the oracle's label-gated NMS (itself pinned to the reference's
    kernel by nms.npz); what the fixture pins is the reference's Python around it: softmax, per-class repeat, clip, score
    gate, numpy label build, BoxList indexing."""
    import oracle
    import pet.lib.ops.boxlist_ops as blo
    from pet.rcnn.modeling.grid_cascade_rcnn.inference import CLSPostProcessor
    from pet.utils.data.structures.bounding_box import BoxList

    def ml_nms_cpu(boxes, scores, labels, thr, topk):
        keep = oracle.nms(boxes.numpy(), scores.numpy(), float(thr), labels=labels.numpy(), topk=int(topk),
                          flavor=oracle.FLAVOR_ML_CUDA)
        return torch.from_numpy(keep)
    blo._box_ml_nms = ml_nms_cpu
    g = torch.Generator().manual_seed(99)
    C, img = 21, (336, 200)
    counts = [37, 0, 52]
    boxes = []
    for n in counts:
        b = coco_like_rois(g, max(n - 8, 0), img[1], img[0], 1)[:n, 1:] if n else torch.zeros(0, 4)
        boxes.append(b)
    logits = torch.randn(sum(counts), C, generator=g) * 2.5
    out = {"logits": logits.numpy(), "counts": np.array(counts), "img_wh": np.array(img), "params": np.array([0.03, 0.3])}
    for i, b in enumerate(boxes):
        out["boxes%d" % i] = b.numpy()
    pp = CLSPostProcessor(0.03, 0.3)
    res = pp(logits.clone(), [BoxList(b.clone(), img, mode="xyxy") for b in boxes])
    for i, bl in enumerate(res):
        out["res_boxes%d" % i] = bl.bbox.numpy()
        out["res_scores%d" % i] = bl.get_field("scores").numpy()
        out["res_labels%d" % i] = bl.get_field("labels").numpy()
    # rescoring branch (:62-76)
    bls = []
    for b in boxes:
        bl = BoxList(b.clone(), img, mode="xyxy")
        bl.add_field("scores", torch.rand(len(b), generator=g))
        bl.add_field("labels", torch.randint(1, C, (len(b),), generator=g))
        bls.append(bl)
    rl = torch.randn(counts[0], C, generator=g)
    out["rescore_logits"] = rl.numpy()
    out["rescore_in_scores"] = bls[0].get_field("scores").numpy()
    out["rescore_in_labels"] = bls[0].get_field("labels").numpy()
    r = CLSPostProcessor(0.03, 0.3)(rl, [bls[0]], rescore=True)
    out["rescore_out"] = r[0].get_field("scores").numpy()
    np.savez_compressed(os.path.join(HERE, "detect.npz"), **out)


def gen_grid_targets():
    """GridLossComputation.prepare_target (grid_cascade_rcnn/loss.py:178-258) for the three cascade stages, on boxes
    that include tiny RoIs (skipped), ground truth far outside the extended RoI (disks clipped / off the map) and exact
    half-pixel cases."""
    from pet.rcnn.modeling.grid_cascade_rcnn.loss import GridLossComputation
    g = torch.Generator().manual_seed(7)
    R = 48
    pos = coco_like_rois(g, R - 8, 800, 1344, 1)[:R, 1:].clone()
    jitter = (torch.rand(R, 4, generator=g) - 0.5) * (pos[:, 2:] - pos[:, :2]).repeat(1, 2) * 0.5
    gt = pos + jitter
    pos[5] = torch.tensor([100.0, 100.0, 102.0, 140.0])      # narrower than the grid -> ignored
    gt[9] = pos[9] + 400.0                                    # ground truth far away: disks off the map
    gt[11] = pos[11].clone()                                  # identical boxes
    out = {"pos": pos.numpy(), "gt": gt.numpy()}
    for stage in range(3):
        lc = GridLossComputation(stage, 1.0, None, 1, 9, 14)
        lc.pos_result = (pos.clone(), gt.clone())
        out["stage%d" % stage] = lc.prepare_target(None, None).numpy().astype(np.uint8)
    lc = GridLossComputation(0, 1.0, None, 2, 9, 14)
    lc.pos_result = (pos.clone(), gt.clone())
    out["radius2"] = lc.prepare_target(None, None).numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "grid_targets.npz"), **out)


def gen_matcher():
    """boxlist_iou + Matcher (pet/utils/data/structures/boxlist_ops.py:123-158, pet/rcnn/utils/matcher.py) on CPU:
    RPN thresholds with low-quality matches, head thresholds without; ground truth incl. a duplicate box (exact ties) and
    a box that overlaps nothing."""
    from pet.rcnn.utils.matcher import Matcher
    from pet.utils.data.structures.bounding_box import BoxList
    from pet.utils.data.structures.boxlist_ops import boxlist_iou
    g = torch.Generator().manual_seed(31)
    img = (1344, 800)
    gt = coco_like_rois(g, 12, img[1], img[0], 1)[:12, 1:].clone()
    gt[5] = gt[2]                                             # duplicate ground truth: ties in the column arg-max
    gt[7] = torch.tensor([1300.0, 760.0, 1343.0, 799.0])
    props = torch.cat([coco_like_rois(g, 600, img[1], img[0], 1)[:600, 1:],
                       gt[:6] + (torch.rand(6, 4, generator=g) - 0.5) * 6, gt[2:4].clone()], 0)
    q = boxlist_iou(BoxList(gt, img), BoxList(props, img))
    out = {"gt": gt.numpy(), "props": props.numpy(), "iou": q.numpy()}
    for tag, (hi, lo, allow) in {"rpn": (0.7, 0.3, True), "head": (0.5, 0.5, False), "grid": (0.5, 0.5, True)}.items():
        out["match_" + tag] = Matcher(hi, lo, allow)(q.clone()).numpy()
        out["params_" + tag] = np.array([hi, lo, float(allow)])
    np.savez_compressed(os.path.join(HERE, "matcher.npz"), **out)


def gen_grid_forward():
    """GridPostProcessor.forward (inference.py:145-187): training (ground-truth rows filtered out, ground truth appended;
    two images, one of them with every proposal equal to a ground truth) and testing (stage 0 on two images; last stage on
    one image with the IoU head's score merge)."""
    from pet.rcnn.modeling.grid_cascade_rcnn.inference import GridPostProcessor
    from pet.utils.data.structures.bounding_box import BoxList
    from pet.rcnn.core.config import cfg
    torch.Tensor.get_device = lambda self: "cpu"
    # the published CPM models' switches (e.g. the R-50 yaml lines 21-34)
    cfg.GRID_RCNN.FUSED_ON = False
    cfg.GRID_RCNN.IOU_HELPER = True
    cfg.GRID_RCNN.IOU_HELPER_MERGE = True
    g = torch.Generator().manual_seed(97531)

    def boxes(n):
        xy = torch.rand(n, 2, generator=g) * 400
        wh = torch.rand(n, 2, generator=g) * 300 + 8
        return torch.cat([xy, xy + wh], 1)

    out = {}
    # ---- training, stage 1
    gts = [boxes(3), boxes(2)]
    props = [torch.cat([boxes(5), gts[0][:2], boxes(1)[:, [0, 1, 2, 3]]], 0), gts[1].clone()]
    props[0][6, 0] = gts[0][2, 0]                       # one coordinate equal to a ground truth's: kept (sum stays > 0)
    counts = [p.shape[0] for p in props]
    logits = torch.randn(sum(counts), 9, 28, 28, generator=g) * 2
    bls, tgts = [], []
    for p, t in zip(props, gts):
        bl = BoxList(p.clone(), (1000, 800), mode="xyxy")
        bl.add_field("labels", torch.randint(1, 81, (p.shape[0],), generator=g))
        bl.add_field("objectness", torch.rand(p.shape[0], generator=g))
        tl = BoxList(t.clone(), (1000, 800), mode="xyxy")
        tl.add_field("labels", torch.randint(1, 81, (t.shape[0],), generator=g))
        bls.append(bl)
        tgts.append(tl)
    out["train_logits"] = logits.numpy()
    for i in range(2):
        out["train_prop%d" % i] = props[i].numpy()
        out["train_prop_labels%d" % i] = bls[i].get_field("labels").numpy()
        out["train_prop_obj%d" % i] = bls[i].get_field("objectness").numpy()
        out["train_gt%d" % i] = gts[i].numpy()
        out["train_gt_labels%d" % i] = tgts[i].get_field("labels").numpy()
    res = GridPostProcessor(1, 9, 14).forward({"unfused": logits.clone()}, bls, None, True, tgts)
    for i, r in enumerate(res):
        out["train_out_bbox%d" % i] = r.bbox.numpy()
        out["train_out_labels%d" % i] = r.get_field("labels").numpy()
        out["train_out_obj%d" % i] = r.get_field("objectness").numpy()
    # ---- testing, stage 0, two images
    props = [boxes(6), boxes(4)]
    logits = torch.randn(10, 9, 28, 28, generator=g) * 2
    bls = []
    for p in props:
        bl = BoxList(p.clone(), (1000, 800), mode="xyxy")
        bl.add_field("scores", torch.rand(p.shape[0], generator=g))
        bls.append(bl)
    out["test0_logits"] = logits.numpy()
    for i in range(2):
        out["test0_prop%d" % i] = props[i].numpy()
        out["test0_scores%d" % i] = bls[i].get_field("scores").numpy()
    res = GridPostProcessor(0, 9, 14).forward({"unfused": logits.clone()}, bls, None, False)
    for i, r in enumerate(res):
        out["test0_out_bbox%d" % i] = r.bbox.numpy()
        out["test0_out_scores%d" % i] = r.get_field("scores").numpy()
    # ---- testing, last stage, one image, score *= iou probability
    p = boxes(7)
    logits = torch.randn(7, 9, 28, 28, generator=g) * 2
    iou = torch.rand(7, 2, generator=g)
    bl = BoxList(p.clone(), (1000, 800), mode="xyxy")
    bl.add_field("scores", torch.rand(7, generator=g))
    out.update(test2_logits=logits.numpy(), test2_prop=p.numpy(), test2_scores=bl.get_field("scores").numpy(),
               test2_iou=iou.numpy())
    r = GridPostProcessor(2, 9, 14).forward({"unfused": logits.clone()}, [bl], iou.clone(), False)[0]
    out["test2_out_bbox"] = r.bbox.numpy()
    out["test2_out_scores"] = r.get_field("scores").numpy()
    np.savez_compressed(os.path.join(HERE, "grid_forward.npz"), **out)


def main():
    build_ref.build(cuda=False)
    ref = build_ref.load("pet_ref_cpu")
    install_shims(ref)
    only = [a for a in sys.argv[1:] if a in ("rpn", "detect", "grid_targets", "matcher", "grid_forward")]
    if only:
        for a in only:
            {"rpn": gen_rpn, "detect": gen_detect, "grid_targets": gen_grid_targets, "matcher": gen_matcher,
             "grid_forward": gen_grid_forward}[a]()
            print(a + ".npz", os.path.getsize(os.path.join(HERE, a + ".npz")))
        return
    gen_rpn()
    gen_detect()
    gen_grid_targets()
    gen_matcher()
    gen_grid_forward()
    feats, rois = gen_roi_align(ref)
    gen_pooler(feats, rois)
    gen_levels()
    gen_nms(ref)
    gen_decode()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
