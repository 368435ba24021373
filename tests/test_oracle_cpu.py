"""Pins the CPU oracle (oracle/cpm_oracle.c) against fixtures produced by the reference itself
(tests/golden/make_golden.py) and against torchvision's CPU nms.  No GPU, no /root/reference needed."""
import numpy as np
import pytest
import torch
import torchvision

import oracle

SCALES = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
CASES = [("p7s2", 7, 7, 2, False, [0, 1, 2, 3]), ("p14s2", 14, 14, 2, False, [0, 2]),
         ("p7s0", 7, 7, 0, False, [1]), ("p7s2a", 7, 7, 2, True, [1, 3]), ("p5x3s3", 5, 3, 3, False, [2])]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_roi_align_matches_reference_bit_exact(golden, case):
    g = golden("roi_align")
    tag, ph, pw, sr, al, levels = case
    rois = g["rois"][g[tag + "_sel"]]
    for l in levels:
        feat = g["feat%d" % l]
        out = oracle.roi_align_forward(feat, rois, SCALES[l], ph, pw, sr, al)
        assert np.array_equal(out, g["%s_l%d_out" % (tag, l)])
        B, C, H, W = feat.shape
        gin = oracle.roi_align_backward(g["%s_l%d_gout" % (tag, l)], rois, SCALES[l], ph, pw, B, C, H, W, sr, al)
        assert np.array_equal(gin, g["%s_l%d_gin" % (tag, l)])


def test_roi_align_aligned_negative_size_raises():
    feat = np.zeros((1, 1, 8, 8), np.float32)
    rois = np.array([[0, 5, 5, 2, 2]], np.float32)
    with pytest.raises(RuntimeError):
        oracle.roi_align_forward(feat, rois, 1.0, 2, 2, 2, True, strict_aligned_check=True)
    # the CUDA reference has no such assert (ROIAlign_cuda.cu:211-214): default is non-strict
    oracle.roi_align_forward(feat, rois, 1.0, 2, 2, 2, True)


def test_roi_align_fp64_and_nearest_run():
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((1, 2, 9, 11))
    rois = np.array([[0, 1.5, 2.0, 30.0, 25.0], [0, -4, -4, 50, 50]])
    o64 = oracle.roi_align_forward(feat, rois, 0.25, 4, 4, 2, False)
    o32 = oracle.roi_align_forward(feat.astype(np.float32), rois.astype(np.float32), 0.25, 4, 4, 2, False)
    assert o64.dtype == np.float64 and np.allclose(o64, o32, atol=1e-5)
    near = oracle.roi_align_forward(feat, rois, 0.25, 4, 4, 2, False, interpolation_method=1)
    assert near.shape == (2, 2, 4, 4) and np.isfinite(near).all()


def test_pooler_composition_matches_reference(golden):
    """levels (poolers.py:29-40) + per-level RoIAlign + scatter == the reference Pooler output."""
    g, ra = golden("pooler"), golden("roi_align")
    rois = ra["rois"][g["order"]]
    levels = oracle.level_map(rois, 2, 5)
    assert np.array_equal(levels, g["levels"])
    feats = [ra["feat%d" % l] for l in range(4)]
    for tag, p in (("p7", 7), ("p14", 14)):
        out = np.zeros((len(rois), feats[0].shape[1], p, p), np.float32)
        for l in range(4):
            idx = np.nonzero(levels == l)[0]
            out[idx] = oracle.roi_align_forward(feats[l], rois[idx], SCALES[l], p, p, 2, False)
            B, C, H, W = feats[l].shape
            gin = oracle.roi_align_backward(g[tag + "_gout"][idx], rois[idx], SCALES[l], p, p, B, C, H, W, 2, False)
            assert np.array_equal(gin, g["%s_gin%d" % (tag, l)])
        assert np.array_equal(out, g[tag + "_out"])


def test_level_map_matches_reference(golden):
    g = golden("levels")
    rois = np.concatenate([np.zeros((len(g["boxes"]), 1), np.float32), g["boxes"]], 1)
    assert np.array_equal(oracle.level_map(rois, 2, 5), g["levels"])
    # torch's CUDA scalar division (a * (1/b)) gives the same levels on this set
    assert np.array_equal(oracle.level_map(rois, 2, 5, recip_div=True), g["levels"])


@pytest.mark.parametrize("thr", [0.3, 0.5, 0.7])
def test_nms_matches_torchvision_and_reference_softnms(golden, thr):
    g = golden("nms")
    k = oracle.nms(g["boxes"], g["scores"], thr)
    assert np.array_equal(k, g["nms_keep_%02d" % int(thr * 10)])
    assert np.array_equal(np.sort(k), g["softnms_hard_keep_%02d" % int(thr * 10)])
    km = oracle.nms(g["boxes"], g["scores"], thr, labels=g["labels"])
    assert np.array_equal(km, g["mlnms_keep_%02d" % int(thr * 10)])
    # topk: the sweep stops after topk keeps (ml_nms.cu:134)
    assert np.array_equal(oracle.nms(g["boxes"], g["scores"], thr, labels=g["labels"], topk=17), km[:17])


def test_nms_ties_and_degenerate(golden):
    g = golden("nms")
    assert np.array_equal(oracle.nms(g["boxes"], g["tie_scores"], 0.5), g["tie_keep_05"])
    assert oracle.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).size == 0
    # two identical zero-area boxes: IoU = NaN, comparison false -> both kept
    b = np.array([[3, 3, 3, 3], [3, 3, 3, 3]], np.float32)
    assert oracle.nms(b, np.array([0.9, 0.8], np.float32), 0.5).tolist() == [0, 1]


def test_nms_random_vs_torchvision_live():
    gen = torch.Generator().manual_seed(5)
    for n in (1, 2, 63, 64, 65, 1000):
        xy = torch.rand(n, 2, generator=gen) * 300
        wh = torch.rand(n, 2, generator=gen) * 120 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=gen)
        for thr in (0.3, 0.7):
            ref = torchvision.ops.nms(boxes, scores, thr).numpy()
            assert np.array_equal(oracle.nms(boxes.numpy(), scores.numpy(), thr), ref)


def test_nms_flavors_agree_away_from_threshold(golden):
    """The three fp32 roundings of the union only differ for IoUs within an ulp of the threshold."""
    g = golden("nms")
    k0 = oracle.nms(g["boxes"], g["scores"], 0.5, flavor=oracle.FLAVOR_PLAIN)
    k1 = oracle.nms(g["boxes"], g["scores"], 0.5, flavor=oracle.FLAVOR_TV_CUDA)
    k2 = oracle.nms(g["boxes"], g["scores"], 0.5, flavor=oracle.FLAVOR_ML_CUDA)
    assert np.array_equal(k0, k1) and np.array_equal(k0, k2)


def test_decode_matches_reference(golden):
    g = golden("decode")
    sub = oracle.calc_sub_regions(9, 3, 56)
    assert np.array_equal(np.asarray(sub, np.int32), g["sub_regions"])
    for stage, ratio in enumerate((1.0, 0.5, 0.25)):
        out = oracle.grid_decode(g["logits"], g["boxes"], sub, ratio)
        np.testing.assert_allclose(out, g["stage%d" % stage], rtol=1e-5, atol=1e-3)


def _rpn_case(g, tag):
    pre, post, thr, min_size, fpn_post, training = g[tag + "_params"]
    N = int(g["N"])
    L = 3
    anchors = [np.broadcast_to(g["anchors%d" % l], (N,) + g["anchors%d" % l].shape) for l in range(L)]
    obj = [g["obj%d" % l] for l in range(L)]
    reg = [g["reg%d" % l] for l in range(L)]
    sizes = [tuple(int(v) for v in g["img_wh"])] * N
    return anchors, obj, reg, sizes, dict(pre_nms_top_n=int(pre), post_nms_top_n=int(post), nms_thresh=float(thr),
                                          min_size=float(min_size), fpn_post_nms_top_n=int(fpn_post), training=bool(training))


@pytest.mark.parametrize("tag", ["train", "test", "minsize"])
def test_rpn_selection_oracle_matches_reference(golden, tag):
    """oracle/rpn.py (numpy restatement) against the reference's RPNPostProcessor run on CPU (tests/golden/rpn.npz).
    The sigmoid is taken from torch here (the reference's own op) so that the comparison pins decode, clip, size test,
    NMS and the cross-level selection bit for bit; np.exp vs torch.exp in the decode is within an ulp."""
    from oracle import rpn
    g = golden("rpn")
    anchors, obj, reg, sizes, kw = _rpn_case(g, tag)
    probs = [torch.from_numpy(o).sigmoid().numpy() for o in obj]
    res = rpn.select(anchors, probs, reg, sizes, scores_are_probs=True, **kw)
    for i, (b, s) in enumerate(res):
        rb, rs = g["%s_boxes%d" % (tag, i)], g["%s_scores%d" % (tag, i)]
        assert b.shape == rb.shape
        assert np.array_equal(s, rs)
        np.testing.assert_allclose(b, rb, rtol=2e-6, atol=2e-4)


def test_detection_postprocess_oracle_matches_reference(golden):
    """oracle/detect.py against the reference's CLSPostProcessor.forward run on CPU (tests/golden/detect.npz; the
    softmax is torch's, the reference's own op, so the Python around the NMS is pinned bit for bit)."""
    from oracle import detect
    g = golden("detect")
    counts = g["counts"].tolist()
    img = tuple(int(v) for v in g["img_wh"])
    prob = torch.softmax(torch.from_numpy(g["logits"]), -1).numpy()
    res = detect.cls_postprocess(prob, [g["boxes%d" % i] for i in range(len(counts))], [img] * len(counts),
                                 float(g["params"][0]), float(g["params"][1]))
    for i, (b, s, l) in enumerate(res):
        assert np.array_equal(b, g["res_boxes%d" % i])
        assert np.array_equal(s, g["res_scores%d" % i])
        assert np.array_equal(l, g["res_labels%d" % i])


def test_grid_targets_oracle_matches_reference(golden):
    """oracle/grid_targets.py against GridLossComputation.prepare_target run on CPU (tests/golden/grid_targets.npz):
    the 0/1 maps are identical for the three cascade stages and for radius 2."""
    from oracle import grid_targets as ogt
    g = golden("grid_targets")
    sub = oracle.calc_sub_regions(9, 3, 56)
    for stage, ratio in enumerate((1.0, 0.5, 0.25)):
        t = ogt.prepare_target(g["pos"], g["gt"], ratio, sub)
        assert np.array_equal(t.astype(np.uint8), g["stage%d" % stage])
    assert np.array_equal(ogt.prepare_target(g["pos"], g["gt"], 1.0, sub, pos_radius=2).astype(np.uint8), g["radius2"])


def test_iou_and_matcher_oracle_match_reference(golden):
    """oracle/matcher.py against the reference's boxlist_iou + Matcher run on CPU (tests/golden/matcher.npz): the IoU
    matrix and the three match vectors are bit-identical."""
    from oracle import matcher as om
    g = golden("matcher")
    q = om.box_iou(g["gt"], g["props"])
    assert np.array_equal(q, g["iou"])
    for tag in ("rpn", "head", "grid"):
        hi, lo, allow = g["params_" + tag]
        assert np.array_equal(om.match(g["iou"], hi, lo, bool(allow)), g["match_" + tag])


@pytest.mark.parametrize("phase", ["train", "eval"])
def test_oracle_reproduces_the_model_trace(golden, phase):
    """The op-boundary trace recorded while the reference's own Generalized_RCNN ran (tests/golden/make_dropin_trace.py):
    the oracle's level map + per-level RoIAlign reproduce every Pooler output bit for bit, its NMS every keep list --
    including the levels that received no RoI and the ground-truth boxes the heads append."""
    g = golden("dropin_trace")
    feats = [g["%s_feat%d" % (phase, l)] for l in range(4)]
    for i in range(int(g["%s_counts" % phase][0])):
        key = "%s_pool%d" % (phase, i)
        ph, pw, sr, aligned, nlev = (int(v) for v in g[key + "_cfg"])
        boxes = [g["%s_boxes%d" % (key, j)] for j in range(int(g[key + "_nimg"][0]))]
        rois = np.concatenate([np.concatenate([np.full((len(b), 1), j, np.float32), b], 1) for j, b in enumerate(boxes)], 0)
        levels = oracle.level_map(rois, 2, 5)
        out = np.zeros(g[key + "_out"].shape, np.float32)
        for l in range(nlev):
            idx = np.nonzero(levels == l)[0]
            if len(idx):
                out[idx] = oracle.roi_align_forward(feats[l], rois[idx], float(g[key + "_scales"][l]), ph, pw, sr, bool(aligned))
        assert np.array_equal(out, g[key + "_out"])
    for i in range(int(g["%s_counts" % phase][1])):
        key = "%s_nms%d" % (phase, i)
        assert np.array_equal(oracle.nms(g[key + "_boxes"], g[key + "_scores"], float(g[key + "_thr"][0])), g[key + "_keep"])
