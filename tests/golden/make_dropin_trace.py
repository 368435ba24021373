"""Records what crosses the op boundary (pet.lib.ops / Pooler / GridPostProcessor.get_boxes) when the REFERENCE'S OWN MODEL
runs: Generalized_RCNN built from the published R-50 CPM config, one training forward + backward and one evaluation pass,
on CPU in the build container, with `pet.lib.ops._C` bound to the reference's unmodified CPU RoIAlign (oracle/_ref) and
its CUDA-only ml_nms bound to the label-gated NMS the nms.npz fixture pins to ml_nms.cu.

    python tests/golden/make_dropin_trace.py        # writes tests/golden/dropin_trace.npz (seeded, deterministic)

/root/reference does not exist on the GPU box, so the model itself cannot run there; the committed trace is what the GPU
test (tests/test_gpu_parity.py::test_dropin_trace_*) replays through cpm_r_cnn_b200 call by call: every Pooler forward of
the iteration (cls 7x7, three grid stages 14x14, re-score 7x7) with the model's own proposals -- empty levels, ground-truth
boxes and duplicates included --, the pooled gradients autograd handed back and the feature gradient the backbone received,
every RPN / detection NMS call, every grid decode.  Sizes are shrunk (FPN.DIM 16, small images, few proposals) to keep the
fixture small; the SEQUENCE and the argument conventions are the model's.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import oracle  # noqa: E402
from oracle import build_ref  # noqa: E402
import make_golden as mg  # noqa: E402

CFG = "/root/reference/cfgs/rcnn/mscoco/grid_cascade/iou_helper/rescore/e2e_grid_cascade@567_rcnn_R-50-FPN_2x.yaml"
IMG_SIZES = ((160, 224), (150, 200))          # (h, w) of the two training images


def main():
    ref = build_ref.load("pet_ref_cpu")

    class RefOps(object):
        roi_align_forward = staticmethod(ref.roi_align_forward)
        roi_align_backward = staticmethod(ref.roi_align_backward)

        @staticmethod
        def ml_nms(dets, scores, labels, thr, topk):
            k = oracle.nms(dets.numpy(), scores.numpy(), float(thr), labels=labels.numpy(), flavor=oracle.FLAVOR_ML_CUDA,
                           topk=int(topk))
            return torch.as_tensor(k, dtype=torch.int64)

    mg.install_shims(RefOps)
    from pet.rcnn.core.config import cfg, merge_cfg_from_file, merge_cfg_from_list
    merge_cfg_from_file(CFG)
    merge_cfg_from_list(["FPN.DIM", 16, "RPN.PRE_NMS_TOP_N_TRAIN", 200, "RPN.POST_NMS_TOP_N_TRAIN", 100,
                         "RPN.FPN_POST_NMS_TOP_N_TRAIN", 100, "RPN.PRE_NMS_TOP_N_TEST", 100, "RPN.POST_NMS_TOP_N_TEST", 50,
                         "RPN.FPN_POST_NMS_TOP_N_TEST", 8, "GRID_RCNN.BATCH_SIZE_PER_IMAGE", 32,
                         "GRID_RCNN.MAX_SAMPLE_NUM_GRID", 8, "DEVICE", "cpu"])
    from pet.rcnn.modeling.model_builder import Generalized_RCNN
    from pet.utils.data.structures.bounding_box import BoxList
    import pet.rcnn.utils.poolers as rp
    import pet.utils.data.structures.boxlist_ops as bo_old
    import pet.lib.ops.boxlist_ops as bo_new
    import pet.rcnn.modeling.grid_cascade_rcnn.inference as ginf

    trace = {}
    state = {"phase": "train", "pool": 0, "nms": 0, "mlnms": 0, "dec": 0, "feats": None, "pending": []}

    # ---- Pooler.forward ----
    orig_pool = rp.Pooler.forward

    def pool_forward(self, x, boxes):
        ph = state["phase"]
        i = state["pool"]
        state["pool"] += 1
        key = "%s_pool%d" % (ph, i)
        xs = list(x)[:len(self.poolers)]
        if state["feats"] is None:
            # the poolers read the FPN maps through one autograd view per level, so that its .grad is the gradient the
            # POOLERS send back (the RPN head reads the same maps and adds its own)
            state["feats"] = xs
            state["taps"] = [t.view_as(t) for t in xs]
            for l, t in enumerate(state["taps"]):
                trace["%s_feat%d" % (ph, l)] = t.detach().numpy().copy()
                if t.requires_grad:
                    t.retain_grad()
        else:
            assert all(a is b for a, b in zip(state["feats"], xs)), "poolers of one pass share the FPN maps"
        x = state["taps"] + list(x)[len(self.poolers):]
        trace[key + "_cfg"] = np.array([self.output_size[0], self.output_size[1], self.poolers[0].sampling_ratio,
                                        int(self.poolers[0].aligned), len(self.poolers)], dtype=np.int64)
        trace[key + "_scales"] = np.array([p.spatial_scale for p in self.poolers], dtype=np.float64)
        trace[key + "_nimg"] = np.array([len(boxes)], dtype=np.int64)
        for j, b in enumerate(boxes):
            trace["%s_boxes%d" % (key, j)] = b.bbox.detach().numpy().copy()
            trace["%s_size%d" % (key, j)] = np.array(b.size, dtype=np.int64)
        out = orig_pool(self, x, boxes)
        trace[key + "_out"] = out.detach().numpy().copy()
        if out.requires_grad:
            out.register_hook(lambda g, key=key: trace.__setitem__(key + "_gout", g.detach().numpy().copy()))
        return out

    rp.Pooler.forward = pool_forward

    # ---- nms (pet.lib.ops.nms = torchvision.ops.nms) and ml_nms ----
    def wrap_nms(orig):
        def f(boxes, scores, thr):
            keep = orig(boxes, scores, thr)
            key = "%s_nms%d" % (state["phase"], state["nms"])
            state["nms"] += 1
            trace[key + "_boxes"], trace[key + "_scores"] = boxes.detach().numpy().copy(), scores.detach().numpy().copy()
            trace[key + "_thr"], trace[key + "_keep"] = np.array([thr], dtype=np.float64), keep.numpy().copy()
            return keep
        return f

    def wrap_ml_nms(orig):
        def f(boxes, scores, labels, thr, topk):
            keep = orig(boxes, scores, labels, thr, topk)
            key = "%s_mlnms%d" % (state["phase"], state["mlnms"])
            state["mlnms"] += 1
            trace[key + "_boxes"], trace[key + "_scores"] = boxes.detach().numpy().copy(), scores.detach().numpy().copy()
            trace[key + "_labels"] = labels.numpy().copy()
            trace[key + "_args"], trace[key + "_keep"] = np.array([thr, topk], dtype=np.float64), keep.numpy().copy()
            return keep
        return f

    bo_old._box_nms = wrap_nms(bo_old._box_nms)
    bo_new._box_nms = wrap_nms(bo_new._box_nms)
    bo_old._box_ml_nms = wrap_ml_nms(bo_old._box_ml_nms)
    bo_new._box_ml_nms = wrap_ml_nms(bo_new._box_ml_nms)

    # ---- GridPostProcessor.get_boxes ----
    orig_dec = ginf.GridPostProcessor.get_boxes

    def get_boxes(self, proposals, grid_pred, is_train=False, **kw):
        out = orig_dec(self, proposals, grid_pred, is_train, **kw)
        key = "%s_dec%d" % (state["phase"], state["dec"])
        state["dec"] += 1
        # decoding is per RoI: the first DEC_ROWS RoIs of the call are recorded (28 KB of logits each)
        DEC_ROWS = 6
        trace[key + "_boxes"] = proposals.bbox.detach().numpy()[:DEC_ROWS].copy()
        trace[key + "_pred"] = grid_pred.detach().numpy()[:DEC_ROWS].copy()
        trace[key + "_cfg"] = np.array([int(is_train), int(getattr(self, "stage", -1)) if hasattr(self, "stage") else -1],
                                       dtype=np.int64)
        trace[key + "_out"] = out.detach().numpy()[:DEC_ROWS].copy()
        trace[key + "_size"] = np.array(proposals.size, dtype=np.int64)
        return out

    ginf.GridPostProcessor.get_boxes = get_boxes

    # ---- the model, one training iteration ----
    torch.manual_seed(0)
    model = Generalized_RCNN(is_train=True)
    g = torch.Generator().manual_seed(1)
    imgs = [torch.randn(3, h, w, generator=g) for (h, w) in IMG_SIZES]
    targets = []
    for (h, w) in IMG_SIZES:
        n = 4
        x1 = torch.rand(n, generator=g) * (w * 0.5)
        y1 = torch.rand(n, generator=g) * (h * 0.5)
        bw = 20 + torch.rand(n, generator=g) * (w * 0.4)
        bh = 20 + torch.rand(n, generator=g) * (h * 0.4)
        b = torch.stack([x1, y1, (x1 + bw).clamp(max=w - 1), (y1 + bh).clamp(max=h - 1)], 1)
        t = BoxList(b, (w, h), mode="xyxy")
        t.add_field("labels", torch.randint(1, 81, (n,), generator=g))
        targets.append(t)
    model.train()
    t0 = time.time()
    out = model(imgs, targets)
    loss = sum(out["losses"].values())
    loss.backward()
    for k, v in out["losses"].items():
        trace["train_loss_" + k] = np.array([float(v)], dtype=np.float64)
    for l, t in enumerate(state["taps"]):
        trace["train_gfeat%d" % l] = t.grad.detach().numpy().copy()      # what the FPN receives from the 5 poolers
    trace["train_counts"] = np.array([state["pool"], state["nms"], state["mlnms"], state["dec"]], dtype=np.int64)
    print("train: %.1fs, %d pooler / %d nms / %d decode calls" % (time.time() - t0, state["pool"], state["nms"], state["dec"]))

    # ---- evaluation pass (box_net) ----
    state.update({"phase": "eval", "pool": 0, "nms": 0, "mlnms": 0, "dec": 0, "feats": None})
    torch.manual_seed(0)
    emodel = Generalized_RCNN(is_train=False)       # model_builder.py:23-28: the test-time model owns the input normalisation
    emodel.load_state_dict(model.state_dict(), strict=False)
    emodel.eval()
    with torch.no_grad():
        res = emodel.box_net([imgs[0] * 60.0 + 110.0])
    conv, result = res if isinstance(res, tuple) else (None, res)
    r0 = result[0]
    trace["eval_result_bbox"] = r0.bbox.numpy().copy()
    trace["eval_result_scores"] = r0.get_field("scores").numpy().copy()
    trace["eval_result_labels"] = r0.get_field("labels").numpy().copy()
    trace["eval_counts"] = np.array([state["pool"], state["nms"], state["mlnms"], state["dec"]], dtype=np.int64)
    print("eval: %d pooler / %d nms / %d ml_nms / %d decode calls, %d detections" % (
        state["pool"], state["nms"], state["mlnms"], state["dec"], len(r0)))
    path = os.path.join(HERE, "dropin_trace.npz")
    np.savez_compressed(path, **trace)
    print("wrote %s (%.2f MB, %d arrays)" % (path, os.path.getsize(path) / 1e6, len(trace)))


if __name__ == "__main__":
    main()
