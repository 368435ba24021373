"""The drop-in boundary without a GPU: libcpm_ops.so loads and exports every symbol include/cpm_ops.h declares, the
Python op layer keeps the reference's names and signatures, the product path has no CPU fallback and
never touches oracle/, and the host-side logic (BoxList, RoI table, argument checks) behaves like the reference's."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

import cpm_r_cnn_b200 as ops
from cpm_r_cnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cpm_ops.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"CPM_API\s+[\w\s\*]+?\b(cpm_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 12, names
    handle = ctypes.CDLL(_lib.LIB_PATH)                # loads without a GPU (cudart is linked statically)
    for n in names:
        assert hasattr(handle, n), "libcpm_ops.so does not export %s" % n
    assert sorted(_lib.exported_symbols()) == names    # the ctypes binding covers exactly the header
    lib = _lib.lib()
    assert lib.cpm_version() >= 100
    assert isinstance(lib.cpm_last_error(), bytes)
    assert lib.cpm_launch_count() == 0                 # nothing has been launched in a CPU-only process


def test_c_abi_has_no_torch_types_and_cites_the_reference():
    text = open(HEADER).read()
    assert "extern \"C\"" in text
    assert "torch::" not in text and "at::Tensor" not in text and "#include <torch" not in text
    for cite in ("ROIAlign.h:57", "ROIAlign.h:98", "ml_nms.h:16", "pet/lib/ops/nms.py", "inference.py:189", "poolers.py"):
        assert cite in text, "header does not cite %s" % cite


def test_argument_errors_reach_python_without_a_gpu():
    """Validation that happens before any CUDA call: status codes + cpm_last_error, re-raised as RuntimeError."""
    lib = _lib.lib()
    pyr = _lib.Pyramid()
    pyr.num_levels = 0
    rc = lib.cpm_roi_align_forward(ctypes.byref(pyr), None, 4, 7, 7, 2, 0, 0, None, None, 0, None, None)
    assert rc < 0 and lib.cpm_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc)
    assert lib.cpm_roi_align_backward_workspace_bytes(1024, 4, 2, 256, 7, 7, 2) > 0
    assert lib.cpm_nms_workspace_bytes(1000) > 0
    assert lib.cpm_nms_batched_workspace_bytes(80000, 80) > 0


def test_no_cpu_fallback():
    """CPU tensors are an error, as for the reference's CUDA-only ops (ml_nms.h:38, ROIAlign_cuda.cu:376)."""
    x = torch.randn(1, 4, 8, 8)
    rois = torch.tensor([[0, 1.0, 1.0, 6.0, 6.0]])
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.roi_align(x, rois, (2, 2), 1.0, 2, False)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.ROIAlign((2, 2), 1.0, 2, False)(x, rois)
    boxes = torch.tensor([[0, 0, 10, 10], [1, 1, 11, 11.0]])
    scores = torch.tensor([0.9, 0.8])
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.nms(boxes, scores, 0.5)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.ml_nms(boxes, scores, torch.zeros(2, dtype=torch.int64), 0.5, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.grid_decode(torch.zeros(1, 9, 28, 28), boxes[:1], ops.calc_sub_regions(9, 3, 56), 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.Pooler("ROIAlign", (7, 7), (0.25, 0.125), 2)([torch.randn(1, 4, 16, 16), torch.randn(1, 4, 8, 8)],
                                                         [ops.BoxList(boxes, (64, 64))])


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "does", "not", "exist.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cpm_r_cnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "cpm_oracle" not in text and "oracle/_ref" not in text, f


def test_reference_signatures_are_kept():
    """pet/lib/ops/roi_align.py:66-86, nms.py:10-11, boxlist_ops.py:15,46, poolers.py:51,103, inference.py:189."""
    assert list(inspect.signature(ops.ROIAlign.__init__).parameters) == [
        "self", "output_size", "spatial_scale", "sampling_ratio", "aligned", "interpolation"]
    assert list(inspect.signature(ops.ROIAlign.forward).parameters) == ["self", "input", "rois"]
    assert list(inspect.signature(ops.Pooler.__init__).parameters) == [
        "self", "method", "output_size", "scales", "sampling_ratio", "rotated", "interpolation"]
    assert list(inspect.signature(ops.Pooler.forward).parameters) == ["self", "x", "boxes"]
    assert list(inspect.signature(ops.nms).parameters)[:3] == ["boxes", "scores", "iou_threshold"]
    assert list(inspect.signature(ops.ml_nms).parameters)[:5] == ["boxes", "scores", "labels", "iou_threshold", "topk"]
    p = inspect.signature(ops.boxlist_nms).parameters
    assert list(p)[:5] == ["boxlist", "nms_thresh", "topk", "score_field", "idxs"] and p["topk"].default == 0
    p = inspect.signature(ops.boxlist_ml_nms).parameters
    assert list(p)[:5] == ["boxlist", "nms_thresh", "topk", "score_field", "label_field"]
    p = inspect.signature(ops.boxlist_nms_legacy).parameters
    assert list(p)[:4] == ["boxlist", "nms_thresh", "max_proposals", "score_field"] and p["max_proposals"].default == -1
    assert list(inspect.signature(ops.GridPostProcessor.get_boxes).parameters)[:4] == [
        "self", "proposals", "grid_pred", "is_train"]
    with pytest.raises(AssertionError):
        ops.Pooler("NoSuchPooler", (7, 7), (0.25,), 2)
    with pytest.raises(NotImplementedError):
        ops.Pooler("ROIPool", (7, 7), (0.25,), 2)
    with pytest.raises(AssertionError):
        ops.ROIAlign((7, 7), 0.25, 2, False, interpolation="cubic")
    with pytest.raises(AssertionError):                                   # roi_align.py:83
        ops.ROIAlign((7, 7), 0.25, 2, False)(torch.zeros(1, 4, 8, 8), torch.zeros(3, 4))


def test_roi_table_and_boxlist_host_logic():
    """Pooler.convert_to_roi_format (poolers.py:90-101) and the BoxList subset the path uses."""
    b0 = ops.BoxList(torch.tensor([[0, 0, 9, 9], [2, 3, 4, 5.0]]), (100, 50))
    b1 = ops.BoxList(torch.zeros(0, 4), (100, 50))
    b2 = ops.BoxList(torch.tensor([[1, 1, 2, 2.0]]), (100, 50))
    pooler = ops.Pooler("ROIAlign", (7, 7), (0.25, 0.125, 0.0625, 0.03125), 2)
    rois = pooler.convert_to_roi_format([b0, b1, b2])
    assert rois.dtype == torch.float32 and rois.shape == (3, 5)
    assert rois[:, 0].tolist() == [0.0, 0.0, 2.0]
    assert (pooler.map_levels.k_min, pooler.map_levels.k_max) == (2.0, 5.0)
    assert pooler.aligned is False and ops.Pooler("ROIAlignV2", (7, 7), (0.25,), 2).aligned is True
    assert b0.area().tolist() == [100.0, 9.0]                             # +1 convention, bounding_box.py:306-310
    xywh = b0.convert("xywh")
    assert xywh.bbox.tolist() == [[0, 0, 10, 10], [2, 3, 3, 3]]
    assert torch.equal(xywh.convert("xyxy").bbox, b0.bbox)
    b0.add_field("scores", torch.tensor([0.1, 0.9]))
    sub = b0[torch.tensor([1])]
    assert len(sub) == 1 and sub.get_field("scores").tolist() == pytest.approx([0.9]) and sub.size == (100, 50)
    with pytest.raises(ValueError):
        ops.BoxList(torch.zeros(3, 5), (10, 10))
    sub_regions = ops.calc_sub_regions(9, 3, 56)                          # grid_rcnn/loss.py:244-273
    assert len(sub_regions) == 9 and all(0 <= v <= 28 for pt in sub_regions for v in pt[:2])


def test_filter_not_gt_matches_the_reference(golden):
    """GridPostProcessor._filter_boxes (inference.py:281-290) as a device-side mask: the rows the reference's own forward
    kept (grid_forward.npz, training case) are exactly the rows the mask keeps."""
    from cpm_r_cnn_b200.grid_decode import filter_not_gt
    g = golden("grid_forward")
    for i in range(2):
        prop, gt = torch.from_numpy(g["train_prop%d" % i]), torch.from_numpy(g["train_gt%d" % i])
        mask = filter_not_gt(prop, gt)
        n_kept = g["train_out_bbox%d" % i].shape[0] - gt.shape[0]
        assert int(mask.sum()) == n_kept
        assert np.array_equal(g["train_out_labels%d" % i][:n_kept], g["train_prop_labels%d" % i][mask.numpy()])
    assert filter_not_gt(torch.zeros(0, 4), torch.zeros(3, 4)).shape == (0,)
    assert filter_not_gt(torch.ones(2, 4), torch.zeros(0, 4)).tolist() == [True, True]


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU RoIAlign on the host cores; no GPU needed): one JSON line with
    the keys the driver reads, `impl: reference`, a cpu_baseline describing the run and an e2e object repeating the value."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "roi_align_fwd_bwd_rois_per_sec" and d["unit"] == "RoIs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
