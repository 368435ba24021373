mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nms or rpn or cls_post or boxlist" > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
timeout 600 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench, cpm_r_cnn_b200 as ops
dev = torch.device("cuda", 0)
class D:  # no dist
    pass
r = bench.bench_nms(ops, dev, 0, 1, None, torch.cuda.synchronize)
for k, v in r.items():
    print(k, "ms %.4f" % v["ms"], "boxes/s %.3g" % v["boxes_per_sec"], "ref same lists", v["reference_gpu"].get("same_keep_lists"), "ref ms %.2f" % v["reference_gpu"].get("ms", -1), "cpu same", v["cpu_baseline"].get("same_keep_set_as_ours"))
PY
