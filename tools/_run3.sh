mkdir -p gpurun_out
timeout 200 python tools/bwd_check.py > gpurun_out/bwd_check.log 2>&1; tail -2 gpurun_out/bwd_check.log
CPM_BWD_IMPL=tma timeout 200 python tools/bwd_check.py > gpurun_out/bwd_check_tma.log 2>&1; tail -1 gpurun_out/bwd_check_tma.log
for impl in staged tma; do CPM_BWD_IMPL=$impl CPM_BENCH_SKIP=cl,bf16,refgpu,config0,clocks timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from cpm_r_cnn_b200 import _lib, synthetic as sy
from cpm_r_cnn_b200.roi_align import pooler_backward
dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
shapes = [tuple(f.shape) for f in feats_h]; rois = rois_h.to(dev); gouts = [g.to(dev) for g in gouts_h]; m = _lib.make_mapper(2, 5)
fns = []
for nchw in (False, True):
    for p, go in zip(bench.POOLERS, gouts):
        fns.append(lambda p=p, go=go, nchw=nchw: pooler_backward(go, shapes, list(sy.FPN_SCALES), rois, p, 2, False, 0, m, nchw_grad=nchw))
for f in fns: f(); f()
torch.cuda.synchronize()
caps = [bench.capture(f) for f in fns]
t = bench.time_graphs([c[0] for c in caps], 20)
print(os.environ["CPM_BWD_IMPL"], "bwd7 nhwc %.4f bwd14 nhwc %.4f | bwd7 nchw %.4f bwd14 nchw %.4f" % tuple(t))
PY
done
