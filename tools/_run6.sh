mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "staging or pooler or smoke or bench_workload" > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
CPM_BENCH_SKIP=cl,bf16,refgpu,config0,clocks timeout 300 python - <<'PY'
import os, sys, torch, importlib
sys.path.insert(0, os.getcwd())
import bench
ra = importlib.import_module("cpm_r_cnn_b200.roi_align")
dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
feats = [f.to(dev) for f in feats_h]
fn = lambda: ra.stage_pyramid_nhwc(feats, cache=False)
fn1 = lambda: [ra.stage_nhwc(f, cache=False) for f in feats]
for f in (fn, fn1): f(); f()
torch.cuda.synchronize()
caps = [bench.capture(f) for f in (fn, fn1)]
t = bench.time_graphs([c[0] for c in caps], 20)
nb = 2 * sum(f.numel() * 4 for f in feats)
print("fused staging %.4f ms (%.0f GB/s) | per-level %.4f ms (%.0f GB/s)" % (t[0], nb / t[0] / 1e6, t[1], nb / t[1] / 1e6))
PY
