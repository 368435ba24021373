"""Grid decode throughput: per-RoI kernel vs streaming kernel (CPM_DECODE_STREAM_MIN picks), R = 16 000 and 4 000."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for mode, env in (("per-RoI", "1000000000"), ("streaming", "1")):
        subprocess.run([sys.executable, __file__, mode], env=dict(os.environ, CPM_DECODE_STREAM_MIN=env))
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch
import cpm_r_cnn_b200 as ops
from cpm_r_cnn_b200 import synthetic as sy
gen = torch.Generator().manual_seed(0)
sub = ops.calc_sub_regions(9, 3, 56)
for R in (1000, 4000, 16000):
    lg = (torch.randn(R, 9, 28, 28, generator=gen) * 2).cuda()
    bx = sy.coco_like_boxes(gen, R).cuda()
    for _ in range(3): ops.grid_decode(lg, bx, sub, 0.5)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): keep = ops.grid_decode(lg, bx, sub, 0.5)
    n = 30
    e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    for i in range(n):
        e[i].record(); g.replay()
    e[n].record(); torch.cuda.synchronize()
    t = sorted(e[i].elapsed_time(e[i + 1]) for i in range(n))[n // 2]
    b = R * (9 * 784 * 4 + 32)
    print("%-9s R=%5d  %.4f ms  %.0f GB/s  frac %.3f" % (sys.argv[1], R, t, b / t / 1e6, b / t / 1e6 / 6540.8))
