import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench
import cpm_r_cnn_b200 as ops
from cpm_r_cnn_b200 import synthetic as sy
dev = torch.device("cuda", 0)
rois_h, feats_h, gouts_h = bench.make_workload(0)
shapes = [tuple(f.shape) for f in feats_h]
K = rois_h.shape[0]
scales = list(sy.FPN_SCALES)
def pinned(shape, cl=False):
    return torch.empty(shape, pin_memory=True, memory_format=torch.channels_last if cl else torch.contiguous_format)
feats_pin = [pinned(f.shape, True).copy_(f) for f in feats_h]
rois_pin = pinned(rois_h.shape).copy_(rois_h)
gouts_pin = [pinned(g.shape).copy_(g) for g in gouts_h]
outs_pin = [pinned((K, 256, p[0], p[1])) for p in bench.POOLERS]
grads_pin = [pinned(s, True) for s in shapes]
st_in, st_c, st_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
dev_sets = [{"feats": [torch.empty_like(f, device=dev) for f in feats_pin], "rois": torch.empty_like(rois_pin, device=dev),
             "gouts": [torch.empty_like(g, device=dev) for g in gouts_pin]} for _ in range(2)]
ev_in = [torch.cuda.Event() for _ in range(2)]; ev_c = [torch.cuda.Event() for _ in range(2)]
poolers = [ops.Pooler("ROIAlign", p, scales, 2) for p in bench.POOLERS]
mode = sys.argv[1] if len(sys.argv) > 1 else "a"
def run(n, marks=None):
    for i in range(n):
        d = dev_sets[i & 1]
        with torch.cuda.stream(st_in):
            st_in.wait_event(ev_c[i & 1])
            for dst, src in zip(d["feats"] + d["gouts"] + [d["rois"]], feats_pin + gouts_pin + [rois_pin]):
                dst.copy_(src, non_blocking=True)
            ev_in[i & 1].record(st_in)
        with torch.cuda.stream(st_c):
            st_c.wait_event(ev_in[i & 1])
            boxlists = [ops.BoxList(d["rois"][j * 512:(j + 1) * 512, 1:], (sy.IMG_W, sy.IMG_H)) for j in range(2)]
            xs = [f.detach().requires_grad_(True) for f in d["feats"]]
            outs = [pl(xs, boxlists) for pl in poolers]
            torch.autograd.backward(outs, d["gouts"])
            ev_c[i & 1].record(st_c)
        with torch.cuda.stream(st_out):
            st_out.wait_event(ev_c[i & 1])
            for dst, src in zip(outs_pin + grads_pin, [o.detach() for o in outs] + [x.grad for x in xs]):
                dst.copy_(src, non_blocking=True)
                src.record_stream(st_out)
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(st_out); marks.append(e)
        if mode == "b":
            time.sleep(0.002)
    for s_ in (st_in, st_c, st_out): s_.synchronize()
run(6); torch.cuda.synchronize()
for rep in range(3):
    marks = []
    t0 = time.perf_counter(); run(20, marks); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3 / 20
    d = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
    print("wall %.2f ms/step; completion deltas: %s" % (wall, " ".join("%.1f" % v for v in d)))
