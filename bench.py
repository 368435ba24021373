#!/usr/bin/env python
"""Benchmark of the CPM R-CNN detection-head op path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): "R-50-FPN CPM R-CNN head: 7x7 box + 14x14 grid-point ROIAlign fwd/bwd, 2 img/GPU,
512 RoIs/img" -- one step = the 7x7 cls-head Pooler and the 14x14 grid-head Pooler, each forward + backward, over the
4-level 256-channel fp32 FPN pyramid of two 800x1344 images and 1024 COCO-shaped RoIs.  One unit of work = one RoI
through one pooler forward+backward, so a step is 2048 units.  `value` = units/s with everything resident in HBM;
`e2e` = the same step through the Python op layer with HOST (pinned) inputs and outputs, copies inside the timed
region.  The NMS half of the metric (configs[2]) is reported in the `nms` object of the same line.

--impl reference times the reference's own CPU RoIAlign (oracle/_ref/pet_ref_cpu.so = unmodified
/root/reference/.../ROIAlign_cpu.cpp, else the C port in oracle/) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of each op's main kernel, from the committed `ncu --set full`
# capture of this same workload (profiles/ncu_r01_summary.txt); not re-measured by bench.py (a run under ncu is never timed)
NCU_DRAM_BYTES = {"fwd7": 157.1e6, "fwd14": 316.0e6, "bwd7": 177.5e6, "bwd14": 361.2e6}
NCU_DRAM_SOURCE = "profiles/ncu_r01_summary.txt (ncu --set full, one capture per kernel)"

METRIC = "roi_align_fwd_bwd_rois_per_sec"
UNIT = "RoIs/s"
IMGS_PER_GPU, ROIS_PER_IMG, CHANNELS = 2, 512, 256
POOLERS = ((7, 7), (14, 14))
SAMPLING = 2


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_workload(rank, device=None):
    from cpm_r_cnn_b200 import synthetic as sy
    gen = torch.Generator().manual_seed(rank)
    rois = sy.coco_like_rois(gen, ROIS_PER_IMG, IMGS_PER_GPU)
    feats = sy.pyramid(gen, IMGS_PER_GPU, CHANNELS)
    gouts = [torch.randn(rois.shape[0], CHANNELS, p[0], p[1], generator=gen) for p in POOLERS]
    return rois, feats, gouts


def algorithmic_bytes(rois):
    """fwd = K*C*PH*PW*4 + U*C*4 + 20K ; bwd = K*C*PH*PW*4 + M*C*4 + 20K (fp32)."""
    from cpm_r_cnn_b200 import synthetic as sy
    shapes = sy.level_shapes()
    lv = sy.fpn_levels_host(rois)
    K = rois.shape[0]
    M = IMGS_PER_GPU * sum(h * w for h, w in shapes)
    out = {}
    for p in POOLERS:
        U = sy.touched_pixels(rois, lv, shapes, sy.FPN_SCALES, p, SAMPLING)
        pooled = K * CHANNELS * p[0] * p[1] * 4
        out["fwd%d" % p[0]] = pooled + U * CHANNELS * 4 + 20 * K
        out["bwd%d" % p[0]] = pooled + M * CHANNELS * 4 + 20 * K
        out["U%d" % p[0]] = U
    return out


# ---------------------------------------------------------------------------------------------------------------------
# reference / CPU arm
# ---------------------------------------------------------------------------------------------------------------------
_WORKLOAD_CACHE = {}


def cpu_reference_rate(sample_rois_per_img=48):
    """The reference's CPU RoIAlign through the reference Pooler's per-level loop (poolers.py:127-130), forward +
    backward at 7x7 and 14x14, on the first `sample_rois_per_img` RoIs of each image of the rank-0 workload."""
    from cpm_r_cnn_b200 import synthetic as sy
    torch.set_num_threads(1)      # the kernel is single-threaded by construction (ROIAlign_cpu.cpp:185-186: omp commented out)
    if 0 not in _WORKLOAD_CACHE:
        _WORKLOAD_CACHE[0] = make_workload(0)
    rois, feats, gouts = _WORKLOAD_CACHE[0]
    sel = torch.cat([torch.arange(i * ROIS_PER_IMG, i * ROIS_PER_IMG + sample_rois_per_img) for i in range(IMGS_PER_GPU)])
    rois = rois[sel]
    gouts = [g[sel] for g in gouts]
    lv = sy.fpn_levels_host(rois)
    kind, fwd, bwd = "port", None, None
    try:
        from oracle import build_ref
        ref = build_ref.load("pet_ref_cpu")
        kind = "reference"
        fwd = lambda f, r, s, p: ref.roi_align_forward(f, r, s, p[0], p[1], SAMPLING, False, 0)
        bwd = lambda g, r, s, p, shp: ref.roi_align_backward(g, r, s, p[0], p[1], shp[0], shp[1], shp[2], shp[3],
                                                             SAMPLING, False, 0)
    except Exception:
        import oracle
        fwd = lambda f, r, s, p: torch.from_numpy(oracle.roi_align_forward(f.numpy(), r.numpy(), s, p[0], p[1], SAMPLING, False))
        bwd = lambda g, r, s, p, shp: torch.from_numpy(oracle.roi_align_backward(g.numpy(), r.numpy(), s, p[0], p[1], shp[0],
                                                                                 shp[1], shp[2], shp[3], SAMPLING, False))
    t0 = time.perf_counter()
    units = 0
    for p, go in zip(POOLERS, gouts):
        for l, f in enumerate(feats):
            idx = torch.nonzero(lv == l).squeeze(1)
            r = rois[idx].contiguous()
            o = fwd(f, r, sy.FPN_SCALES[l], p)
            g = bwd(go[idx].contiguous(), r, sy.FPN_SCALES[l], p, tuple(f.shape))
            assert o.shape[0] == r.shape[0] and g.shape == f.shape
        units += rois.shape[0]
    dt = time.perf_counter() - t0
    sample = ("%d of %d RoIs/img x %d img, 7x7 + 14x14 fwd+bwd through the per-level Pooler loop, fp32, 4 levels x %d ch"
              % (sample_rois_per_img, ROIS_PER_IMG, IMGS_PER_GPU, CHANNELS))
    return units / dt, dt, kind, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates, times = [], []
    # each step is a bounded sample sized so that the whole run stays within ~2 minutes of single-thread CPU work
    n_step = min(ROIS_PER_IMG, max(8, 3000 // max(args.steps, 1)))
    for i in range(args.warmup + args.steps):
        n = 8 if i < args.warmup else n_step
        r, dt, kind, sample = cpu_reference_rate(n)
        if i >= args.warmup:
            rates.append(r)
            times.append(dt)
    value = sum(rates) / len(rates)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: R-50-FPN CPM head RoIAlign 7x7 + 14x14 fwd+bwd, 2 img, 512 RoIs/img, "
                                   "256 ch fp32 (bounded sample per step)", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import cpm_r_cnn_b200 as ops
    from cpm_r_cnn_b200 import _lib, sharding, synthetic as sy
    from cpm_r_cnn_b200.roi_align import pooler_backward, pooler_forward

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the cpm_ops path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly one line, the JSON: everything libraries print while the job runs (NCCL writes its version
    # banner to stdout when the first communicator is created) is routed to stderr, and the real stdout is restored for
    # the result
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    rois_h, feats_h, gouts_h = make_workload(rank)
    shapes = [tuple(f.shape) for f in feats_h]
    scales = list(sy.FPN_SCALES)
    mapper = _lib.make_mapper(2, 5)
    K = rois_h.shape[0]
    # device-resident inputs: channels_last pyramid (the layout a channels_last backbone/FPN emits; zero-copy NHWC)
    feats = [f.to(dev).contiguous(memory_format=torch.channels_last) for f in feats_h]
    rois = rois_h.to(dev)
    gouts = [g.to(dev) for g in gouts_h]

    # The step is captured once into four CUDA graphs (one per op, through the Python op layer) and replayed: the ops are
    # tens of microseconds each, so an eager Python loop would time the host's enqueue rate, not the kernels.
    def op_fwd(p):
        return lambda: pooler_forward(feats, scales, rois, p, SAMPLING, False, 0, mapper)

    def op_bwd(p, go):
        return lambda: pooler_backward(go, shapes, scales, rois, p, SAMPLING, False, 0, mapper)

    op_fns = []
    for p, go in zip(POOLERS, gouts):
        op_fns += [op_fwd(p), op_bwd(p, go)]
    names = ["fwd7", "bwd7", "fwd14", "bwd14"]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            for fn in op_fns:
                fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    sync_all()
    graphs, keep, launches_per_step = [], [], 0
    for fn in op_fns:
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(g):
            keep.append(fn())
        launches_per_step += _lib.launch_count() - l0
        graphs.append(g)

    def step(evs=None):
        for j, g in enumerate(graphs):
            if evs: evs[j].record()
            g.replay()
        if evs: evs[len(graphs)].record()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0 and "clocks" not in os.environ.get("CPM_BENCH_SKIP", ""):
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    sync_all()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_beg.record()
    for i in range(args.steps):
        step(evs[i])
    t_end.record()
    sync_all()
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    # whole-job figure: units of all ranks / the slowest rank's device time (cpm_r_cnn_b200/sharding.py)
    units_total, sec_total, value = sharding.aggregate(K * len(POOLERS) * args.steps, t_beg.elapsed_time(t_end) * 1e-3)
    ms_step = sec_total * 1e3 / args.steps
    units_per_step = K * len(POOLERS) * world
    op_ms = {n: sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.steps for i, n in enumerate(names)}

    # ---- the same four ops with channels_last pooled tensors (extension, not part of `value`): the forward returns the
    #      pooled block with channels_last strides and the backward reads a channels_last gradient in place ----
    cl_ms = {}
    try:
        if "cl" in os.environ.get("CPM_BENCH_SKIP", ""):
            raise RuntimeError("skipped")
        gouts_cl = [g.contiguous(memory_format=torch.channels_last) for g in gouts]
        cl_fns = []
        for p, go in zip(POOLERS, gouts_cl):
            cl_fns += [(lambda p=p: pooler_forward(feats, scales, rois, p, SAMPLING, False, 0, mapper, channels_last=True)),
                       op_bwd(p, go)]
        with torch.cuda.stream(side):
            for fn in cl_fns:
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        sync_all()
        cl_graphs, cl_keep = [], []
        for fn in cl_fns:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                cl_keep.append(fn())
            cl_graphs.append(g)
        n_cl = max(3, min(args.steps, 20))
        acc = [0.0] * 4
        for it in range(n_cl + 2):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            for j, g in enumerate(cl_graphs):
                e[j].record()
                g.replay()
            e[4].record()
            torch.cuda.synchronize()
            if it >= 2:
                for j in range(4):
                    acc[j] += e[j].elapsed_time(e[j + 1])
        cl_ms = {n: acc[j] / n_cl for j, n in enumerate(names)}
        del cl_graphs, cl_keep, gouts_cl
    except Exception as ex:      # reported, never fatal for the headline
        cl_ms = {"error": repr(ex)[:200]}
    sync_all()

    # ---- forward on a bf16 pyramid (bf16 pooled output, fp32 arithmetic): reported separately, own algorithmic bytes
    #      (2 bytes per stored element); the tolerance of this path is stated in tests/test_gpu_parity.py ----
    bf16_ms = {}
    try:
        if "bf16" in os.environ.get("CPM_BENCH_SKIP", ""):
            raise RuntimeError("skipped")
        feats16 = [f.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in feats]
        b_fns = [(lambda p=p: pooler_forward(feats16, scales, rois, p, SAMPLING, False, 0, mapper)) for p in POOLERS]
        with torch.cuda.stream(side):
            for fn in b_fns:
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        sync_all()
        b_graphs, b_keep = [], []
        for fn in b_fns:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                b_keep.append(fn())
            b_graphs.append(g)
        n_b = max(3, min(args.steps, 20))
        acc = [0.0] * len(b_graphs)
        for it in range(n_b + 2):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(len(b_graphs) + 1)]
            for j, g in enumerate(b_graphs):
                e[j].record()
                g.replay()
            e[len(b_graphs)].record()
            torch.cuda.synchronize()
            if it >= 2:
                for j in range(len(b_graphs)):
                    acc[j] += e[j].elapsed_time(e[j + 1])
        bf16_ms = {"fwd%d" % p[0]: acc[j] / n_b for j, p in enumerate(POOLERS)}
        del b_graphs, b_keep, feats16
    except Exception as ex:
        bf16_ms = {"error": repr(ex)[:200]}
    sync_all()

    # ---- the red.global.add fallback of the backward (non-deterministic order), reported separately ----
    atomic_ms = {}
    try:
        at_fns = [(lambda p=p, go=go: pooler_backward(go, shapes, scales, rois, p, SAMPLING, False, 0, mapper, mode="atomic"))
                  for p, go in zip(POOLERS, gouts)]
        with torch.cuda.stream(side):
            for fn in at_fns:
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        sync_all()
        at_graphs, at_keep = [], []
        for fn in at_fns:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                at_keep.append(fn())
            at_graphs.append(g)
        acc = [0.0, 0.0]
        for it in range(7):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            for j, g in enumerate(at_graphs):
                e[j].record()
                g.replay()
            e[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                acc[0] += e[0].elapsed_time(e[1])
                acc[1] += e[1].elapsed_time(e[2])
        atomic_ms = {"bwd7": acc[0] / 5, "bwd14": acc[1] / 5}
        del at_graphs, at_keep
    except Exception as ex:
        atomic_ms = {"error": repr(ex)[:200]}
    sync_all()

    # ---- e2e: host (pinned) buffers in and out, copies inside the timed region ----
    # One step = H2D of the pyramid, the RoIs and both pooled gradients, the two Pooler modules forward + one autograd
    # backward (the feature gradient of both poolers accumulates into x.grad, as in the head), D2H of both pooled outputs
    # and the gradient pyramid.  Three streams (H2D / compute / D2H) and two device input sets let step i+1's upload run
    # under step i's compute and download (PCIe is full duplex); every step still moves all of its bytes.
    def pinned(shape, channels_last=False):
        return torch.empty(shape, pin_memory=True, memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    feats_pin = [pinned(f.shape, True).copy_(f) for f in feats_h]
    rois_pin = pinned(rois_h.shape).copy_(rois_h)
    gouts_pin = [pinned(g.shape).copy_(g) for g in gouts_h]
    outs_pin = [pinned((K, CHANNELS, p[0], p[1])) for p in POOLERS]
    grads_pin = [pinned(s, True) for s in shapes]
    assert all(t.is_pinned() for t in feats_pin + gouts_pin + outs_pin + grads_pin + [rois_pin])
    h2d = sum(t.numel() * 4 for t in feats_pin + gouts_pin) + rois_pin.numel() * 4
    d2h = sum(t.numel() * 4 for t in outs_pin + grads_pin)
    st_in, st_c, st_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dev_sets = [{"feats": [torch.empty_like(f, device=dev) for f in feats_pin], "rois": torch.empty_like(rois_pin, device=dev),
                 "gouts": [torch.empty_like(g, device=dev) for g in gouts_pin]} for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_c = [torch.cuda.Event() for _ in range(2)]
    poolers = [ops.Pooler("ROIAlign", p, scales, SAMPLING) for p in POOLERS]

    ev_out = [torch.cuda.Event() for _ in range(2)]
    hold = [None, None]       # a step's results stay referenced until the compute stream has waited for their download:
                              # their memory then returns to the allocator in stream order (no record_stream, whose
                              # deferred frees make the caching allocator fall back to cudaMalloc at unpredictable times)

    def e2e_steps_run(n):
        for i in range(n):
            d = dev_sets[i & 1]
            with torch.cuda.stream(st_in):
                st_in.wait_event(ev_c[i & 1])          # the compute that last read this input set has finished
                for dst, src in zip(d["feats"] + d["gouts"] + [d["rois"]], feats_pin + gouts_pin + [rois_pin]):
                    dst.copy_(src, non_blocking=True)
                ev_in[i & 1].record(st_in)
            with torch.cuda.stream(st_c):
                st_c.wait_event(ev_in[i & 1])
                st_c.wait_event(ev_out[i & 1])         # the download of step i - 2 has read its tensors ...
                hold[i & 1] = None                     # ... so they can be recycled by this step
                boxlists = [ops.BoxList(d["rois"][j * ROIS_PER_IMG:(j + 1) * ROIS_PER_IMG, 1:], (sy.IMG_W, sy.IMG_H))
                            for j in range(IMGS_PER_GPU)]
                xs = [f.detach().requires_grad_(True) for f in d["feats"]]
                outs = [pl(xs, boxlists) for pl in poolers]
                torch.autograd.backward(outs, d["gouts"])
                ev_c[i & 1].record(st_c)
            with torch.cuda.stream(st_out):
                st_out.wait_event(ev_c[i & 1])
                res = [o.detach() for o in outs] + [x.grad for x in xs]
                for dst, src in zip(outs_pin + grads_pin, res):
                    dst.copy_(src, non_blocking=True)
                ev_out[i & 1].record(st_out)
            hold[i & 1] = (outs, xs, res, boxlists)
            del outs, xs, res, boxlists
        for s_ in (st_in, st_c, st_out):
            s_.synchronize()

    e2e_steps = max(10, min(args.steps, 50))
    e2e_steps_run(6)          # untimed: lets torch's caching allocator reach its steady state on all three streams
    sync_all()
    t0 = time.perf_counter()
    e2e_steps_run(e2e_steps)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = units_per_step / (e2e_ms * 1e-3)

    # ---- NMS half of the metric (configs[2]) ----
    nms = bench_nms(ops, dev, rank, world, dist, sync_all)
    decode = bench_decode(ops, dev, rank)
    rpn = bench_rpn(ops, dev, rank)
    det = bench_detect(ops, dev, rank)
    gtg = bench_grid_targets(ops, dev, rank)
    mat = bench_matcher(ops, dev, rank)

    if rank == 0:
        peak, peak_src = measured_peak()
        ab = algorithmic_bytes(rois_h)
        rl_ops = {n: {"ms": op_ms[n], "bytes": ab[n], "gbs": ab[n] / (op_ms[n] * 1e-3) / 1e9,
                      "frac": ab[n] / (op_ms[n] * 1e-3) / 1e9 / peak, "frac_of_nominal_8TBs": ab[n] / (op_ms[n] * 1e-3) / 1e9 / 8000.0}
                  for n in names}
        top = max(names, key=lambda n: op_ms[n])
        total_bytes = sum(ab[n] for n in names)
        # CPU baseline beside it: the full workload of the step (512 RoIs/img x 2 img), 4 passes (~10 s of CPU work)
        cpu_rate = cpu_dt = cpu_kind = cpu_sample = None
        if world == 1:
            runs = [cpu_reference_rate(ROIS_PER_IMG) for _ in range(4)]
            cpu_dt = sum(r[1] for r in runs)
            cpu_rate = sum(r[0] * r[1] for r in runs) / cpu_dt          # units / total seconds
            cpu_kind, cpu_sample = runs[0][2], runs[0][3] + ", 4 passes"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: R-50-FPN CPM head RoIAlign 7x7 + 14x14 fwd+bwd, %d img/GPU, %d RoIs/img, "
                                       "4-level 256-ch fp32 pyramid of 800x1344 images, sampling_ratio 2" % (IMGS_PER_GPU, ROIS_PER_IMG),
                           "layout": "pyramid channels_last (NHWC, zero-copy), pooled output (K,C,PH,PW) contiguous",
                           "backward": "deterministic tile-owner gather (no atomics)",
                           "unit_of_work": "one RoI through one pooler forward+backward; %d per step per GPU" % (K * len(POOLERS)),
                           "launch": "each op captured once in a CUDA graph through the Python op layer and replayed; per-op time = CUDA "
                                     "events between consecutive graph replays on the launching stream",
                           "l2": "not flushed: per-step working set (pyramid 183 MB + pooled/grad_out 514 MB + gradients 366 MB) "
                                 "exceeds the 126 MB L2",
                           "parallelism": "dp%d (images sharded per GPU, no collective inside the ops)" % world},
                "roofline": {"bound": "hbm", "kernel": {"fwd7": "roi_align_fwd_cols<1> (7x7)", "fwd14": "roi_align_fwd_cols<2> (14x14)",
                                                         "bwd7": "bwd_tiles_staged<0,7,2> (7x7) + bwd_prepare",
                                                         "bwd14": "bwd_tiles_staged<0,14,2> (14x14) + bwd_prepare"}[top],
                             "achieved": rl_ops[top]["gbs"], "peak": peak, "unit": "GB/s", "frac": rl_ops[top]["frac"],
                             "traffic": NCU_DRAM_BYTES.get(top), "traffic_source": NCU_DRAM_SOURCE, "peak_source": peak_src,
                             "step": {"bytes": total_bytes, "gbs": total_bytes / (ms_step * 1e-3) / 1e9,
                                      "frac": total_bytes / (ms_step * 1e-3) / 1e9 / peak},
                             "ops": rl_ops, "U_px": {"7x7": ab["U7"], "14x14": ab["U14"]},
                             "backward_atomic_fallback_ms": atomic_ms,
                             "fwd_bf16_storage": ({n: {"ms": bf16_ms[n], "bytes": (ab[n] - 20 * K) // 2 + 20 * K,
                                                       "gbs": ((ab[n] - 20 * K) // 2 + 20 * K) / (bf16_ms[n] * 1e-3) / 1e9,
                                                       "frac": ((ab[n] - 20 * K) // 2 + 20 * K) / (bf16_ms[n] * 1e-3) / 1e9 / peak}
                                                   for n in bf16_ms} if "error" not in bf16_ms else bf16_ms),
                             "ops_channels_last_pooled": ({n: {"ms": cl_ms[n], "gbs": ab[n] / (cl_ms[n] * 1e-3) / 1e9,
                                                               "frac": ab[n] / (cl_ms[n] * 1e-3) / 1e9 / peak} for n in names}
                                                          if "error" not in cl_ms else cl_ms)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms, "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": clocks, "nms": nms, "grid_decode": decode, "rpn_proposals": rpn, "detection_postprocess": det, "grid_targets": gtg, "iou_matcher": mat}
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": cpu_kind, "sample": cpu_sample,
                                    "seconds": cpu_dt}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_nms(ops, dev, rank, world, dist, sync_all, iters=10):
    """configs[2]: batched NMS, 16 images: RPN flavour (5 levels x 1000 proposals, thr 0.7) and detection flavour
    (1000 proposals x 80 classes, score > 0.03 gate and un-gated stress, thr 0.3).  boxes/s = input boxes / time."""
    from cpm_r_cnn_b200 import synthetic as sy
    gen = torch.Generator().manual_seed(1000 + rank)
    res = {}
    cases = {}
    b, s, seg = sy.rpn_like_candidates(gen, 16, 5, 1000)
    cases["rpn_16img_x5lvl_x1000_thr0.7"] = (b, s, seg, 80, 0.7)
    b, s, seg, lab, img = sy.detection_candidates(gen, 16, 1000, 80, 0.03)
    cases["det_16img_x80cls_gated0.03_thr0.3"] = (b, s, seg, 16 * 80, 0.3)
    b, s, seg, lab, img = sy.detection_candidates(gen, 16, 1000, 80, -1.0)
    cases["det_16img_x80cls_x1000_ungated_thr0.3"] = (b, s, seg, 16 * 80, 0.3)
    for name, (b, s, seg, nseg, thr) in cases.items():
        b, s, seg = b.to(dev), s.to(dev), seg.to(dev)
        for _ in range(3):
            ops.batched_nms(b, s, seg, nseg, thr, sync=False)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            keep, counts, total = ops.batched_nms(b, s, seg, nseg, thr, sync=False)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        res[name] = {"boxes": int(b.shape[0]) * world, "segments": nseg * world, "ms": ms, "kept": int(total.item()),
                     "boxes_per_sec": b.shape[0] * world / (ms * 1e-3)}
        if rank == 0:
            res[name]["reference_gpu"] = reference_gpu_nms(name, b, s, seg, nseg, thr, int(total.item()))
            if res[name]["reference_gpu"].get("ms"):
                res[name]["speedup_vs_reference_gpu"] = res[name]["reference_gpu"]["ms"] * world / ms if world == 1 else None
    return res


def bench_decode(ops, dev, rank, iters=20):
    """Grid-point decode (GridPostProcessor.get_boxes): R RoIs x 9 points x 28x28 logits, one stage; streaming read of
    R*9*28*28*4 + 32R bytes.  R = 1000 (one test image) and 16000 (16 images)."""
    from cpm_r_cnn_b200 import synthetic as sy
    peak, _ = measured_peak()
    res = {}
    if rank != 0:
        return res
    gen = torch.Generator().manual_seed(77)
    sub = ops.calc_sub_regions(9, 3, 56)
    for R in (1000, 16000):
        logits = (torch.randn(R, 9, 28, 28, generator=gen) * 2).to(dev)
        boxes = sy.coco_like_boxes(gen, R).to(dev)
        for _ in range(3):
            ops.grid_decode(logits, boxes, sub, 0.5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.grid_decode(logits, boxes, sub, 0.5)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        nbytes = R * 9 * 28 * 28 * 4 + 32 * R
        res["R%d" % R] = {"ms": ms, "rois_per_sec": R / (ms * 1e-3), "bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9,
                          "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
    return res


def bench_rpn(ops, dev, rank, iters=10):
    """Next row: RPN proposal selection, 2 images x 5 FPN levels x 3 anchors of an 800x1344 image,
    PRE_NMS_TOP_N_TRAIN 2000 / POST 2000 / FPN_POST_NMS_TOP_N_TRAIN 2000, thr 0.7 (training flavour, top-k over the
    batch).  Ours = RPNPostProcessor (one decode launch + one batched NMS); reference_gpu = the reference's own loop
    structure (inference.py:96-113: per level, per image clip -> remove_small_boxes -> torchvision nms) restated with the
    same torch ops on the same GPU.  proposals/s = candidates entering the selection / time."""
    import torchvision
    res = {}
    if rank != 0:
        return res
    gen = torch.Generator().manual_seed(5)
    N, A, img = 2, 3, (1344, 800)
    strides = (4, 8, 16, 32, 64)
    shapes = [((800 + s - 1) // s, (1344 + s - 1) // s) for s in strides]
    anchors, obj, reg = [], [], []
    for (h, w), st in zip(shapes, strides):
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        ctr = torch.stack([xs, ys], -1).reshape(-1, 1, 2) * st + st / 2
        half = torch.tensor([[5.6 * st, 2.8 * st], [4.0 * st, 4.0 * st], [2.8 * st, 5.6 * st]]).reshape(1, A, 2)
        anchors.append(torch.cat([ctr - half, ctr + half - 1], -1).reshape(-1, 4).to(dev))
        obj.append((torch.randn(N, A, h, w, generator=gen) * 2).to(dev))
        reg.append((torch.randn(N, 4 * A, h, w, generator=gen) * 0.3).to(dev))
    pre, post, thr, min_size, fpn_post = 2000, 2000, 0.7, 0, 2000
    pp = ops.RPNPostProcessor(pre, post, thr, min_size, None, fpn_post, True).train()
    alist = [[ops.BoxList(a, img) for a in anchors] for _ in range(N)]

    def ours():
        return pp(alist, obj, reg)

    def reference_loop():
        out = [[] for _ in range(N)]
        for a, o, r in zip(anchors, obj, reg):
            s, d, an = pp._level_candidates([ops.BoxList(a, img)] * N, o, r)
            w_, h_ = (an[..., 2] - an[..., 0]) + 1, (an[..., 3] - an[..., 1]) + 1
            cx, cy = an[..., 0] + 0.5 * w_, an[..., 1] + 0.5 * h_
            dw, dh = d[..., 2].clamp(max=pp.box_coder.bbox_xform_clip), d[..., 3].clamp(max=pp.box_coder.bbox_xform_clip)
            pcx, pcy, pw_, ph_ = d[..., 0] * w_ + cx, d[..., 1] * h_ + cy, torch.exp(dw) * w_, torch.exp(dh) * h_
            boxes = torch.stack([pcx - 0.5 * pw_, pcy - 0.5 * ph_, pcx + 0.5 * pw_ - 1, pcy + 0.5 * ph_ - 1], -1)
            for i in range(N):
                b = boxes[i].clone()
                b[:, 0::2].clamp_(min=0, max=img[0] - 1)
                b[:, 1::2].clamp_(min=0, max=img[1] - 1)
                keep = ((b[:, 2] - b[:, 0] + 1 >= min_size) & (b[:, 3] - b[:, 1] + 1 >= min_size)).nonzero().squeeze(1)
                b, sc = b[keep], s[i][keep]
                k = torchvision.ops.nms(b, sc, thr)[:post]
                out[i].append((b[k], sc[k]))
        allsc = torch.cat([sc for lv in out for _, sc in lv])
        return torch.topk(allsc, min(fpn_post, allsc.numel()))[0]

    cand = sum(min(pre, A * h * w) for (h, w) in shapes) * N
    for name, fn in (("ours", ours), ("reference_gpu_loop", reference_loop)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / iters       # wall clock on purpose: both arms are host-sync bound
        res[name] = {"ms": ms, "candidates": cand, "proposals_per_sec": cand / (ms * 1e-3)}
    res["kept_per_image"] = [len(b) for b in ours()]
    res["speedup"] = res["reference_gpu_loop"]["ms"] / res["ours"]["ms"]
    return res


def bench_detect(ops, dev, rank, iters=5):
    """Next row: detection post-processing, 16 images x 1000 proposals x 81 classes, score > 0.03,
    thr 0.3.  Ours = CLSPostProcessor (whole batch, one NMS launch); reference_gpu = the reference's per-image loop
    (inference.py:91-124: repeat, clip, numpy label build + H2D, mask, `_C.ml_nms` = the unmodified ml_nms.cu when
    oracle/_ref is present, else torchvision.ops.batched_nms) with the same torch ops on the same GPU."""
    import numpy as np
    import torchvision
    from cpm_r_cnn_b200 import synthetic as sy
    res = {}
    if rank != 0:
        return res
    gen = torch.Generator().manual_seed(21)
    B, R, C, img = 16, 1000, 81, (sy.IMG_W, sy.IMG_H)
    boxes = [sy.coco_like_boxes(gen, R).to(dev) for _ in range(B)]
    logits = torch.randn(B * R, C, generator=gen).to(dev)
    pp = ops.CLSPostProcessor(0.03, 0.3)
    blists = [ops.BoxList(b, img) for b in boxes]
    try:
        from oracle import build_ref
        refk = build_ref.load("pet_ref_cuda")
        ml = lambda b, s, l: refk.ml_nms(b, s, l, 0.3, 0)
        op = "_C.ml_nms (reference ml_nms.cu, unmodified) per image"
    except Exception:
        ml = lambda b, s, l: torchvision.ops.batched_nms(b, s, l, 0.3)
        op = "torchvision.ops.batched_nms per image (reference build unavailable)"

    def ours():
        return pp(logits, blists)

    def reference_loop():
        prob = torch.softmax(logits, -1)
        out = []
        for i in range(B):
            p = prob[i * R:(i + 1) * R]
            b = boxes[i].repeat(1, C).reshape(-1, 4)
            b[:, 0::2].clamp_(min=0, max=img[0] - 1)
            b[:, 1::2].clamp_(min=0, max=img[1] - 1)
            sc = p.reshape(-1)
            labels = torch.from_numpy(np.tile(np.arange(C), R)).to(dtype=torch.int64, device=dev)
            fg = torch.from_numpy((np.arange(R * C) % C != 0).astype(int)).to(dtype=torch.bool, device=dev)
            m = (sc > 0.03) & fg
            keep = ml(b[m], sc[m], labels[m])
            out.append(keep.numel())
        return out

    for name, fn in (("ours", ours), ("reference_gpu_loop", reference_loop)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            r = fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / iters
        res[name] = {"ms": ms, "proposals_x_classes": B * R * (C - 1), "pairs_per_sec": B * R * (C - 1) / (ms * 1e-3)}
    res["reference_gpu_loop"]["op"] = op
    res["detections"] = sum(len(b) for b in ours())
    res["same_count_as_reference"] = res["detections"] == sum(reference_loop())
    res["speedup"] = res["reference_gpu_loop"]["ms"] / res["ours"]["ms"]
    return res


def bench_grid_targets(ops, dev, rank, iters=20):
    """Next row: grid-point training targets for 2 x 96 positives (MAX_SAMPLE_NUM_GRID), stage 0.
    Ours = one kernel writing (R,9,28,28) on the device (bytes = R*9*28*28*4 + 32R); cpu_baseline = the reference's
    algorithm on the host (oracle port of the Python triple loop, loss.py:214-237) plus the upload of its result, which
    is what the reference does every iteration and stage (:256-257)."""
    from cpm_r_cnn_b200 import synthetic as sy
    res = {}
    if rank != 0:
        return res
    gen = torch.Generator().manual_seed(3)
    R = 192
    pos = sy.coco_like_boxes(gen, R)
    gt = pos + (torch.rand(R, 4, generator=gen) - 0.5) * (pos[:, 2:] - pos[:, :2]).repeat(1, 2) * 0.6
    pd, gd = pos.to(dev), gt.to(dev)
    for _ in range(3):
        ops.prepare_grid_target(pd, gd, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        t = ops.prepare_grid_target(pd, gd, 1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = R * 9 * 28 * 28 * 4 + 32 * R
    res["ours"] = {"ms": ms, "rois_per_sec": R / (ms * 1e-3), "bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9}
    try:
        import oracle
        from oracle import grid_targets as ogt
        sub = oracle.calc_sub_regions(9, 3, 56)
        t0 = time.perf_counter()
        ref = ogt.prepare_target(pos.numpy(), gt.numpy(), 1.0, sub)
        up = torch.from_numpy(ref).to(dev)
        torch.cuda.synchronize()
        cms = (time.perf_counter() - t0) * 1e3
        res["cpu_baseline"] = {"ms": cms, "rois_per_sec": R / (cms * 1e-3), "kind": "port", "cores": 1,
                               "identical": bool(torch.equal(up, t))}
        res["speedup"] = cms / ms
    except Exception as ex:
        res["cpu_baseline"] = {"unavailable": repr(ex)[:200]}
    return res


def bench_matcher(ops, dev, rank, iters=20):
    """Next row: IoU matrix + Matcher for one image, 50 ground-truth boxes x 2000 proposals (RPN
    thresholds, low-quality matches on).  Ours = cpm_box_iou + cpm_matcher (3 launches, no host sync); reference_gpu = the
    reference's torch expressions (boxlist_ops.py:123-158, matcher.py:52-112) on the same GPU."""
    from cpm_r_cnn_b200 import synthetic as sy
    res = {}
    if rank != 0:
        return res
    gen = torch.Generator().manual_seed(12)
    M, N, img = 50, 2000, (sy.IMG_W, sy.IMG_H)
    gt, pr = sy.coco_like_boxes(gen, M).to(dev), sy.coco_like_boxes(gen, N).to(dev)
    g, p = ops.BoxList(gt, img), ops.BoxList(pr, img)
    matcher = ops.Matcher(0.7, 0.3, True)

    def ours():
        return matcher(ops.boxlist_iou(g, p))

    def reference():
        a1 = (gt[:, 2] - gt[:, 0] + 1) * (gt[:, 3] - gt[:, 1] + 1)
        a2 = (pr[:, 2] - pr[:, 0] + 1) * (pr[:, 3] - pr[:, 1] + 1)
        lt, rb = torch.max(gt[:, None, :2], pr[:, :2]), torch.min(gt[:, None, 2:], pr[:, 2:])
        wh = (rb - lt + 1).clamp(min=0)
        inter = wh[:, :, 0] * wh[:, :, 1]
        q = inter / (a1[:, None] + a2 - inter)
        vals, m = q.max(dim=0)
        allm = m.clone()
        m[vals < 0.3] = -1
        m[(vals >= 0.3) & (vals < 0.7)] = -2
        best, _ = q.max(dim=1)
        upd = torch.nonzero(q == best[:, None])[:, 1]
        m[upd] = allm[upd]
        return m

    same = bool(torch.equal(ours(), reference()))
    for name, fn in (("ours", ours), ("reference_gpu", reference)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / iters
        res[name] = {"ms": ms, "pairs_per_sec": M * N / (ms * 1e-3)}
    res["identical"] = same
    res["speedup"] = res["reference_gpu"]["ms"] / res["ours"]["ms"]
    return res


def reference_gpu_nms(name, b, s, seg, nseg, thr, kept_ours):
    """The reference's own single-GPU op path on the same boxes, timed with CUDA events in this process:
    RPN flavour  = one pet.lib.ops.nms (= torchvision.ops.nms, pet/lib/ops/nms.py:2,10) call per (image, level), the loop
                   of rpn/inference.py:102-113;
    detection    = one _C.ml_nms call per image (grid_cascade_rcnn/inference.py:91-97 -> ml_nms.cu:82-146, mask D2H +
                   host sweep included), from oracle/_ref/pet_ref_cuda.so (the unmodified reference kernel); when that
                   build is absent, torchvision.ops.batched_nms per image is timed instead and named."""
    import torchvision
    out = {}
    try:
        order = torch.argsort(seg.to(torch.int64), stable=True)
        counts = torch.bincount(seg.to(torch.int64), minlength=nseg).tolist()
        bs, ss = b[order].contiguous(), s[order].contiguous()
        if name.startswith("rpn"):
            chunks, pos = [], 0
            for c in counts:
                chunks.append((bs[pos:pos + c], ss[pos:pos + c]))
                pos += c
            fn = lambda: sum(int(torchvision.ops.nms(bb, sc, thr).numel()) for bb, sc in chunks if bb.shape[0])
            out["op"] = "torchvision.ops.nms per (image, level)"
        else:
            n_cls = 80
            labels = (seg.to(torch.int64) % n_cls + 1)[order].contiguous()
            per_img, pos = [], 0
            for i in range(nseg // n_cls):
                c = sum(counts[i * n_cls:(i + 1) * n_cls])
                per_img.append((bs[pos:pos + c], ss[pos:pos + c], labels[pos:pos + c]))
                pos += c
            try:
                from oracle import build_ref
                ref = build_ref.load("pet_ref_cuda")
                fn = lambda: sum(int(ref.ml_nms(bb, sc, lb, thr, 0).numel()) for bb, sc, lb in per_img if bb.shape[0])
                out["op"] = "_C.ml_nms (reference ml_nms.cu, unmodified) per image"
            except Exception:
                fn = lambda: sum(int(torchvision.ops.batched_nms(bb, sc, lb, thr).numel()) for bb, sc, lb in per_img if bb.shape[0])
                out["op"] = "torchvision.ops.batched_nms per image (reference build unavailable)"
        kept = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out.update({"ms": ms, "boxes_per_sec": b.shape[0] / (ms * 1e-3), "kept": kept, "same_kept_count": kept == kept_ours})
    except Exception as e:
        out["unavailable"] = repr(e)[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
