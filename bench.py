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

# dram__bytes_read.sum + dram__bytes_write.sum per launch of each op's main kernel come from the committed digest of the
# `ncu --set full` capture of this same workload (written by tools/ncu_digest.py --json); a run under ncu is never timed,
# so bench.py only READS the file -- absent or stale (kernel name mismatch) => traffic is null
NCU_DRAM_FILE = os.path.join(ROOT, "profiles", "ncu_dram_r02.json")


def ncu_dram_bytes():
    try:
        return json.load(open(NCU_DRAM_FILE))
    except Exception:
        return {}


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
    prev_threads = torch.get_num_threads()
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
    torch.set_num_threads(prev_threads)
    sample = ("%d of %d RoIs/img x %d img, 7x7 + 14x14 fwd+bwd through the per-level Pooler loop, fp32, 4 levels x %d ch"
              % (sample_rois_per_img, ROIS_PER_IMG, IMGS_PER_GPU, CHANNELS))
    return units / dt, dt, kind, sample


def run_reference(args):
    """`--impl reference`: the reference's own CPU RoIAlign (unmodified ROIAlign_cpu.cpp, single-threaded by construction)
    on the SAME workload as the repo arm: every step pools all 512 RoIs/img of both images through both poolers, forward
    + backward.  Rank 0 alone runs it (one host core = one process); the line therefore says n_gpus 1 whatever --gpus is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates, times = [], []
    # one full step is ~4.5 s of single-thread CPU work; only a run asked for more than 40 steps is sampled down
    n_step = ROIS_PER_IMG if args.steps <= 40 else max(8, (40 * ROIS_PER_IMG) // args.steps)
    for i in range(args.warmup + args.steps):
        n = 8 if i < args.warmup else n_step
        r, dt, kind, sample = cpu_reference_rate(n)
        if i >= args.warmup:
            rates.append(r)
            times.append(dt)
    value = sum(rates) / len(rates)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "requested_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: R-50-FPN CPM head RoIAlign 7x7 + 14x14 fwd+bwd, %d img/GPU, %d RoIs/img, "
                                   "4-level 256-ch fp32 pyramid of 800x1344 images, sampling_ratio 2" % (IMGS_PER_GPU, ROIS_PER_IMG),
                       "sample": sample, "same_workload_as_repo_arm": n_step == ROIS_PER_IMG,
                       "note": "one CPU process on rank 0 (the kernel is single-threaded: ROIAlign_cpu.cpp:185-186 has its omp "
                               "pragma commented out); compare with the repo arm at N = 1"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def host_threads():
    """CPU threads this process may use (its affinity mask, not the machine's core count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_cores(index):
    """Pins this process to the CPU cores NVML reports as local to GPU `index` (before any pinned allocation, so that the
    staging buffers are first-touched on that NUMA node).  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"cpus": "%d-%d (%d)" % (allowed[0], allowed[-1], len(allowed)) if allowed else None}
    except Exception as ex:
        return {"unavailable": repr(ex)[:120]}


def capture(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    return g, keep


def time_graphs(graphs, iters, skip=2):
    """Median per-graph time (ms) over `iters` passes of the graph list, CUDA events on the launching stream."""
    n = len(graphs)
    rec = [[] for _ in range(n)]
    for it in range(iters + skip):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        for j, g in enumerate(graphs):
            e[j].record()
            g.replay()
        e[n].record()
        torch.cuda.synchronize()
        if it >= skip:
            for j in range(n):
                rec[j].append(e[j].elapsed_time(e[j + 1]))
    return [sorted(r)[len(r) // 2] for r in rec]


def reference_levels(rois):
    """LevelMapper with the reference's torch expressions (pet/rcnn/utils/poolers.py:29-40, k_min 2, k_max 5)."""
    area = (rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1)
    lv = torch.floor(4 + torch.log2(torch.sqrt(area) / 224 + 1e-6))
    return torch.clamp(lv, min=2, max=5).to(torch.int64) - 2


def reference_gpu_roi_align(feats_nchw, rois, gouts, scales, iters=5):
    """The reference's own single-GPU path for the same step, timed with CUDA events in this process: the UNMODIFIED
    ROIAlign_cuda.cu (oracle/_ref/pet_ref_cuda.so) through the reference Pooler's per-level loop -- LevelMapper in torch
    ops, nonzero() per level (host sync), _C.roi_align_forward per level (which ends in cudaDeviceSynchronize,
    ROIAlign_cuda.cu:422), index assignment (poolers.py:117-131) -- and the matching autograd backward: the index
    gather of the pooled gradient and one _C.roi_align_backward (at::zeros + atomicAdd scatter) per level."""
    out = {}
    try:
        from oracle import build_ref
        ref = build_ref.load("pet_ref_cuda")
        K, C = rois.shape[0], feats_nchw[0].shape[1]

        def fwd(p):
            lv = reference_levels(rois)
            res = torch.zeros((K, C, p[0], p[1]), dtype=feats_nchw[0].dtype, device=rois.device)
            for l, f in enumerate(feats_nchw):
                idx = torch.nonzero(lv == l).squeeze(1)
                res[idx] = ref.roi_align_forward(f, rois[idx], scales[l], p[0], p[1], SAMPLING, False, 0)
            return res

        def bwd(p, go):
            lv = reference_levels(rois)
            grads = []
            for l, f in enumerate(feats_nchw):
                idx = torch.nonzero(lv == l).squeeze(1)
                B, _, H, W = f.shape
                grads.append(ref.roi_align_backward(go[idx].contiguous(), rois[idx], scales[l], p[0], p[1], B, C, H, W,
                                                    SAMPLING, False, 0))
            return grads

        fns = []
        for p, go in zip(POOLERS, gouts):
            fns += [(lambda p=p: fwd(p)), (lambda p=p, go=go: bwd(p, go))]
        for fn in fns:
            fn()
        torch.cuda.synchronize()
        ms = [0.0] * 4
        for it in range(iters):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            for j, fn in enumerate(fns):
                e[j].record()
                fn()
            e[4].record()
            torch.cuda.synchronize()
            for j in range(4):
                ms[j] += e[j].elapsed_time(e[j + 1]) / iters
        out = {"op": "unmodified ROIAlign_cuda.cu (pet_ref_cuda.so) through the reference Pooler's per-level loop, NCHW fp32",
               "ms": dict(zip(["fwd7", "bwd7", "fwd14", "bwd14"], ms)), "ms_per_step": sum(ms),
               "rois_per_sec": K * len(POOLERS) / (sum(ms) * 1e-3)}
    except Exception as ex:
        out = {"unavailable": repr(ex)[:200]}
    return out


def bench_config0(dev):
    """configs[0]: ROIAlign 7x7 forward on ONE synthetic 800x1333 (padded 800x1344) R-50-FPN image, 512 RoIs, 256 channels:
    the reference's CPU kernel through the per-level loop (1 thread by construction) and torchvision's CPU roi_align
    (all torch threads) beside the GPU op on the same inputs."""
    from cpm_r_cnn_b200 import _lib, synthetic as sy
    from cpm_r_cnn_b200.roi_align import pooler_forward
    res = {}
    gen = torch.Generator().manual_seed(0)
    rois = sy.coco_like_rois(gen, 512, 1)
    feats = sy.pyramid(gen, 1, CHANNELS)
    lv = sy.fpn_levels_host(rois)
    scales = list(sy.FPN_SCALES)
    mapper = _lib.make_mapper(2, 5)
    fd = [f.to(dev) for f in feats]
    rd = rois.to(dev)
    fn = lambda: pooler_forward(fd, scales, rd, (7, 7), SAMPLING, False, 0, mapper)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g, keep = capture(fn)
    ms = time_graphs([g], 10)[0]
    res["ours_gpu"] = {"ms": ms, "rois_per_sec": 512 / (ms * 1e-3),
                       "note": "NCHW input: the graph holds the NHWC staging of the pyramid + the forward kernel"}
    out_gpu = keep.cpu()
    try:
        from oracle import build_ref
        ref = build_ref.load("pet_ref_cpu")
        prev_threads = torch.get_num_threads()
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        out = torch.zeros((512, CHANNELS, 7, 7))
        for l, f in enumerate(feats):
            idx = torch.nonzero(lv == l).squeeze(1)
            out[idx] = ref.roi_align_forward(f, rois[idx].contiguous(), scales[l], 7, 7, SAMPLING, False, 0)
        dt = time.perf_counter() - t0
        rms = float(out.pow(2).mean().sqrt())
        res["reference_cpu"] = {"ms": dt * 1e3, "rois_per_sec": 512 / dt, "cores": 1, "kind": "reference",
                                "max_abs_diff_vs_gpu": float((out - out_gpu).abs().max()),
                                "within_1e-5_rel": bool(((out - out_gpu).abs() <= 1e-5 * (out.abs() + rms)).all())}
        res["speedup_vs_reference_cpu"] = dt * 1e3 / ms
        torch.set_num_threads(prev_threads)
    except Exception as ex:
        res["reference_cpu"] = {"unavailable": repr(ex)[:200]}
    try:
        import torchvision
        prev_threads = torch.get_num_threads()
        torch.set_num_threads(host_threads())
        t0 = time.perf_counter()
        for l, f in enumerate(feats):
            idx = torch.nonzero(lv == l).squeeze(1)
            torchvision.ops.roi_align(f, rois[idx].contiguous(), (7, 7), scales[l], SAMPLING, False)
        dt = time.perf_counter() - t0
        res["torchvision_cpu"] = {"ms": dt * 1e3, "rois_per_sec": 512 / dt, "cores": torch.get_num_threads()}
        torch.set_num_threads(prev_threads)
    except Exception as ex:
        res["torchvision_cpu"] = {"unavailable": repr(ex)[:200]}
    return res


def bench_config4(dev, ra):
    """configs[4], op-level proxy (the model itself -- X-101-64x4d-FPN-DCN -- is out of scope): the
    inference-side use of the path on one rank: 2 images x 1000 RoIs/img, NCHW maps in, 7x7 box-head pooling and 14x14
    grid-head pooling forward only, staging of the pyramid inside the timed graph; the reference's own CUDA kernels through
    its per-level loop beside it."""
    from cpm_r_cnn_b200 import _lib, synthetic as sy
    from cpm_r_cnn_b200.roi_align import pooler_forward
    res = {}
    gen = torch.Generator().manual_seed(4)
    B, R = IMGS_PER_GPU, 1000
    rois_h = sy.coco_like_rois(gen, R, B)
    feats_h = sy.pyramid(gen, B, CHANNELS)
    feats = [f.to(dev) for f in feats_h]
    rois = rois_h.to(dev)
    scales = list(sy.FPN_SCALES)
    mapper = _lib.make_mapper(2, 5)

    def step():
        ra.STAGING_CACHE.clear()
        return [pooler_forward(feats, scales, rois, p, SAMPLING, False, 0, mapper) for p in POOLERS]

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g, _ = capture(step)
    ms = time_graphs([g], 10)[0]
    res["ours"] = {"ms": ms, "rois_per_sec": 2 * B * R / (ms * 1e-3), "rois": B * R,
                   "note": "7x7 + 14x14 forward, NCHW in, NHWC staging inside; RoI-units = RoIs x 2 poolers"}
    try:
        from oracle import build_ref
        ref = build_ref.load("pet_ref_cuda")
        lv = sy.fpn_levels_host(rois_h).to(dev)

        def ref_step():
            outs = []
            for p in POOLERS:
                out = torch.zeros((B * R, CHANNELS, p[0], p[1]), device=dev)
                for l, f in enumerate(feats):
                    idx = torch.nonzero(lv == l).squeeze(1)       # poolers.py:127-130 (host sync included)
                    out[idx] = ref.roi_align_forward(f, rois[idx], scales[l], p[0], p[1], SAMPLING, False, 0)
                outs.append(out)
            return outs

        for _ in range(2):
            ref_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ref_step()
        e1.record()
        torch.cuda.synchronize()
        rms = e0.elapsed_time(e1) / 5
        res["reference_gpu"] = {"ms": rms, "rois_per_sec": 2 * B * R / (rms * 1e-3),
                                "op": "unmodified ROIAlign_cuda.cu through the reference Pooler's per-level loop"}
        res["speedup_vs_reference_gpu"] = rms / ms
    except Exception as ex:
        res["reference_gpu"] = {"unavailable": repr(ex)[:200]}
    return res


def bench_allreduce(dev, world, dist, sync_all):
    """configs[3], the collective of the surrounding training step (north_star: NCCL is used only for DDP's gradient
    all-reduce, never inside the ops): one all-reduce of a gradient set the size of R-101-FPN CPM R-CNN's (~62 M fp32
    parameters, 25 MB buckets as DDP would issue them), timed on the device, max over ranks."""
    if world <= 1:
        return {}
    n_params, bucket = 62_000_000, 25 * 1024 * 1024 // 4
    bufs = [torch.zeros(min(bucket, n_params - i), device=dev) for i in range(0, n_params, bucket)]

    def step():
        for b in bufs:
            dist.all_reduce(b)

    for _ in range(3):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    nbytes = 4 * n_params
    return {"ms": ms, "bytes": nbytes, "buckets": len(bufs), "algbw_gbs": nbytes / (ms * 1e-3) / 1e9,
            "busbw_gbs": nbytes / (ms * 1e-3) / 1e9 * 2 * (world - 1) / world,
            "note": "NCCL all-reduce of a 62 M-parameter fp32 gradient set in 25 MB buckets (what DDP adds around the ops)"}


def run_ours(args):
    import torch.distributed as dist
    import cpm_r_cnn_b200 as ops
    import importlib
    from cpm_r_cnn_b200 import _lib, sharding, synthetic as sy
    ra = importlib.import_module("cpm_r_cnn_b200.roi_align")    # the package re-exports a function of the same name
    from cpm_r_cnn_b200.roi_align import pooler_backward, pooler_forward

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the cpm_ops path has no CPU fallback")
    affinity = bind_to_gpu_cores(local)       # before CUDA / pinned allocations: NUMA-local staging buffers
    torch.set_num_threads(host_threads())     # torch's intra-op pool follows the affinity mask, not the machine's core count
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly one line, the JSON: everything libraries print while the job runs (NCCL writes its version
    # banner to stdout when the first communicator is created) is routed to stderr, and the real stdout is restored for
    # the result
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")
    _lib.lib()
    skip = os.environ.get("CPM_BENCH_SKIP", "")

    rois_h, feats_h, gouts_h = make_workload(rank)
    shapes = [tuple(f.shape) for f in feats_h]
    scales = list(sy.FPN_SCALES)
    mapper = _lib.make_mapper(2, 5)
    K = rois_h.shape[0]
    rois = rois_h.to(dev)
    gouts = [g.to(dev) for g in gouts_h]
    # what the reference's FPN hands the Pooler: NCHW-contiguous maps (FPN.py:96-121) ...
    feats_nchw = [f.to(dev) for f in feats_h]
    # ... and the same maps as a channels_last backbone would emit them (zero-copy NHWC for the kernels)
    feats = [f.contiguous(memory_format=torch.channels_last) for f in feats_nchw]
    names = ["fwd7", "bwd7", "fwd14", "bwd14"]

    def sync_all():
        # Ranks that arrive early wait on the HOST (gloo) first, so that the NCCL barrier kernel never sits spinning on a
        # GPU while rank 0 is still running one of its rank-0-only arms: a pending NCCL kernel on the peer was measured to
        # slow rank 0's host-synchronising arms 14x (detection post-process 0.98 -> 14 ms at N = 2).
        if world > 1:
            dist.barrier(group=host_group)
            dist.barrier()
        torch.cuda.synchronize()

    side = torch.cuda.Stream(dev)

    def warm(fns, n=3):
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(n):
                for fn in fns:
                    fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        sync_all()

    # ---- HEADLINE: the step as the unmodified reference model feeds it -- NCHW maps in, NCHW gradients out.  The NHWC
    #      staging of the pyramid is INSIDE the timed region, once per step (the staging cache serves the second pooler,
    #      as it serves poolers 2..5 of a CPM iteration); everything is captured once in ONE CUDA graph through the Python
    #      op layer and replayed, because the ops are tens of microseconds each ----
    def step_nchw():
        ra.STAGING_CACHE.clear()
        keep = []
        for p, go in zip(POOLERS, gouts):
            keep.append(pooler_forward(feats_nchw, scales, rois, p, SAMPLING, False, 0, mapper))
            keep.append(pooler_backward(go, shapes, scales, rois, p, SAMPLING, False, 0, mapper, nchw_grad=True))
        return keep

    warm([step_nchw])
    l0 = _lib.launch_count()
    g_step, keep_step = capture(step_nchw)
    launches_per_step = _lib.launch_count() - l0
    assert all(t.is_contiguous() for t in keep_step[1]) and all(t.is_contiguous() for t in keep_step[3])
    for _ in range(max(args.warmup, 3)):
        g_step.replay()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0 and "clocks" not in skip:
        sampler.start()
    # the contract's K-step region, repeated: the median repeat is the value (a 10 ms region is at the mercy of one hiccup)
    repeats = []
    n_rep = 5
    for _ in range(n_rep):
        sync_all()
        t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_beg.record()
        for i in range(args.steps):
            g_step.replay()
        t_end.record()
        sync_all()
        repeats.append(t_beg.elapsed_time(t_end) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    launches = launches_per_step * args.steps
    sec_med = sorted(repeats)[n_rep // 2]
    # whole-job figure: units of all ranks / the slowest rank's device time (cpm_r_cnn_b200/sharding.py)
    units_total, sec_total, value = sharding.aggregate(K * len(POOLERS) * args.steps, sec_med)
    ms_step = sec_total * 1e3 / args.steps
    units_per_step = K * len(POOLERS) * world

    # ---- per-op times: pyramid resident as channels_last (NHWC, zero-copy), one graph per op ----
    def op_fwd(p, cl=False):
        return lambda: pooler_forward(feats, scales, rois, p, SAMPLING, False, 0, mapper, channels_last=cl)

    def op_bwd(p, go, **kw):
        return lambda: pooler_backward(go, shapes, scales, rois, p, SAMPLING, False, 0, mapper, **kw)

    op_fns = []
    for p, go in zip(POOLERS, gouts):
        op_fns += [op_fwd(p), op_bwd(p, go)]
    warm(op_fns)
    caps = [capture(fn) for fn in op_fns]
    n_it = max(5, min(args.steps, 30))
    op_list = time_graphs([c[0] for c in caps], n_it)
    op_ms = dict(zip(names, op_list))
    if world > 1:
        t = torch.tensor(op_list, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        op_ms = dict(zip(names, [float(v) for v in t.tolist()]))
    del caps

    # ---- pieces of the NCHW step, one graph each: the staging alone, the backward writing NCHW ----
    nchw_ms = {}
    try:
        st_fn = lambda: ra.stage_pyramid_nhwc(feats_nchw, cache=False)
        b_fns = [op_bwd(p, go, nchw_grad=True) for p, go in zip(POOLERS, gouts)]
        warm([st_fn] + b_fns, 2)
        caps = [capture(fn) for fn in [st_fn] + b_fns]
        t = time_graphs([c[0] for c in caps], n_it)
        nchw_ms = {"stage_pyramid_nhwc": t[0], "bwd7_nchw_out": t[1], "bwd14_nchw_out": t[2]}
        del caps
    except Exception as ex:
        nchw_ms = {"error": repr(ex)[:200]}
    sync_all()

    # ---- the same four ops with channels_last pooled tensors (extension): the forward returns the pooled block with
    #      channels_last strides and the backward reads a channels_last gradient in place ----
    cl_ms = {}
    try:
        if "cl" in skip:
            raise RuntimeError("skipped")
        gouts_cl = [g.contiguous(memory_format=torch.channels_last) for g in gouts]
        cl_fns = []
        for p, go in zip(POOLERS, gouts_cl):
            cl_fns += [op_fwd(p, True), op_bwd(p, go)]
        warm(cl_fns, 2)
        caps = [capture(fn) for fn in cl_fns]
        cl_ms = dict(zip(names, time_graphs([c[0] for c in caps], n_it)))
        del caps, gouts_cl
    except Exception as ex:      # reported, never fatal for the headline
        cl_ms = {"error": repr(ex)[:200]}
    sync_all()

    # ---- forward on a bf16 pyramid (bf16 pooled output, fp32 arithmetic): reported separately, own algorithmic bytes
    #      (2 bytes per stored element); the tolerance of this path is stated in tests/test_gpu_parity.py ----
    bf16_ms = {}
    try:
        if "bf16" in skip:
            raise RuntimeError("skipped")
        feats16 = [f.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in feats]
        b_fns = [(lambda p=p: pooler_forward(feats16, scales, rois, p, SAMPLING, False, 0, mapper)) for p in POOLERS]
        warm(b_fns, 2)
        caps = [capture(fn) for fn in b_fns]
        bf16_ms = dict(zip(["fwd%d" % p[0] for p in POOLERS], time_graphs([c[0] for c in caps], n_it)))
        del caps, feats16
    except Exception as ex:
        bf16_ms = {"error": repr(ex)[:200]}
    sync_all()

    # ---- the red.global.add fallback of the backward (non-deterministic order), reported separately ----
    atomic_ms = {}
    try:
        at_fns = [op_bwd(p, go, mode="atomic") for p, go in zip(POOLERS, gouts)]
        warm(at_fns, 2)
        caps = [capture(fn) for fn in at_fns]
        atomic_ms = dict(zip(["bwd7", "bwd14"], time_graphs([c[0] for c in caps], 5)))
        del caps
    except Exception as ex:
        atomic_ms = {"error": repr(ex)[:200]}
    sync_all()

    # ---- the reference's own GPU RoIAlign on the same step (rank 0) ----
    ref_gpu = reference_gpu_roi_align(feats_nchw, rois, gouts, scales) if rank == 0 and "refgpu" not in skip else {}
    sync_all()

    # ---- e2e: host (pinned) buffers in and out, copies inside the timed region ----
    # One step = H2D of the NCHW pyramid, the RoIs and both pooled gradients, the two Pooler modules forward + one autograd
    # backward (the feature gradient of both poolers accumulates into x.grad, as in the head), D2H of both pooled outputs
    # and the NCHW gradient pyramid.  Three streams (H2D / compute / D2H) and two device input sets let step i+1's upload
    # run under step i's compute and download (PCIe is full duplex); every step still moves all of its bytes.
    def pinned(shape):
        return torch.empty(shape, pin_memory=True)
    feats_pin = [pinned(f.shape).copy_(f) for f in feats_h]
    rois_pin = pinned(rois_h.shape).copy_(rois_h)
    gouts_pin = [pinned(g.shape).copy_(g) for g in gouts_h]
    outs_pin = [pinned((K, CHANNELS, p[0], p[1])) for p in POOLERS]
    grads_pin = [pinned(s_) for s_ in shapes]
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

    def e2e_steps_run(n, upload_pyramid=True, download_grads=True):
        for i in range(n):
            d = dev_sets[i & 1]
            with torch.cuda.stream(st_in):
                st_in.wait_event(ev_c[i & 1])          # the compute that last read this input set has finished
                src = (feats_pin if upload_pyramid else []) + gouts_pin + [rois_pin]
                dst = (d["feats"] if upload_pyramid else []) + d["gouts"] + [d["rois"]]
                for dd, ss in zip(dst, src):
                    dd.copy_(ss, non_blocking=True)
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
                res = [o.detach() for o in outs] + ([x.grad for x in xs] if download_grads else [])
                for dd, ss in zip(outs_pin + grads_pin, res):
                    dd.copy_(ss, non_blocking=True)
                ev_out[i & 1].record(st_out)
            hold[i & 1] = (outs, xs, res, boxlists)
            del outs, xs, res, boxlists
        for s_ in (st_in, st_c, st_out):
            s_.synchronize()

    def e2e_measure(**kw):
        n = max(10, min(args.steps, 50))
        e2e_steps_run(6, **kw)    # untimed: lets torch's caching allocator reach its steady state on all three streams
        sync_all()
        t0 = time.perf_counter()
        e2e_steps_run(n, **kw)
        sync_all()
        ms = (time.perf_counter() - t0) * 1e3 / n
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n

    e2e_ms, e2e_steps = e2e_measure()
    e2e_value = units_per_step / (e2e_ms * 1e-3)
    # the case a head inside a resident model sees: the pyramid and its gradient stay on the device; per step the RoIs and
    # the pooled gradients come from the host and the pooled outputs go back
    res_ms, _ = e2e_measure(upload_pyramid=False, download_grads=False)
    h2d_res = sum(t.numel() * 4 for t in gouts_pin) + rois_pin.numel() * 4
    d2h_res = sum(t.numel() * 4 for t in outs_pin)
    # bare copies of the same bytes, both directions at once: the PCIe / host-memory ceiling of this rank on this box
    def bare_copies(n=6):
        t0 = None
        for i in range(n + 2):
            if i == 2:
                sync_all()
                t0 = time.perf_counter()
            with torch.cuda.stream(st_in):
                for dd, ss in zip(dev_sets[0]["feats"] + dev_sets[0]["gouts"], feats_pin + gouts_pin):
                    dd.copy_(ss, non_blocking=True)
            with torch.cuda.stream(st_out):
                for dd, ss in zip(outs_pin + grads_pin, keep_step[0:1] + keep_step[2:3] + keep_step[1] + keep_step[3]):
                    dd.copy_(ss, non_blocking=True)
        st_in.synchronize()
        st_out.synchronize()
        sync_all()
        return (time.perf_counter() - t0) * 1e3 / n
    try:
        bare_ms = bare_copies()
        if world > 1:
            t = torch.tensor([bare_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            bare_ms = float(t.item())
    except Exception:
        bare_ms = None

    # ---- NMS half of the metric (configs[2]) ----
    nms = bench_nms(ops, dev, rank, world, dist, sync_all)
    decode = bench_decode(ops, dev, rank)
    rpn = bench_rpn(ops, dev, rank)
    det = bench_detect(ops, dev, rank)
    gtg = bench_grid_targets(ops, dev, rank)
    mat = bench_matcher(ops, dev, rank)
    cfg0 = bench_config0(dev) if rank == 0 and "config0" not in skip else {}
    cfg4 = bench_config4(dev, ra) if rank == 0 and "config4" not in skip else {}
    sync_all()
    ddp = bench_allreduce(dev, world, dist, sync_all) if "allreduce" not in skip else {}

    if rank == 0:
        peak, peak_src = measured_peak()
        ab = algorithmic_bytes(rois_h)
        dram = ncu_dram_bytes()
        rl_ops = {n: {"ms": op_ms[n], "bytes": ab[n], "gbs": ab[n] / (op_ms[n] * 1e-3) / 1e9,
                      "frac": ab[n] / (op_ms[n] * 1e-3) / 1e9 / peak, "frac_of_nominal_8TBs": ab[n] / (op_ms[n] * 1e-3) / 1e9 / 8000.0,
                      "traffic": dram.get(n)}
                  for n in names}
        top = max(names, key=lambda n: op_ms[n])
        total_bytes = sum(ab[n] for n in names)
        resident_ms = sum(op_ms[n] for n in names)
        # CPU baseline beside it: the full workload of the step (512 RoIs/img x 2 img), 4 passes (~18 s of CPU work)
        cpu_rate = cpu_dt = cpu_kind = cpu_sample = None
        if world == 1:
            runs = [cpu_reference_rate(ROIS_PER_IMG) for _ in range(4)]
            cpu_dt = sum(r[1] for r in runs)
            cpu_rate = sum(r[0] * r[1] for r in runs) / cpu_dt          # units / total seconds
            cpu_kind, cpu_sample = runs[0][2], runs[0][3] + ", 4 passes"
        kernel_names = {"fwd7": "roi_align_fwd_cols<1> (7x7)", "fwd14": "roi_align_fwd_cols<2> (14x14)",
                        "bwd7": "bwd_tiles_staged<0,7,2> (7x7) + bwd_prepare", "bwd14": "bwd_tiles_staged<0,14,2> (14x14) + bwd_prepare"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: R-50-FPN CPM head RoIAlign 7x7 + 14x14 fwd+bwd, %d img/GPU, %d RoIs/img, "
                                       "4-level 256-ch fp32 pyramid of 800x1344 images, sampling_ratio 2" % (IMGS_PER_GPU, ROIS_PER_IMG),
                           "layout": "what the unmodified reference feeds and expects: NCHW-contiguous feature maps in (FPN.py:96-121), "
                                     "pooled output (K,C,PH,PW), NCHW-contiguous feature gradients out; the NHWC staging of the "
                                     "pyramid runs INSIDE the timed step, once per step (staging cache shared by the poolers)",
                           "backward": "deterministic tile-owner gather (no atomics)",
                           "unit_of_work": "one RoI through one pooler forward+backward; %d per step per GPU" % (K * len(POOLERS)),
                           "launch": "the whole step captured once in a CUDA graph through the Python op layer and replayed; the "
                                     "K-step region is timed with CUDA events on the launching stream, %d times, median reported "
                                     "(all repeats in `repeats_ms`)" % n_rep,
                           "l2": "not flushed: per-step working set (pyramid 183 MB + staged copy 183 MB + pooled/grad_out 514 MB + "
                                 "gradients 366 MB) exceeds the 126 MB L2",
                           "parallelism": "dp%d (images sharded per GPU, no collective inside the ops)" % world,
                           "cpu_affinity": affinity},
                "repeats_ms": [r * 1e3 for r in repeats],
                "value_resident_channels_last": {"value": units_per_step / (resident_ms * 1e-3), "ms_per_step": resident_ms,
                                                 "note": "the round-1 headline: the same four ops on a pyramid that is already "
                                                         "channels_last (zero-copy NHWC) with channels_last-strided gradients, sum "
                                                         "of the per-op graph times; comparable with BENCH_r01.value"},
                "roofline": {"bound": "hbm", "kernel": kernel_names[top],
                             "achieved": rl_ops[top]["gbs"], "peak": peak, "unit": "GB/s", "frac": rl_ops[top]["frac"],
                             "traffic": dram.get(top), "traffic_source": os.path.relpath(NCU_DRAM_FILE, ROOT) if dram else None,
                             "peak_source": peak_src,
                             "step": {"bytes": total_bytes, "gbs": total_bytes / (ms_step * 1e-3) / 1e9,
                                      "frac": total_bytes / (ms_step * 1e-3) / 1e9 / peak,
                                      "note": "algorithmic bytes of the four ops (staging is overhead, SURVEY.md 8d) / the NCHW step"},
                             "step_resident_channels_last": {"bytes": total_bytes, "gbs": total_bytes / (resident_ms * 1e-3) / 1e9,
                                                             "frac": total_bytes / (resident_ms * 1e-3) / 1e9 / peak},
                             "ops": rl_ops, "nchw_pieces_ms": nchw_ms, "U_px": {"7x7": ab["U7"], "14x14": ab["U14"]},
                             "backward_atomic_fallback_ms": atomic_ms,
                             "fwd_bf16_storage": ({n: {"ms": bf16_ms[n], "bytes": (ab[n] - 20 * K) // 2 + 20 * K,
                                                       "gbs": ((ab[n] - 20 * K) // 2 + 20 * K) / (bf16_ms[n] * 1e-3) / 1e9,
                                                       "frac": ((ab[n] - 20 * K) // 2 + 20 * K) / (bf16_ms[n] * 1e-3) / 1e9 / peak}
                                                   for n in bf16_ms} if "error" not in bf16_ms else bf16_ms),
                             "ops_channels_last_pooled": ({n: {"ms": cl_ms[n], "gbs": ab[n] / (cl_ms[n] * 1e-3) / 1e9,
                                                               "frac": ab[n] / (cl_ms[n] * 1e-3) / 1e9 / peak} for n in names}
                                                          if "error" not in cl_ms else cl_ms)},
                "roi_align": {"reference_gpu": ref_gpu,
                              "speedup_vs_reference_gpu": (ref_gpu["ms_per_step"] / ms_step if world == 1 and "ms_per_step" in ref_gpu
                                                           else None)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms, "steps": e2e_steps,
                        "h2d_gbs_per_rank": h2d / (e2e_ms * 1e-3) / 1e9, "d2h_gbs_per_rank": d2h / (e2e_ms * 1e-3) / 1e9,
                        "bare_copies_ms": bare_ms,
                        "note": "NCHW pyramid, RoIs and pooled gradients up, pooled outputs and NCHW gradient pyramid down, every "
                                "step; bare_copies_ms = the same bytes with no kernels (this rank's PCIe / host-memory ceiling)",
                        "pyramid_resident": {"value": units_per_step / (res_ms * 1e-3), "ms_per_step": res_ms,
                                             "h2d_bytes_per_step": h2d_res, "d2h_bytes_per_step": d2h_res,
                                             "note": "the case a head inside a resident model sees: pyramid and its gradient stay "
                                                     "on the device; RoIs + pooled gradients up, pooled outputs down"}},
                "gpu_launches": int(launches), "clocks": clocks, "nms": nms, "grid_decode": decode, "rpn_proposals": rpn,
                "detection_postprocess": det, "grid_targets": gtg, "iou_matcher": mat, "config0": cfg0,
                "config4_inference_proxy": cfg4, "config3_ddp_allreduce": ddp}
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": cpu_kind, "sample": cpu_sample,
                                    "seconds": cpu_dt}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        sync_all()
        dist.destroy_process_group()


def bench_nms(ops, dev, rank, world, dist, sync_all, iters=10):
    """configs[2]: batched NMS, 16 images: RPN flavour (5 levels x 1000 proposals, thr 0.7) and detection flavour
    (1000 proposals x 80 classes, score > 0.03 gate and un-gated stress, thr 0.3).  boxes/s = input boxes / time."""
    from cpm_r_cnn_b200 import _lib, synthetic as sy
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
            flavor = _lib.IOU_TV_CUDA if name.startswith("rpn") else _lib.IOU_ML_CUDA
            ours_keep, ours_counts = ops.batched_nms(b, s, seg, nseg, thr, iou_flavor=flavor, return_counts=True)
            res[name]["reference_gpu"] = reference_gpu_nms(name, b, s, seg, nseg, thr, ours_keep, ours_counts)
            if res[name]["reference_gpu"].get("ms"):
                res[name]["speedup_vs_reference_gpu"] = res[name]["reference_gpu"]["ms"] * world / ms if world == 1 else None
            res[name]["cpu_baseline"] = cpu_nms_baseline(name, b, s, seg, nseg, thr, ours_keep)
    return res


def cpu_nms_baseline(name, b, s, seg, nseg, thr, ours_keep):
    """The extension ships no CPU nms / ml_nms (ml_nms.h:38 "CPU version not implemented"), so -- as BASELINE.json's
    north_star prescribes -- the equivalent torchvision.ops CPU op is timed on the host cores: one torchvision.ops.nms per
    (image, level) segment for the RPN flavour, one torchvision.ops.batched_nms per image for the detection flavour."""
    import torchvision
    out = {}
    try:
        bc, sc, sg = b.cpu(), s.cpu(), seg.cpu().to(torch.int64)
        order = torch.argsort(sg, stable=True)
        counts = torch.bincount(sg, minlength=nseg).tolist()
        bs, ss, gs = bc[order].contiguous(), sc[order].contiguous(), sg[order].contiguous()
        prev_threads = torch.get_num_threads()
        torch.set_num_threads(host_threads())
        t0 = time.perf_counter()
        kept, pos = [], 0
        if name.startswith("rpn"):
            for c in counts:
                if c:
                    kept.append(order[pos + torchvision.ops.nms(bs[pos:pos + c], ss[pos:pos + c], thr)])
                pos += c
            op = "torchvision.ops.nms (CPU) per (image, level)"
        else:
            n_cls = 80
            for i in range(nseg // n_cls):
                c = sum(counts[i * n_cls:(i + 1) * n_cls])
                if c:
                    kept.append(order[pos + torchvision.ops.batched_nms(bs[pos:pos + c], ss[pos:pos + c], gs[pos:pos + c], thr)])
                pos += c
            op = "torchvision.ops.batched_nms (CPU) per image"
        dt = time.perf_counter() - t0
        cores = torch.get_num_threads()
        torch.set_num_threads(prev_threads)
        kept = torch.cat(kept) if kept else torch.empty(0, dtype=torch.int64)
        out = {"op": op, "ms": dt * 1e3, "boxes_per_sec": b.shape[0] / dt, "cores": cores, "kind": "torchvision",
               "same_keep_set_as_ours": bool(torch.equal(torch.sort(kept)[0], torch.sort(ours_keep.cpu())[0]))}
    except Exception as e:
        out = {"unavailable": repr(e)[:200]}
    return out


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
        eager_ms = e0.elapsed_time(e1) / iters          # back-to-back eager calls: host enqueue rate for small R
        g, keep = capture(lambda: ops.grid_decode(logits, boxes, sub, 0.5))
        ms = time_graphs([g], iters)[0]                 # the kernel itself (CUDA-graph replay)
        nbytes = R * 9 * 28 * 28 * 4 + 32 * R
        res["R%d" % R] = {"ms": ms, "eager_call_ms": eager_ms, "rois_per_sec": R / (ms * 1e-3), "bytes": nbytes,
                          "gbs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
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


def reference_gpu_nms(name, b, s, seg, nseg, thr, ours_keep, ours_counts):
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
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        if name.startswith("rpn"):
            chunks = [(bs[starts[i]:starts[i + 1]], ss[starts[i]:starts[i + 1]], starts[i]) for i in range(nseg) if counts[i]]
            fn = lambda: [order[p0 + torchvision.ops.nms(bb, sc, thr)] for bb, sc, p0 in chunks]
            out["op"] = "torchvision.ops.nms per (image, level)"
        else:
            n_cls = 80
            labels = (seg.to(torch.int64) % n_cls + 1)[order].contiguous()
            per_img = []
            for i in range(nseg // n_cls):
                p0, p1 = starts[i * n_cls], starts[(i + 1) * n_cls]
                if p1 > p0:
                    per_img.append((bs[p0:p1], ss[p0:p1], labels[p0:p1], p0))
            try:
                from oracle import build_ref
                ref = build_ref.load("pet_ref_cuda")
                fn = lambda: [order[p0 + ref.ml_nms(bb, sc, lb, thr, 0)] for bb, sc, lb, p0 in per_img]
                out["op"] = "_C.ml_nms (reference ml_nms.cu, unmodified) per image"
            except Exception:
                fn = lambda: [order[p0 + torchvision.ops.batched_nms(bb, sc, lb, thr)] for bb, sc, lb, p0 in per_img]
                out["op"] = "torchvision.ops.batched_nms per image (reference build unavailable)"
        ref_keep = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        # keep LISTS, not counts: per (image, level) segment in the reference's own order for the RPN flavour; per image as a
        # set for the detection flavour (ml_nms orders an image's survivors by score over all classes, ours groups by class)
        if name.startswith("rpn"):
            same = bool(torch.equal(torch.cat(ref_keep), ours_keep))
        else:
            oc = torch.cumsum(ours_counts.reshape(-1, 80).sum(1), 0).tolist()
            ours_img = [ours_keep[(oc[i - 1] if i else 0):oc[i]] for i in range(len(oc))]
            ours_img = [k for k in ours_img if k.numel()]
            same = len(ours_img) == len(ref_keep) and all(
                torch.equal(torch.sort(a)[0], torch.sort(r)[0]) for a, r in zip(ours_img, ref_keep))
        out.update({"ms": ms, "boxes_per_sec": b.shape[0] / (ms * 1e-3), "kept": int(sum(k.numel() for k in ref_keep)),
                    "same_keep_lists": same})
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
