"""Builds libcpm_ops.so (the C-ABI CUDA library, include/cpm_ops.h) in-tree with nvcc for sm_100a.

    python -m cpm_r_cnn_b200.build [--force] [--verbose]

No torch involved: the library depends on the CUDA runtime only (cudart is linked statically), so the .so that
is built in the (GPU-less) build container is the one that runs on the B200 box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libcpm_ops.so")
OBJ_DIR = os.path.join(HERE, "build")

SOURCES = ["api.cu", "roi_align_fwd.cu", "roi_align_fwd_cols.cu", "roi_align_fwd_rows.cu", "roi_align_bwd.cu", "roi_align_bwd_tma.cu", "nms.cu", "grid_decode.cu", "rpn_decode.cu", "grid_targets.cu", "matcher.cu", "layout.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "roi_align_bwd.cuh"), os.path.join(INCLUDE, "cpm_ops.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: no implicit contraction -- every fused multiply-add in the kernels is an explicit fmaf(), which is what
# lets the arithmetic be pinned against the reference (see DESIGN.md "Numerics").
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", INCLUDE, "--expt-relaxed-constexpr"] + os.environ.get("CPM_NVCC_EXTRA", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
