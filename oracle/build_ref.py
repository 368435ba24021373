"""TEST INFRASTRUCTURE ONLY -- builds the reference's own ops into oracle/_ref/ (git-ignored).

Compiles the reference sources *in place* (from /root/reference, nothing is copied) together with
oracle/ref_bind.cpp:

  oracle/_ref/pet_ref_cpu.so   ROIAlign/ROIAlign_cpu.cpp + NMS/soft_nms.cpp            (g++)
  oracle/_ref/pet_ref_cuda.so  the above + ROIAlign/ROIAlign_cuda.cu + NMS/ml_nms.cu   (nvcc, sm_100a)

The reference's own build system (setup.py / make.sh) is not run.  /root/reference only exists in the
build container; on the GPU box the prebuilt .so files are used as they travelled with the snapshot.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/pet/lib/ops/csrc"
OUT = os.path.join(HERE, "_ref")


def have_reference():
    return os.path.isdir(REF_CSRC)


def build(cuda=True, verbose=False):
    """Build (or rebuild) the reference ops. Returns the list of built .so paths."""
    if not have_reference():
        return [p for p in (os.path.join(OUT, "pet_ref_cpu.so"), os.path.join(OUT, "pet_ref_cuda.so"))
                if os.path.exists(p)]
    from torch.utils import cpp_extension

    os.makedirs(OUT, exist_ok=True)
    built = []
    jobs = [("pet_ref_cpu", False)]
    if cuda:
        jobs.append(("pet_ref_cuda", True))
    for name, with_cuda in jobs:
        so = os.path.join(OUT, name + ".so")
        if os.path.exists(so):
            built.append(so)
            continue
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        srcs = [os.path.join(HERE, "ref_bind.cpp"),
                os.path.join(REF_CSRC, "ROIAlign", "ROIAlign_cpu.cpp"),
                os.path.join(REF_CSRC, "NMS", "soft_nms.cpp")]
        cflags = ["-O2", "-w"]
        cuda_flags = []
        if with_cuda:
            srcs += [os.path.join(REF_CSRC, "ROIAlign", "ROIAlign_cuda.cu"),
                     os.path.join(REF_CSRC, "NMS", "ml_nms.cu")]
            cflags.append("-DWITH_CUDA")
            # the reference's setup.py:35-40 flags + an explicit Blackwell target
            cuda_flags = ["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                          "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                          "-gencode", "arch=compute_100a,code=sm_100a", "-w"]
            os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        cpp_extension.load(name=name, sources=srcs, extra_include_paths=[REF_CSRC],
                           extra_cflags=cflags, extra_cuda_cflags=cuda_flags,
                           with_cuda=with_cuda, build_directory=bdir, verbose=verbose,
                           is_python_module=False)
        shutil.copy(os.path.join(bdir, name + ".so"), so)
        shutil.rmtree(bdir, ignore_errors=True)
        built.append(so)
    return built


def load(name):
    """Import oracle/_ref/<name>.so as a python module (torch must be importable)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols)
    so = os.path.join(OUT, name + ".so")
    if not os.path.exists(so):
        raise FileNotFoundError(so)
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("\n".join(build(cuda="--no-cuda" not in sys.argv, verbose="-v" in sys.argv)))
