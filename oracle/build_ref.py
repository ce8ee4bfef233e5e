"""Build oracle/_ref/libchipmunk_ref_indexed_io.so from the reference's OWN sources, where they lie.

    python oracle/build_ref.py [--force]

Test infrastructure only.  The reference's attention / MLP kernels are ThunderKittens + wgmma
(`-arch=sm_90a`, /root/reference/setup.py:76,101-105) and cannot be built for B200; its three index
kernels are arch-generic CUDA + CUB + cuRAND and compile unmodified for sm_100a:

    /root/reference/csrc/indexed_io/mask_to_indices.cu
    /root/reference/csrc/indexed_io/topk_indices.cu
    /root/reference/csrc/indexed_io/copy_indices.cu
    /root/reference/csrc/indexed_io/scatter_add.cu

plus `scatter_add.cu` (needs the vendored ThunderKittens headers of the reference tree, no Hopper-only code is
instantiated).  They are compiled straight from
/root/reference (never copied into this repo) together with `oracle/ref_shim.cpp`, which registers them
as `torch.ops.chipmunk_ref.*`.  Output goes to `oracle/_ref/` only: git-ignored, NOT gpurun-ignored, so
the built library travels to the GPU box where `tests/test_ref_kernels_gpu.py` runs the reference kernels
next to ours.  /root/reference does not exist on the GPU box: nothing there rebuilds this.
nvcc spends several minutes per file on <torch/extension.h>; the three files build in parallel.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CHIPMUNK_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libchipmunk_ref_indexed_io.so")
SOURCES = [os.path.join(REF, "csrc", "indexed_io", f)
           for f in ("mask_to_indices.cu", "topk_indices.cu", "copy_indices.cu", "scatter_add.cu")]
# scatter_add.cu includes the reference's vendored ThunderKittens headers (header-only, in the reference tree) and
# csrc/common; only its bulk reduce-add helper is instantiated, which is plain sm_90+ PTX and builds for sm_100a
TK_INCLUDES = [os.path.join(REF, "submodules", "ThunderKittens", "include"), os.path.join(REF, "csrc", "common")]
SHIM = os.path.join(HERE, "ref_shim.cpp")


# The reference's two Triton kernels (the ONLY mm2 implementation it has, and the fp8-capable mm1) are arch-portable:
# they JIT for sm_100 on the GPU box.  They are staged -- byte-for-byte, by this recipe, into the git-ignored
# oracle/_ref/triton_ref/ -- so that tests/test_ref_triton_gpu.py and bench.py's `reference_gpu` leg can import them
# there (/root/reference does not exist on the GPU box; nothing of them enters the repo's history).
TRITON_SOURCES = [os.path.join(REF, "src", "chipmunk", "triton", f) for f in ("csp_mlp_mm1.py", "csp_mlp_mm2.py")]
TRITON_OUT = os.path.join(OUT, "triton_ref")


def stage_triton(force: bool = False) -> str | None:
    if not all(os.path.exists(s) for s in TRITON_SOURCES):
        return TRITON_OUT if os.path.isdir(TRITON_OUT) else None
    os.makedirs(TRITON_OUT, exist_ok=True)
    for s in TRITON_SOURCES:
        dst = os.path.join(TRITON_OUT, os.path.basename(s))
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(s):
            shutil.copyfile(s, dst)
    return TRITON_OUT


# The reference's Python host side of the path -- its op wrappers (padding / slicing around the kernels), its modules and its
# util package -- is staged the same way into oracle/_ref/chipmunk_py/chipmunk/{ops,modules,util}: tests/test_ref_python_gpu.py
# runs THAT code, unmodified, on top of this repo's `torch.ops.chipmunk.*` kernels (INTEGRATION.md option 2) and compares with the
# golden vectors.  `triton/` and the compiled `cuda` module are not staged: the runner supplies empty modules for them.
PY_PACKAGES = ("ops", "modules", "util")
PY_OUT = os.path.join(OUT, "chipmunk_py")


def stage_python(force: bool = False) -> str | None:
    src_root = os.path.join(REF, "src", "chipmunk")
    if not all(os.path.isdir(os.path.join(src_root, p)) for p in PY_PACKAGES):
        return PY_OUT if os.path.isdir(PY_OUT) else None
    for pkg in PY_PACKAGES:
        for d, _, files in os.walk(os.path.join(src_root, pkg)):
            rel = os.path.relpath(d, src_root)
            if "__pycache__" in rel:
                continue
            os.makedirs(os.path.join(PY_OUT, "chipmunk", rel), exist_ok=True)
            for f in files:
                if not f.endswith(".py"):
                    continue
                s, dst = os.path.join(d, f), os.path.join(PY_OUT, "chipmunk", rel, f)
                if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(s):
                    shutil.copyfile(s, dst)
    return PY_OUT


def available() -> bool:
    return all(os.path.exists(s) for s in SOURCES)


def build(force: bool = False) -> str | None:
    stage_triton(force)
    stage_python(force)
    if os.path.exists(LIB) and not force:
        return LIB
    if not available():
        return None
    import torch
    from torch.utils import cpp_extension as ce

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OUT, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")]
    import sysconfig
    inc.append(f"-I{sysconfig.get_paths()['include']}")
    inc += [f"-I{p}" for p in TK_INCLUDES]
    # -std=c++20 / -DKITTENS_HOPPER: the reference's own flags (setup.py:91-103)
    common = ["-std=c++20", "-O3", "-DKITTENS_HOPPER", "-Xcompiler", "-fPIC", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=1", *inc]
    cuflags = ["-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
               "--expt-extended-lambda", "-DNDEBUG", "--use_fast_math", "-DTORCH_COMPILE",   # setup.py:91-98
               # torch.utils.cpp_extension's defaults (the reference is built through CUDAExtension, setup.py:124-133):
               # without them at::Half comparisons in topk_indices.cu:20 are ambiguous with cuda_fp16.h's operators
               "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
               "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"]

    def cc(src: str) -> str:
        obj = os.path.join(OUT, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        flags = cuflags if src.endswith(".cu") else ["-x", "cu", *cuflags]
        cmd = [nvcc, *common, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj[:-2] + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr[-4000:]}")
        return obj

    with cf.ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(cc, [*SOURCES, SHIM]))
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
           "-ltorch", "-lcurand"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build("--force" in sys.argv))
