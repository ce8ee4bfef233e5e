"""Build the reference's binding TU on top of the chipmunk_b200 C ABI (INTEGRATION.md option 3).

    python bindings/build_binding.py            -> bindings/_build/cuda.so   (needs /root/reference: its csrc/chipmunk.cpp is
                                                   compiled UNMODIFIED, from where it lies; nothing of it is copied here)
    python bindings/build_binding.py --check    -> compile bindings/chipmunk_ops_b200.cpp only (no reference needed)

`cuda.so` is what the reference installs as `chipmunk/cuda.*.so`: importing it runs the reference's own
`TORCH_LIBRARY(chipmunk, m)` / `TORCH_LIBRARY_IMPL(chipmunk, CUDA, m)` initialisers (csrc/chipmunk.cpp:45-80), whose ten
`extern` operator functions (:27-43) are defined by bindings/chipmunk_ops_b200.cpp instead of csrc/attn, csrc/mlp and
csrc/indexed_io.  g++ only: there is no device code on this side of the C ABI.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
OURS = os.path.join(HERE, "chipmunk_ops_b200.cpp")
REF_TU = os.path.join(os.environ.get("CHIPMUNK_REFERENCE", "/root/reference"), "csrc", "chipmunk.cpp")
LIB = os.path.join(OUT, "cuda.so")


def _flags():
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{os.path.join(ROOT, 'include')}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    return ["-std=c++17", "-O2", "-fPIC", abi, "-DTORCH_API_INCLUDE_EXTENSION_H", *inc], os.path.join(os.path.dirname(torch.__file__), "lib")


def _cc(src: str, obj: str | None):
    cxx, _ = _flags()
    cmd = ["g++", *cxx, "-c", src, "-o", obj] if obj else ["g++", *cxx, "-fsyntax-only", src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed on {src}:\n{r.stderr[-4000:]}")
    return obj


def compile_check() -> None:
    _cc(OURS, None)


def build(force: bool = False) -> str | None:
    if not os.path.exists(REF_TU):
        return LIB if os.path.exists(LIB) else None
    os.makedirs(OUT, exist_ok=True)
    newest = max(os.path.getmtime(p) for p in (OURS, REF_TU, os.path.join(ROOT, "include", "chipmunk_b200.h")))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    objs = [os.path.join(OUT, "chipmunk_ops_b200.o"), os.path.join(OUT, "reference_chipmunk.o")]
    with cf.ThreadPoolExecutor(max_workers=2) as ex:
        list(ex.map(lambda a: _cc(*a), [(OURS, objs[0]), (REF_TU, objs[1])]))
    _, tlib = _flags()
    cmlib = os.path.join(ROOT, "chipmunk_b200")
    cmd = ["g++", "-shared", "-o", LIB, *objs, f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
           "-ltorch_python", f"-L{cmlib}", "-lchipmunk_b200", f"-Wl,-rpath,{tlib}", f"-Wl,-rpath,{cmlib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    if "--check" in sys.argv:
        compile_check()
        print("bindings/chipmunk_ops_b200.cpp compiles")
    else:
        print(build("--force" in sys.argv))
