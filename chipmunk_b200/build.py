"""In-tree build of libchipmunk_b200.so (sm_100a only).

    python chipmunk_b200/build.py [--force] [--verbose]

Each `csrc/*.cu` is compiled to an object with nvcc (in parallel) and linked into
`chipmunk_b200/libchipmunk_b200.so`.  nvcc cross-compiles without a GPU, so this also runs in
the authoring container; the built library travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libchipmunk_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libchipmunk_b200.so cannot be built")
    return nvcc


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "chipmunk_b200.h"))
    return max(os.path.getmtime(h) for h in hs if os.path.exists(h))


EXTRA_FLAGS: list = []      # --debug-stages adds -DCM_DEBUG_STAGES (stage-isolation timing builds; never shipped)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, *EXTRA_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, src[:-3] + ".log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    hdr_t = _headers_mtime()
    todo, objs = [], []
    for s in srcs:
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        st = os.path.getmtime(os.path.join(CSRC, s))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(st, hdr_t):
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--debug-stages", action="store_true",
                    help="compile the CM_DEBUG_FLAGS stage-isolation switches in (timing experiments; results are wrong when set)")
    a = ap.parse_args()
    if a.debug_stages:
        EXTRA_FLAGS.append("-DCM_DEBUG_STAGES")
        a.force = True
    print(build(a.force, a.verbose))
