"""ctypes binding of libchipmunk_b200.so (the C ABI declared in include/chipmunk_b200.h).

There is no CPU or PyTorch fallback: if the library is missing the import fails loudly, and
every call raises RuntimeError on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CHIPMUNK_B200_LIB: development override for same-box A/B timing of two builds (never a fallback: it must exist)
LIB_PATH = os.environ.get("CHIPMUNK_B200_LIB") or os.path.join(_HERE, "libchipmunk_b200.so")

ABI_VERSION = 2          # must equal CM_ABI_VERSION of include/chipmunk_b200.h (checked at load time and by the CPU tests)
CM_BF16, CM_F16, CM_F32 = 0, 1, 2
_DTYPE_TAG = {torch.bfloat16: CM_BF16, torch.float16: CM_F16, torch.float32: CM_F32}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: chipmunk_b200 has no fallback path. "
            "Build it with `python chipmunk_b200/build.py` (needs nvcc, no GPU required).")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    p64 = C.POINTER(C.c_int64)
    sigs = {
        "cm_abi_version": ([], i32),
        "cm_sm_count": ([], i32),
        "cm_strerror": ([i32], C.c_char_p),
        "cm_csp_attn": ([vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, p64, p64, p64, p64, i64, i32, i32, vp], i32),
        "cm_csp_attn_add": ([vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, p64, p64, p64, p64, p64, i64, i32, vp], i32),
        "cm_csp_attn_add_bcast": ([vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, p64, p64, p64, p64, p64, i64, i32, vp], i32),
        "cm_csp_attn_add_peers": ([vp, vp, vp, vp, vp, p64, i32, vp, vp, i32, i32, i32, i32, p64, p64, p64, p64, p64, i64, i32, vp], i32),
        "cm_dense_attn": ([vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, vp], i32),
        "cm_dense_attn_strided": ([vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, p64, p64, p64, p64, i64, vp], i32),
        "cm_csp_mlp_mm1": ([vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, i32, vp], i32),
        "cm_csp_mlp_mm2": ([vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, i32, vp], i32),
        "cm_csp_scatter_add": ([vp, vp, vp, vp, i32, i32, i64, vp], i32),
        "cm_mask_to_indices": ([vp, vp, vp, i64, i32, i32, i32, vp], i32),
        "cm_bitmask_to_indices": ([vp, vp, vp, i64, i32, i32, i32, vp], i32),
        "cm_select_columns": ([vp, i64, i64, i32, i32, f32, C.c_uint64, vp, i64, i32, vp, vp, vp, vp, i32, i32, vp], i32),
        "cm_topk_indices": ([vp, i32, vp, vp, i32, i32, i32, f32, i32, f32, vp], i32),
        "cm_copy_indices": ([vp, vp, i32, vp, vp, i32, i32, i32, i32, vp], i32),
        "cm_gather_rows": ([vp, vp, vp, i64, i64, i64, i64, vp], i32),
        "cm_bitpack": ([vp, vp, i64, vp], i32),
        "cm_bitunpack": ([vp, vp, i64, vp], i32),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(lib, name)      # AttributeError here = header and library disagree
        fn.argtypes = args
        fn.restype = res
    got = lib.cm_abi_version()
    if got != ABI_VERSION:     # a stale build: same symbol names, possibly different argument lists
        raise ImportError(f"{LIB_PATH} reports ABI version {got}, this package binds version {ABI_VERSION}: "
                          "rebuild it with `python chipmunk_b200/build.py --force`")
    return lib


lib = _load()
EXPORTS = ("cm_abi_version", "cm_sm_count", "cm_strerror", "cm_csp_attn", "cm_csp_attn_add", "cm_csp_attn_add_bcast", "cm_csp_attn_add_peers", "cm_dense_attn", "cm_dense_attn_strided",
           "cm_csp_mlp_mm1", "cm_csp_mlp_mm2", "cm_csp_scatter_add", "cm_mask_to_indices",
           "cm_bitmask_to_indices", "cm_select_columns", "cm_topk_indices", "cm_copy_indices", "cm_gather_rows", "cm_bitpack",
           "cm_bitunpack")


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib.cm_strerror(code).decode()
        raise RuntimeError(f"{what}: {msg} (code {code})")


def stream_ptr(device=None) -> int:
    """The caller's current CUDA stream (the reference used the legacy
    default stream for half of its kernels; here every kernel runs on the current stream)."""
    return torch.cuda.current_stream(device).cuda_stream


def strides3(t: torch.Tensor):
    """(batch, head, row) element strides of a [B,H,N,D] tensor as a C int64[3]."""
    return (C.c_int64 * 3)(t.stride(0), t.stride(1), t.stride(2))


def dtype_tag(dt: torch.dtype) -> int:
    try:
        return _DTYPE_TAG[dt]
    except KeyError:
        raise RuntimeError(f"Unsupported dtype {dt}") from None


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("chipmunk_b200 kernels need CUDA tensors; there is no CPU path")
