"""`torch.ops.chipmunk.*` — the reference's dispatcher surface (csrc/chipmunk.cpp:45-80),
registered from Python with torch.library and implemented by the C ABI.

The ten schemas are the reference's (chipmunk.cpp:47-60) with one deliberate difference:
`csp_attn` annotates `o` as mutated (`Tensor(o!)`), which the reference omits although its
kernel writes `o` in place.  Argument order, names, dtypes and the TORCH_CHECK
error behaviour (RuntimeError) are kept.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from ._lib import check, lib, require_cuda, stream_ptr, strides3

QG = 192      # query rows per index group
MLP_BM = 128  # token rows per MLP index group

_SCHEMAS = {
    "csp_mlp_mm1": "(Tensor a, Tensor b_colmajor, Tensor(c!) c, Tensor bias, Tensor pa_cache_colmajor, Tensor indices, Tensor indices_counts) -> ()",
    "csp_mlp_mm2_and_scatter_add": "(Tensor packed, Tensor(unpacked_colmajor!) unpacked_colmajor, Tensor sp_inds, Tensor sp_counts, Tensor mma_a, Tensor mma_b, Tensor(mma_c!) mma_c, int num_sms_scatter_add, int matmul_kernel) -> ()",
    "csp_attn": "(Tensor q, Tensor k, Tensor v, Tensor(o!) o, Tensor indices, Tensor indices_counts, int o_scale) -> ()",
    "csp_128_attn": "(Tensor q, Tensor k, Tensor v, Tensor indices, Tensor indices_counts) -> Tensor",
    "dense_attn": "(Tensor q, Tensor k, Tensor v) -> Tensor[]",
    "dense_colsum_attn": "(Tensor q, Tensor k, Tensor v, Tensor p) -> Tensor[]",
    "copy_indices": "(Tensor bmfc1, Tensor(bm_mid_cache!) bm_mid_cache, Tensor sp_inds, Tensor sp_counts) -> ()",
    "topk_indices": "(Tensor activation, Tensor(indices!) indices, Tensor(counts!) counts, float sparsity_amount, int multiple_of, float random_amount) -> ()",
    "csp_scatter_add": "(Tensor packed, Tensor(unpacked_colmajor!) unpacked_colmajor, Tensor sp_inds, Tensor sp_counts, int num_sms) -> ()",
    "mask_to_indices": "(Tensor mask, int multiple_of, int pad_to_multiple_of) -> Tensor[]",
}


def _chk(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


# ------------------------------------------------------------------------------- attention
def _attn_common_checks(q, k, v, indices, counts, what):
    require_cuda(q, k, v, indices, counts)
    _chk(q.dim() == 4 and k.dim() == 4 and v.dim() == 4, f"{what}: q, k, v must be [B,H,N,D]")
    B, H, Nq, D = q.shape
    _chk(D == 128, "Head dimension must be 128")
    _chk(q.dtype == k.dtype == v.dtype == torch.bfloat16, f"{what}: q, k, v must be bfloat16")
    _chk(k.shape[0] == B and v.shape[0] == B, "K/V batch dimension - idx 0 - must match for all inputs")
    _chk(k.shape[1] == H and v.shape[1] == H, "QO heads must be equal to KV heads")
    _chk(k.shape[2] == v.shape[2], "V sequence length dimension - idx 2 - must match for all inputs")
    _chk(k.shape[3] == D and v.shape[3] == D, "K/V head dimension - idx 3 - must match")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t.stride(3) == 1, f"{what}: {n}.stride(3) must be 1")
        _chk(all(s % 8 == 0 for s in t.stride()[:3]) and t.data_ptr() % 16 == 0,
             f"{what}: {n} must be 16-byte aligned in every stride")
    G = (Nq + QG - 1) // QG
    _chk(indices.dim() == 4, "Indices must be a 4D tensor")
    _chk(counts.dim() == 3, "Indices counts must be a 3D tensor")
    _chk(indices.dtype == torch.int32, "Indices must be a 32-bit integer tensor")
    _chk(counts.dtype == torch.int32, "Indices counts must be a 32-bit integer tensor")
    _chk(indices.is_contiguous(), "Indices must be contiguous")
    _chk(counts.is_contiguous(), "Indices counts must be contiguous")
    _chk(tuple(indices.shape[:3]) == (B, H, G), "Indices [batch, head, query group] dimensions must match q")
    _chk(tuple(counts.shape) == (B, H, G), "Indices counts [batch, head, query group] dimensions must match q")
    return B, H, Nq, k.shape[2]


def _launch_csp_attn(q, k, v, o, indices, counts, o_scale, accumulate):
    B, H, Nq, Nk = _attn_common_checks(q, k, v, indices, counts, "csp_attn")
    _chk(o.shape == q.shape and o.dtype == torch.bfloat16 and o.stride(3) == 1, "O must match Q")
    _chk(all(s % 8 == 0 for s in o.stride()[:3]) and o.data_ptr() % 16 == 0, "O must be 16-byte aligned")
    if B * H * Nq == 0:
        return
    with torch.cuda.device(q.device):
        check(lib.cm_csp_attn(_ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(indices), _ptr(counts),
                              B, H, Nq, Nk, strides3(q), strides3(k), strides3(v), strides3(o),
                              indices.shape[3], int(o_scale), int(accumulate), stream_ptr(q.device)),
              "csp_attn")


def csp_attn(q, k, v, o, indices, indices_counts, o_scale):
    _chk(o_scale in (1, -1), "o_scale must be 1 or -1")
    _launch_csp_attn(q, k, v, o, indices, indices_counts, o_scale, 1)


def csp_attn_add(q, k, v, cache, indices, indices_counts, o_scale: int = 1, out=None, multicast_delta: int = 0, peer_deltas=None):
    """o = bf16(cache + o_scale * delta) as a fresh tensor, or into `out` (any [B,H,N,128] bf16 view with
    16-byte-aligned strides, e.g. this rank's slice of an all-gather buffer).  B200 addition: the sparse
    step's `o = cache.clone(); csp_attn(..., o, ...)` in one pass; `cache` is left untouched."""
    _chk(o_scale in (1, -1), "o_scale must be 1 or -1")
    B, H, Nq, Nk = _attn_common_checks(q, k, v, indices, indices_counts, "csp_attn_add")
    _chk(cache.shape == q.shape and cache.dtype == torch.bfloat16 and cache.stride(3) == 1, "cache must match Q")
    _chk(all(s % 8 == 0 for s in cache.stride()[:3]) and cache.data_ptr() % 16 == 0, "cache must be 16-byte aligned")
    require_cuda(cache)
    if out is None:
        o = torch.empty(q.shape, dtype=q.dtype, device=q.device)
    else:
        o = out
        require_cuda(o)
        _chk(o.shape == q.shape and o.dtype == torch.bfloat16 and o.stride(3) == 1, "out must match Q")
        _chk(all(s % 8 == 0 for s in o.stride()[:3]) and o.data_ptr() % 16 == 0, "out must be 16-byte aligned")
        _chk(o.data_ptr() != cache.data_ptr(), "out must not alias cache (use csp_attn for the in-place form)")
    if B * H * Nq == 0:
        return o
    with torch.cuda.device(q.device):
        if peer_deltas is not None:
            # `out` is this rank's slice of a symmetric buffer; rows are stored into every peer's copy over NVLink (parallel.py)
            _chk(out is not None and not multicast_delta, "peer_deltas needs out= (a slice of a symmetric-memory buffer) and no multicast_delta")
            arr = (C.c_int64 * len(peer_deltas))(*[int(d) for d in peer_deltas])
            check(lib.cm_csp_attn_add_peers(_ptr(q), _ptr(k), _ptr(v), _ptr(cache), _ptr(o), arr, len(peer_deltas),
                                            _ptr(indices), _ptr(indices_counts), B, H, Nq, Nk, strides3(q), strides3(k),
                                            strides3(v), strides3(cache), strides3(o), indices.shape[3], int(o_scale),
                                            stream_ptr(q.device)), "csp_attn_add_peers")
        elif multicast_delta:
            # `out` is this rank's slice of a symmetric buffer; rows go to its NVLS multicast alias (parallel.py)
            _chk(out is not None, "multicast_delta needs out= (a slice of a symmetric-memory buffer)")
            check(lib.cm_csp_attn_add_bcast(_ptr(q), _ptr(k), _ptr(v), _ptr(cache), _ptr(o), int(multicast_delta),
                                            _ptr(indices), _ptr(indices_counts), B, H, Nq, Nk, strides3(q), strides3(k),
                                            strides3(v), strides3(cache), strides3(o), indices.shape[3], int(o_scale),
                                            stream_ptr(q.device)), "csp_attn_add_bcast")
        else:
            check(lib.cm_csp_attn_add(_ptr(q), _ptr(k), _ptr(v), _ptr(cache), _ptr(o), _ptr(indices), _ptr(indices_counts),
                                      B, H, Nq, Nk, strides3(q), strides3(k), strides3(v), strides3(cache), strides3(o),
                                      indices.shape[3], int(o_scale), stream_ptr(q.device)), "csp_attn_add")
    return o


def csp_128_attn(q, k, v, indices, indices_counts):
    o = torch.empty(q.shape, dtype=q.dtype, device=q.device)
    _launch_csp_attn(q, k, v, o, indices, indices_counts, 1, 0)
    return o


def _launch_dense(q, k, v, p):
    require_cuda(q, k, v)
    _chk(q.dim() == 4 and q.shape[3] == 128, "Head dimension must be 128")
    _chk(q.dtype == k.dtype == v.dtype == torch.bfloat16, "dense_attn: q, k, v must be bfloat16")
    _chk(k.shape == v.shape and k.shape[:2] == q.shape[:2], "dense_attn: K/V shapes must match Q's batch and heads")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t.stride(3) == 1, f"dense_attn: {n}.stride(3) must be 1")
        _chk(all(s % 8 == 0 for s in t.stride()[:3]) and t.data_ptr() % 16 == 0,
             f"dense_attn: {n} must be 16-byte aligned in every stride")
    B, H, Nq, _ = q.shape
    Nk = k.shape[2]
    G = (Nq + QG - 1) // QG
    o = torch.empty(q.shape, dtype=q.dtype, device=q.device)
    l = torch.empty(B, H, Nq, 1, dtype=torch.float32, device=q.device)
    cs = None
    cs_stride = (Nk + 7) // 8 * 8
    if p is not None:
        _chk(p.dtype == torch.float32 and p.numel() >= B * H * Nq, "dense_colsum_attn: p must be fp32 [B,H,N,1]")
        _chk(tuple(p.shape[:2]) == (B, H), "dense_colsum_attn: p batch/head must match q")
        p = p.reshape(B, H, -1)[:, :, :Nq].contiguous()
        cs = torch.empty(B, H, G, cs_stride, dtype=torch.bfloat16, device=q.device)
    with torch.cuda.device(q.device):
        check(lib.cm_dense_attn_strided(_ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(l),
                                        _ptr(cs) if cs is not None else None, _ptr(p) if p is not None else None,
                                        B, H, Nq, Nk, strides3(q), strides3(k), strides3(v), strides3(o),
                                        cs_stride, stream_ptr(q.device)), "dense_attn")
    if cs is not None and cs_stride != Nk:
        cs = cs[..., :Nk]
    return o, cs, l


def dense_attn(q, k, v) -> List[torch.Tensor]:
    o, _, l = _launch_dense(q, k, v, None)
    return [o, l]


def dense_colsum_attn(q, k, v, p) -> List[torch.Tensor]:
    o, cs, l = _launch_dense(q, k, v, p)
    return [o, cs, l]


# ------------------------------------------------------------------------------------- MLP
def _mlp_index_checks(indices, counts, M, F, what):
    _chk(indices.dtype == torch.int32 and counts.dtype == torch.int32, f"{what}: indices/counts must be int32")
    _chk(indices.is_contiguous() and counts.is_contiguous(), f"{what}: indices/counts must be contiguous")
    _chk(M % MLP_BM == 0, f"{what}: M must be a multiple of {MLP_BM}")
    _chk(indices.numel() == (M // MLP_BM) * indices.shape[-1] and counts.numel() == M // MLP_BM,
         f"{what}: indices must be [M/128, F] and counts [M/128]")


def mlp_mm1(a, w1, c, bias, pa_T, indices, counts, update_pa: bool = False):
    """csp_mlp_mm1 with the optional fused cache update (see include/chipmunk_b200.h)."""
    require_cuda(a, w1, c, bias, pa_T, indices, counts)
    for t in (a, w1, c, bias, pa_T):
        _chk(t.dtype == torch.bfloat16, "csp_mlp_mm1: a, b, c, bias, pa_cache must be bfloat16")
        _chk(t.is_contiguous(), "csp_mlp_mm1: tensors must be contiguous")
    _chk(a.dim() == 2 and w1.dim() == 2 and c.dim() == 2, "csp_mlp_mm1: a [M,K], b_colmajor [F,K], c [M,F]")
    M, K = a.shape
    F = w1.shape[0]
    _chk(w1.shape[1] == K, "csp_mlp_mm1: K must match")
    _chk(tuple(c.shape) == (M, F), "csp_mlp_mm1: c must be [M,F]")
    _chk(bias.numel() == F, "csp_mlp_mm1: bias must be [F]")
    _chk(tuple(pa_T.shape) == (F, M), "csp_mlp_mm1: pa_cache_colmajor must be [F,M]")
    _chk(K % 64 == 0, "csp_mlp_mm1: K must be a multiple of 64")
    _mlp_index_checks(indices, counts, M, F, "csp_mlp_mm1")
    with torch.cuda.device(a.device):
        check(lib.cm_csp_mlp_mm1(_ptr(a), _ptr(w1), _ptr(c), _ptr(bias), _ptr(pa_T), _ptr(indices),
                                 _ptr(counts), M, K, F, indices.shape[-1], int(update_pa),
                                 stream_ptr(a.device)), "csp_mlp_mm1")


def csp_mlp_mm1(a, b_colmajor, c, bias, pa_cache_colmajor, indices, indices_counts):
    mlp_mm1(a, b_colmajor, c, bias, pa_cache_colmajor, indices, indices_counts, False)


def mlp_mm2(packed, w2_T, out, pa_T, indices, counts, do_scatter: bool):
    require_cuda(packed, w2_T, out, indices, counts)
    for t in (packed, w2_T, out):
        _chk(t.dtype == torch.bfloat16 and t.is_contiguous(), "csp_mlp_mm2: tensors must be contiguous bfloat16")
    M, F = packed.shape[-2:]
    N = w2_T.shape[-1]
    _chk(w2_T.shape[-2] == F and tuple(out.shape[-2:]) == (M, N), "csp_mlp_mm2: shapes must be packed [M,F], w2_T [F,N], out [M,N]")
    _chk(N % 256 == 0, "csp_mlp_mm2: N must be a multiple of 256")
    if do_scatter:
        _chk(pa_T is not None and pa_T.dtype == torch.bfloat16 and pa_T.is_contiguous()
             and tuple(pa_T.shape[-2:]) == (F, M), "csp_mlp_mm2: unpacked_colmajor must be contiguous bf16 [F,M]")
    _mlp_index_checks(indices, counts, M, F, "csp_mlp_mm2")
    with torch.cuda.device(packed.device):
        check(lib.cm_csp_mlp_mm2(_ptr(packed), _ptr(w2_T), _ptr(out), _ptr(pa_T) if do_scatter else None,
                                 _ptr(indices), _ptr(counts), M, F, N, indices.shape[-1],
                                 int(do_scatter), stream_ptr(packed.device)), "csp_mlp_mm2")


def csp_mlp_mm2_and_scatter_add(packed, unpacked_colmajor, sp_inds, sp_counts, mma_a, mma_b, mma_c,
                                num_sms_scatter_add, matmul_kernel):
    # `matmul_kernel` was a raw CUfunction of the reference's Triton kernel and
    # `num_sms_scatter_add` its SM split (csp_mlp_mm2_and_scatter_add.cu:181-256); both are
    # accepted and ignored: one sm_100a kernel does the GEMM and the scatter.
    _chk(packed.dim() == 3 and packed.shape[0] == 1, "csp_mlp_mm2_and_scatter_add: batch must be 1")
    _chk(mma_a.data_ptr() == packed.data_ptr(), "csp_mlp_mm2_and_scatter_add: mma_a must alias packed")
    mlp_mm2(mma_a[0], mma_b[0], mma_c[0], unpacked_colmajor[0], sp_inds[0], sp_counts[0], True)


def csp_scatter_add(packed, unpacked_colmajor, sp_inds, sp_counts, num_sms):
    require_cuda(packed, unpacked_colmajor, sp_inds, sp_counts)
    _chk(packed.dim() == 3 and packed.shape[0] == 1, "csp_scatter_add: batch must be 1")
    _chk(packed.dtype == torch.bfloat16 and unpacked_colmajor.dtype == torch.bfloat16, "csp_scatter_add: bf16 only")
    _chk(packed.is_contiguous() and unpacked_colmajor.is_contiguous(), "csp_scatter_add: tensors must be contiguous")
    M, F = packed.shape[-2:]
    _chk(tuple(unpacked_colmajor.shape[-2:]) == (F, M), "csp_scatter_add: unpacked_colmajor must be [1,F,M]")
    _mlp_index_checks(sp_inds, sp_counts, M, F, "csp_scatter_add")
    with torch.cuda.device(packed.device):
        check(lib.cm_csp_scatter_add(_ptr(packed), _ptr(unpacked_colmajor), _ptr(sp_inds), _ptr(sp_counts),
                                     M, F, sp_inds.shape[-1], stream_ptr(packed.device)), "csp_scatter_add")


# ------------------------------------------------------------------------------ indexed IO
def copy_indices(bmfc1, bm_mid_cache, sp_inds, sp_counts):
    require_cuda(bmfc1, bm_mid_cache, sp_inds, sp_counts)
    _chk(sp_inds.dtype == torch.int32, "sp_inds must be int32")
    _chk(sp_counts.dtype == torch.int32, "sp_counts must be int32")
    _chk(bmfc1.dtype == bm_mid_cache.dtype and bmfc1.shape == bm_mid_cache.shape, "copy_indices: src/dst must match")
    _chk(bmfc1.dtype in (torch.bfloat16, torch.float16, torch.float32), "Unsupported tensor type")
    _chk(bmfc1.is_contiguous() and bm_mid_cache.is_contiguous() and sp_inds.is_contiguous()
         and sp_counts.is_contiguous(), "copy_indices: tensors must be contiguous")
    B, M, F = bmfc1.shape[0], sp_inds.shape[1], sp_inds.shape[2]
    _chk(bmfc1.shape[2] == F and bm_mid_cache.shape[1] % M == 0, "copy_indices: shapes must be [B,M*R,F] / [B,M,F]")
    R = bm_mid_cache.shape[1] // M
    with torch.cuda.device(bmfc1.device):
        check(lib.cm_copy_indices(_ptr(bmfc1), _ptr(bm_mid_cache), bmfc1.element_size(), _ptr(sp_inds),
                                  _ptr(sp_counts), B, M, R, F, stream_ptr(bmfc1.device)), "copy_indices")


def topk_indices(activation, indices, counts, sparsity_amount, multiple_of, random_amount):
    require_cuda(activation, indices, counts)
    _chk(activation.dim() == 3, "activation must be 3-dimensional [batch, rows, cols]")
    _chk(indices.dim() == 3, "indices must be 3-dimensional [batch, rows, cols]")
    _chk(counts.dim() == 2, "counts must be 2-dimensional [batch, rows]")
    _chk(0 <= sparsity_amount <= 1, "sparsity_amount must be between 0 and 1")
    _chk(indices.dtype == torch.int32 and counts.dtype == torch.int32, "indices/counts must be int32")
    _chk(activation.is_contiguous() and indices.is_contiguous() and counts.is_contiguous(),
         "topk_indices: tensors must be contiguous")
    _chk(indices.shape == activation.shape and tuple(counts.shape) == tuple(activation.shape[:2]),
         "topk_indices: indices/counts shapes must match activation")
    _chk(activation.dtype in (torch.bfloat16, torch.float16, torch.float32), "Unsupported dtype for activation tensor")
    B, R, Cn = activation.shape
    with torch.cuda.device(activation.device):
        check(lib.cm_topk_indices(_ptr(activation), _lib.dtype_tag(activation.dtype), _ptr(indices), _ptr(counts),
                                  B, R, Cn, float(sparsity_amount), int(multiple_of), float(random_amount),
                                  stream_ptr(activation.device)), "topk_indices")


def mask_to_indices(mask, multiple_of, pad_to_multiple_of=192) -> List[torch.Tensor]:
    require_cuda(mask)
    _chk(mask.dim() == 4, "mask must be 4-dimensional [b, h, m, n]")
    _chk(mask.dtype == torch.bool, "mask must be bool type")
    mask = mask.contiguous()
    b, h, m, n = mask.shape
    pad_n = ((n + pad_to_multiple_of - 1) // pad_to_multiple_of) * pad_to_multiple_of
    indices = torch.empty(b, h, m, pad_n, dtype=torch.int32, device=mask.device)
    counts = torch.empty(b, h, m, dtype=torch.int32, device=mask.device)
    with torch.cuda.device(mask.device):
        check(lib.cm_mask_to_indices(_ptr(mask), _ptr(indices), _ptr(counts), b * h * m, n, pad_n,
                                     int(multiple_of), stream_ptr(mask.device)), "mask_to_indices")
    return [indices, counts]


def bitmask_to_indices(packed, mask_shape, multiple_of, pad_to_multiple_of=192) -> List[torch.Tensor]:
    """Fused bitunpack + mask_to_indices (not a reference op; replaces the pair of calls in
    src/chipmunk/modules/attn.py:173-176)."""
    require_cuda(packed)
    _chk(packed.dtype == torch.uint8 and packed.is_contiguous(), "packed must be contiguous uint8")
    b, h, m, n = mask_shape
    _chk(packed.numel() * 8 >= b * h * m * n, "packed is too short for mask_shape")
    pad_n = ((n + pad_to_multiple_of - 1) // pad_to_multiple_of) * pad_to_multiple_of
    indices = torch.empty(b, h, m, pad_n, dtype=torch.int32, device=packed.device)
    counts = torch.empty(b, h, m, dtype=torch.int32, device=packed.device)
    with torch.cuda.device(packed.device):
        check(lib.cm_bitmask_to_indices(_ptr(packed), _ptr(indices), _ptr(counts), b * h * m, n, pad_n,
                                        int(multiple_of), stream_ptr(packed.device)), "bitmask_to_indices")
    return [indices, counts]


def pack_rows_to_words(mask2d: torch.Tensor) -> torch.Tensor:
    """bool [R, n] -> int32 words [R, ceil(n/32)], bit c%32 of word c/32 = column c (each row padded to whole words):
    the layout select_columns takes its static mask in.  One-time preprocessing (initialize_static_mask)."""
    require_cuda(mask2d)
    _chk(mask2d.dim() == 2 and mask2d.dtype == torch.bool, "pack_rows_to_words: bool [R, n] expected")
    R, n = mask2d.shape
    W = (n + 31) // 32
    padded = torch.zeros(R, W * 32, dtype=torch.bool, device=mask2d.device)
    padded[:, :n] = mask2d
    packed, _ = bitpack(padded)
    return packed.view(torch.int32).view(R, W)


def select_columns(cs, k: int, multiple_of: int, random_prob: float = 0.0, static_words=None, group_is_sparse=None,
                   seed=None, pad_to_multiple_of: int = 192, want_packed: bool = True, want_indices: bool = True):
    """Top-k (+ random, + static mask) column selection of one full step, in one kernel (cm_select_columns).
    cs [B,H,G,n] bf16 (last-dim stride 1, rows equally strided).  Returns (packed uint8 | None, mask_shape,
    indices [B,H,G,pad_n] | None, counts [B,H,G] | None).  `seed` defaults to a draw from torch's CUDA generator, so
    torch.manual_seed() governs the random columns as it governs the reference's torch.randint."""
    require_cuda(cs)
    _chk(cs.dim() == 4 and cs.dtype == torch.bfloat16, "select_columns: cs must be bf16 [B,H,G,n]")
    _chk(cs.stride(3) == 1, "select_columns: cs.stride(3) must be 1")
    B, H, G, n = cs.shape
    rows = B * H * G
    rs = cs.stride(2)
    _chk(rows <= 1 or (cs.stride(1) == G * rs and cs.stride(0) == H * G * rs) or cs.is_contiguous(),
         "select_columns: cs rows must be equally strided")
    _chk(0 <= k, "select_columns: k must be >= 0")
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,), device="cpu").item()) if random_prob > 0 else 0
    packed = indices = counts = None
    if want_packed:
        nbits = rows * n
        buf = torch.empty(4 * ((nbits + 31) // 32), dtype=torch.uint8, device=cs.device)
        packed = buf[: (nbits + 7) // 8]
    if want_indices:
        pad_n = ((n + pad_to_multiple_of - 1) // pad_to_multiple_of) * pad_to_multiple_of
        indices = torch.empty(B, H, G, pad_n, dtype=torch.int32, device=cs.device)
        counts = torch.empty(B, H, G, dtype=torch.int32, device=cs.device)
    else:
        pad_n = n
    srows = 0
    sstride = 0
    if static_words is not None:
        require_cuda(static_words)
        _chk(static_words.dtype == torch.int32 and static_words.dim() == 2 and static_words.is_contiguous(),
             "select_columns: static_words must be contiguous int32 [G, ceil(n/32)]")
        _chk(static_words.shape[1] >= (n + 31) // 32, "select_columns: static_words rows are too short")
        srows, sstride = static_words.shape
    if group_is_sparse is not None:
        require_cuda(group_is_sparse)
        group_is_sparse = group_is_sparse.reshape(-1)
        _chk(group_is_sparse.dtype in (torch.bool, torch.uint8) and group_is_sparse.is_contiguous(),
             "select_columns: group_is_sparse must be bool [G]")
        _chk(srows in (0, group_is_sparse.numel()), "select_columns: static_words and group_is_sparse must have the same rows")
        srows = group_is_sparse.numel()
    if srows:
        _chk(G % srows == 0 or srows >= G, "select_columns: static rows must cover the query groups")
        if srows > G:      # the reference slices [..., :qg, :n]
            srows = G
    with torch.cuda.device(cs.device):
        check(lib.cm_select_columns(_ptr(cs), rs, rows, n, int(k), float(random_prob), int(seed) & (2 ** 64 - 1),
                                    _ptr(static_words) if static_words is not None else None, sstride, srows,
                                    _ptr(group_is_sparse) if group_is_sparse is not None else None,
                                    _ptr(packed) if packed is not None else None,
                                    _ptr(indices) if indices is not None else None,
                                    _ptr(counts) if counts is not None else None,
                                    pad_n, int(multiple_of), stream_ptr(cs.device)), "select_columns")
    return packed, torch.Size((B, H, G, n)), indices, counts


def gather_rows(x: torch.Tensor, dim: int, perm: torch.Tensor) -> torch.Tensor:
    """x.index_select(dim, perm) for a CUDA tensor, as ONE gather kernel (cm_gather_rows): the token reorderings of
    chipmunk.ops.patch / chipmunk.ops.voxel.  `perm` int32 on the same device."""
    require_cuda(x, perm)
    _chk(perm.dtype == torch.int32 and perm.is_contiguous() and perm.dim() == 1, "gather_rows: perm must be contiguous int32 [n]")
    dim = dim % x.dim()
    x = x.contiguous()
    outer = 1
    for d in x.shape[:dim]:
        outer *= int(d)
    inner = x.element_size()
    for d in x.shape[dim + 1:]:
        inner *= int(d)
    n_src, n_dst = int(x.shape[dim]), int(perm.numel())
    out = torch.empty((*x.shape[:dim], n_dst, *x.shape[dim + 1:]), dtype=x.dtype, device=x.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        check(lib.cm_gather_rows(_ptr(x), _ptr(out), _ptr(perm), outer, n_src, n_dst, inner, stream_ptr(x.device)), "gather_rows")
    return out


def bitpack(mask):
    require_cuda(mask)
    _chk(mask.dtype == torch.bool, "mask must be bool type")
    shape = mask.shape
    flat = mask.contiguous().view(-1)
    n = flat.numel()
    packed = torch.empty((n + 7) // 8, dtype=torch.uint8, device=mask.device)
    with torch.cuda.device(mask.device):
        check(lib.cm_bitpack(_ptr(flat), _ptr(packed), n, stream_ptr(mask.device)), "bitpack")
    return packed, shape


def bitunpack(packed, original_shape):
    require_cuda(packed)
    _chk(packed.dtype == torch.uint8 and packed.is_contiguous(), "packed must be contiguous uint8")
    n = 1
    for d in original_shape:
        n *= int(d)
    _chk(packed.numel() * 8 >= n, "packed is too short for original_shape")
    mask = torch.empty(n, dtype=torch.bool, device=packed.device)
    with torch.cuda.device(packed.device):
        check(lib.cm_bitunpack(_ptr(packed), _ptr(mask), n, stream_ptr(packed.device)), "bitunpack")
    return mask.view(*original_shape)


# ---------------------------------------------------------------------------- registration
_IMPLS = {
    "csp_mlp_mm1": csp_mlp_mm1,
    "csp_mlp_mm2_and_scatter_add": csp_mlp_mm2_and_scatter_add,
    "csp_attn": csp_attn,
    "csp_128_attn": csp_128_attn,
    "dense_attn": dense_attn,
    "dense_colsum_attn": dense_colsum_attn,
    "copy_indices": copy_indices,
    "topk_indices": topk_indices,
    "csp_scatter_add": csp_scatter_add,
    "mask_to_indices": mask_to_indices,
}

_lib_def = None
_lib_impl = None


def _fake_impls():
    def g(n):
        return (n + QG - 1) // QG

    def f_csp_128(q, k, v, indices, indices_counts):
        return torch.empty_like(q)

    def f_dense(q, k, v):
        return [torch.empty_like(q), q.new_empty((*q.shape[:3], 1), dtype=torch.float32)]

    def f_colsum(q, k, v, p):
        return [torch.empty_like(q),
                q.new_empty((q.shape[0], q.shape[1], g(q.shape[2]), k.shape[2])),
                q.new_empty((*q.shape[:3], 1), dtype=torch.float32)]

    def f_m2i(mask, multiple_of, pad_to_multiple_of):
        b, h, m, n = mask.shape
        pad_n = ((n + pad_to_multiple_of - 1) // pad_to_multiple_of) * pad_to_multiple_of
        return [mask.new_empty((b, h, m, pad_n), dtype=torch.int32), mask.new_empty((b, h, m), dtype=torch.int32)]

    def f_none(*a, **k):
        return None

    fakes = {n: f_none for n in _IMPLS}
    fakes.update(csp_128_attn=f_csp_128, dense_attn=f_dense, dense_colsum_attn=f_colsum, mask_to_indices=f_m2i)
    return fakes


def register() -> None:
    """Define the `chipmunk` operator library once per process and attach the CUDA kernels."""
    global _lib_def, _lib_impl
    if _lib_def is not None:
        return
    _lib_def = torch.library.Library("chipmunk", "DEF")
    for name, schema in _SCHEMAS.items():
        _lib_def.define(name + schema)
    _lib_impl = torch.library.Library("chipmunk", "IMPL", "CUDA")
    for name, fn in _IMPLS.items():
        _lib_impl.impl(name, fn)
    for name, fn in _fake_impls().items():
        try:
            torch.library.register_fake(f"chipmunk::{name}", fn, lib=_lib_def)
        except Exception:        # older torch: no register_fake(lib=...); eager path unaffected
            pass
