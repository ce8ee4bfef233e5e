"""Column-sparse MLP wrappers (reference: src/chipmunk/ops/mlp.py:7-92).

`run_e2e` is the call SparseDiffMlp makes.  The reference runs mm1, then a CUDA-graph that
overlaps a scatter-add kernel with a Triton mm2 on disjoint SM sets.  Here mm1's epilogue
already holds both gelu(new) - cache and the cache tile, so it writes the refreshed cache
itself (`update_pa`), and mm2 is a single tcgen05 GEMM: two launches, no graph, no Triton.
"""
from __future__ import annotations

import torch

from .. import torch_ops as _t

USE_FUSED_MLP_MATMUL_2 = True


def mm1(x, fc1w, sparse_act_packed, fc1b, sparse_act_T, indices, counts, scale_a=None, scale_b=None) -> None:
    assert x.dtype == torch.bfloat16 and sparse_act_packed.dtype == torch.bfloat16
    assert sparse_act_T.dtype == torch.bfloat16
    if fc1w.dtype != torch.bfloat16:
        raise ValueError(f"Unsupported dtype: {fc1w.dtype}")   # the fp8 preview path is out of scope
    torch.ops.chipmunk.csp_mlp_mm1(x, fc1w, sparse_act_packed, fc1b, sparse_act_T, indices, counts)


def mm2_fused(packed, unpacked_colmajor, indices, counts, sparse_act_packed, fc2wT, cached_out,
              num_sms_scatter_add: int) -> None:
    assert sparse_act_packed.dtype == fc2wT.dtype == cached_out.dtype == torch.bfloat16
    torch.ops.chipmunk.csp_mlp_mm2_and_scatter_add(
        packed.unsqueeze(0), unpacked_colmajor.unsqueeze(0), indices.unsqueeze(0), counts.unsqueeze(0),
        sparse_act_packed.unsqueeze(0), fc2wT.unsqueeze(0), cached_out.unsqueeze(0), num_sms_scatter_add, 0)


def mm2_unfused(sparse_act_packed, fc2wT, cached_out, unpacked_colmajor, indices, counts,
                num_sms_scatter_add: int) -> None:
    assert sparse_act_packed.dtype == fc2wT.dtype == cached_out.dtype == torch.bfloat16
    torch.ops.chipmunk.csp_scatter_add(sparse_act_packed.unsqueeze(0), unpacked_colmajor.unsqueeze(0),
                                       indices.unsqueeze(0), counts.unsqueeze(0), num_sms_scatter_add)
    _t.mlp_mm2(sparse_act_packed, fc2wT, cached_out, None, indices, counts, False)


@torch.compiler.disable
def run_e2e(x, fc1w, fc1b, fc2w_T, indices, counts, sparse_act_T, cached_out, num_sms_scatter_add: int = 0,
            mm1_scale_a=None, mm1_scale_b=None) -> None:
    """cached_out += (gelu(x W1[idx]^T + b1[idx]) - sparse_act_T[idx]) W2^T[idx];
    sparse_act_T[idx] <- gelu(...), both in place.  x [M,K1], fc1w [F,K1], fc2w_T [F,N]."""
    M, K1 = x.shape
    F, K1_ = fc1w.shape
    assert K1 == K1_, "K1 must match"
    F_, N = fc2w_T.shape
    assert F == F_, "K2 must match"
    if fc1w.dtype != torch.bfloat16:
        raise ValueError(f"Unsupported dtype: {fc1w.dtype}")
    packed = torch.empty((M, F), device=x.device, dtype=x.dtype)
    _t.mlp_mm1(x, fc1w, packed, fc1b, sparse_act_T, indices, counts, update_pa=True)
    _t.mlp_mm2(packed, fc2w_T, cached_out, None, indices, counts, False)


__all__ = ["mm1", "mm2_fused", "mm2_unfused", "run_e2e"]
