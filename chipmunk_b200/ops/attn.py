"""Attention operator wrappers (reference: src/chipmunk/ops/attn.py:42-169).

The reference pads Q (and the index rows) to a multiple of 192 and forces contiguity before
calling its Hopper kernels, then slices the result.  The sm_100a kernels take any Nq and
strided [B,H,N,128] views directly (rows past Nq are zero-filled on load and clipped on
store), so these wrappers do no copies; only the shapes they return follow the reference.
"""
from __future__ import annotations

import torch

QG = 192  # queries per index group


def _padded(n: int) -> int:
    return ((n + QG - 1) // QG) * QG


def csp_attn(q, k, v, indices, indices_counts):
    """delta = softmax(Q K[idx]^T / sqrt(d)) V[idx] per 192-row group, as a fresh tensor
    (reference ops/attn.py:134-169 -> csp_128_attn)."""
    return torch.ops.chipmunk.csp_128_attn(q, k, v, indices, indices_counts)


def csp_attn_add(q, k, v, cache, indices, indices_counts, o_scale: int = 1, out=None):
    """cache + o_scale * csp(q, k, v) in one kernel, out of place (B200 addition; replaces the
    `clone` + in-place `torch.ops.chipmunk.csp_attn` pair of reference modules/attn.py:165-190)."""
    from .. import torch_ops as _t
    return _t.csp_attn_add(q, k, v, cache, indices, indices_counts, o_scale, out)


def dense_attn(q, k, v):
    """Returns (o [B,H,N,128], l [B,H,pad192(N),1]) with l zero past N, the shape the reference
    hands back so that it can be fed to dense_colsum_attn (ops/attn.py:42-84)."""
    o, l = torch.ops.chipmunk.dense_attn(q, k, v)
    n, pn = q.shape[-2], _padded(q.shape[-2])
    if pn != n:
        lp = l.new_zeros((*l.shape[:2], pn, 1))
        lp[..., :n, :] = l
        l = lp
    return o, l


def dense_colsum_attn(q, k, v, p):
    """Returns (o, cs [B,H,ceil(Nk/192),Nk] bf16, l padded like dense_attn)
    (reference ops/attn.py:86-129).  `p` may be padded past N; the extra rows are ignored."""
    n, pn = q.shape[-2], _padded(q.shape[-2])
    assert p.shape[-2] in (n, pn), "p must have N or pad192(N) rows"
    o, cs, l = torch.ops.chipmunk.dense_colsum_attn(q, k, v, p)
    if pn != n:
        lp = l.new_zeros((*l.shape[:2], pn, 1))
        lp[..., :n, :] = l
        l = lp
        kseq = k.shape[-2]
        cs = cs[..., : (kseq + QG - 1) // QG, :kseq]
    return o, cs, l


__all__ = ["csp_attn", "csp_attn_add", "dense_attn", "dense_colsum_attn"]
