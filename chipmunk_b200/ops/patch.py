"""2-level 2-D token patching (reference: src/chipmunk/ops/patch.py:7-80).

The reference chains four einops rearranges per call.  The transform is a fixed permutation of the
h*w token positions, so here it is computed ONCE per (h, w, chunk sizes, device) as an index vector
and applied with a single gather, as is its inverse: CUDA tensors go through the library's gather
kernel (cm_gather_rows), CPU tensors (the reference also calls these on host-side tables at set-up time) through
torch.index_select.
"""
from __future__ import annotations

from functools import lru_cache

import torch

from ..util import GLOBAL_CONFIG


def _chunks():
    cfg = GLOBAL_CONFIG["patchify"]
    return int(cfg["chunk_size_1"]), int(cfg["chunk_size_2"])


@lru_cache(maxsize=32)
def _perm(h: int, w: int, c1: int, c2: int, device: str):
    """perm[i] = raster position (row*w + col) of the token that lands at patched position i.
    Patched order: c1 x c1 patches in raster order; inside each, c2 x c2 sub-patches in raster order;
    inside each sub-patch, raster order."""
    assert h % c1 == 0 and w % c1 == 0, "Height and width must be divisible by chunk_size_1."
    assert c1 % c2 == 0, "chunk_size_1 must be divisible by chunk_size_2."
    dev = torch.device(device)
    r = torch.arange(h, device=dev)[:, None].expand(h, w)
    c = torch.arange(w, device=dev)[None, :].expand(h, w)
    n2 = c1 // c2
    key = ((((r // c1) * (w // c1) + (c // c1)) * n2 + (r % c1) // c2) * n2 + (c % c1) // c2) * (c2 * c2) \
        + (r % c2) * c2 + (c % c2)
    perm = torch.empty(h * w, dtype=torch.long, device=dev)
    perm[key.reshape(-1)] = torch.arange(h * w, device=dev)
    inv = key.reshape(-1).clone()
    if dev.type == "cuda":
        return perm.to(torch.int32), inv.to(torch.int32)
    return perm, inv


def _gather(x: torch.Tensor, dim: int, perm: torch.Tensor) -> torch.Tensor:
    if x.is_cuda:
        from .. import torch_ops as _t
        return _t.gather_rows(x, dim, perm)
    return x.index_select(dim, perm)


def patchify(x: torch.Tensor) -> torch.Tensor:
    """[b, h, w] -> [b, h*w] in patched order."""
    assert x.ndim == 3, "Input tensor must have 3 dimensions (b, h, w)."
    b, h, w = x.shape
    c1, c2 = _chunks()
    perm, _ = _perm(h, w, c1, c2, str(x.device))
    return _gather(x.reshape(b, h * w), 1, perm)


def unpatchify(x_chunk_flat: torch.Tensor, original_shape) -> torch.Tensor:
    """[b, h*w] in patched order -> [b, h, w]."""
    b, h, w = original_shape
    c1, c2 = _chunks()
    _, inv = _perm(h, w, c1, c2, str(x_chunk_flat.device))
    return _gather(x_chunk_flat, 1, inv).reshape(x_chunk_flat.shape[0], h, w)


def patchify_rope(x_shape, pe: torch.Tensor, width_rope: int, height_rope: int) -> torch.Tensor:
    """Reorder the image-token part of the rotary table pe [a, b, tokens, d, e, 2] in place
    (reference patch.py:63-80)."""
    img_tokens = x_shape[1]
    c1, c2 = _chunks()
    perm, _ = _perm(height_rope, width_rope, c1, c2, str(pe.device))
    tail = pe[:, :, -img_tokens:]
    pe[:, :, -img_tokens:] = _gather(tail, 2, perm)
    return pe
