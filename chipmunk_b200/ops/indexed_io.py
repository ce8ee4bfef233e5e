"""Index selection / copy wrappers (reference: src/chipmunk/ops/indexed_io.py:4-36)."""
from __future__ import annotations

from typing import List

import torch

from .. import torch_ops as _t


def copy_indices(bm_fc1, bm_mid_cache, indices, counts) -> None:
    torch.ops.chipmunk.copy_indices(bm_fc1, bm_mid_cache, indices, counts)


def topk_indices(activations, indices_out, counts_out, sparsity_amount: float, multiple_of: int, rk: float) -> None:
    torch.ops.chipmunk.topk_indices(activations, indices_out, counts_out, sparsity_amount, multiple_of, rk)


def scatter_add(packed, unpacked, indices, counts, num_sms: int) -> None:
    """packed [M,F], unpacked [F,M], indices [M/128,F], counts [M/128] (batch dim added here)."""
    torch.ops.chipmunk.csp_scatter_add(packed.unsqueeze(0), unpacked.unsqueeze(0), indices.unsqueeze(0),
                                       counts.unsqueeze(0), num_sms)


def mask_to_indices(mask, multiple_of: int, pad_to_multiple_of: int) -> List[torch.Tensor]:
    return torch.ops.chipmunk.mask_to_indices(mask, multiple_of, pad_to_multiple_of)


def bitmask_to_indices(packed, mask_shape, multiple_of: int, pad_to_multiple_of: int) -> List[torch.Tensor]:
    """bitunpack + mask_to_indices in one pass over the packed bits (B200 addition)."""
    return _t.bitmask_to_indices(packed, mask_shape, multiple_of, pad_to_multiple_of)


def select_columns(cs, k: int, multiple_of: int, random_prob: float = 0.0, static_words=None, group_is_sparse=None,
                   seed=None, pad_to_multiple_of: int = 192, want_packed: bool = True, want_indices: bool = True):
    """random_and_topk + bitpack + mask_to_indices of a full step in one kernel (B200 addition; reference
    src/chipmunk/modules/attn.py:76-84,132-150).  Returns (packed, mask_shape, indices, counts)."""
    return _t.select_columns(cs, k, multiple_of, random_prob, static_words, group_is_sparse, seed,
                             pad_to_multiple_of, want_packed, want_indices)


def pack_rows_to_words(mask2d):
    return _t.pack_rows_to_words(mask2d)
