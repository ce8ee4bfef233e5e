"""Operator wrappers with the names and argument meaning of the reference's `chipmunk.ops`
(src/chipmunk/ops/__init__.py:1-7).  Every function dispatches to `torch.ops.chipmunk.*`,
i.e. to the sm_100a kernels behind the C ABI; there is no eager or CPU fallback."""
from .mlp import run_e2e as mlp, mm1, mm2_fused, mm2_unfused
from .indexed_io import copy_indices, topk_indices, mask_to_indices, scatter_add, bitmask_to_indices, select_columns, pack_rows_to_words
from .attn import csp_attn, csp_attn_add, dense_attn, dense_colsum_attn
from .bitpack import bitpack, bitunpack
from .patch import patchify, unpatchify, patchify_rope
from . import voxel

__all__ = ["mlp", "copy_indices", "topk_indices", "mask_to_indices", "scatter_add", "csp_attn", "csp_attn_add",
           "dense_attn", "dense_colsum_attn", "bitpack", "bitunpack", "bitmask_to_indices", "select_columns", "pack_rows_to_words",
           "mm1", "mm2_fused", "mm2_unfused", "patchify", "unpatchify", "patchify_rope", "voxel"]
