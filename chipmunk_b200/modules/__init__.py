"""Sparse-delta modules with the reference's names (src/chipmunk/modules/__init__.py)."""
from .attn import SparseDiffAttn
from .mlp import SparseDiffMlp


def quantize_fp8(model, *args, **kwargs):
    """The reference's fp8 preview path (src/chipmunk/modules/mlp_fp8.py:353, taken when `mlp.is_fp8: true`) is outside
    the scope of this library (bf16 path only).  The name exists because the FLUX example imports it
    unconditionally (examples/flux/src/flux/util.py:15) and calls it only under `GLOBAL_CONFIG['mlp']['is_fp8']`
    (:349-350): calling it fails loudly instead of silently running bf16."""
    raise RuntimeError("chipmunk_b200 implements the bf16 column-sparse path only: set `mlp.is_fp8: false` "
                       "(the reference's fp8 preview, modules/mlp_fp8.py, is not part of this library)")


__all__ = ["SparseDiffAttn", "SparseDiffMlp", "quantize_fp8"]
