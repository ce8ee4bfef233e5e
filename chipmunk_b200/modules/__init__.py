"""Sparse-delta modules with the reference's names (src/chipmunk/modules/__init__.py)."""
from .attn import SparseDiffAttn
from .mlp import SparseDiffMlp

__all__ = ["SparseDiffAttn", "SparseDiffMlp"]
