"""Boolean mask <-> little-endian bit-packed uint8 (reference: src/chipmunk/ops/bitpack.py:4-69,
there as torch.compile'd elementwise code; here one HBM-bound CUDA kernel each way)."""
from .. import torch_ops as _t


def bitpack(mask):
    """-> (uint8 [ceil(numel/8)], original shape)"""
    return _t.bitpack(mask)


def bitunpack(packed, original_shape):
    return _t.bitunpack(packed, original_shape)
