"""GPU parity for the integer path: bit-exact against the oracle and the reference-generated
fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bitpack_matches_reference_fixture(cm, cuda):
    z = np.load(os.path.join(GOLDEN, "bitpack.npz"))
    n = len([k for k in z.files if k.startswith("mask")])
    for i in range(n):
        mask = torch.from_numpy(z[f"mask{i}"]).to(cuda)
        packed, shape = cm.ops.bitpack(mask)
        assert packed.dtype == torch.uint8 and tuple(shape) == tuple(mask.shape)
        assert np.array_equal(packed.cpu().numpy(), z[f"packed{i}"])
        assert torch.equal(cm.ops.bitunpack(packed, shape), mask)


@pytest.mark.parametrize("shape", [(1, 3, 5, 1000), (2, 2, 7, 4608), (1, 1, 3, 119056), (1, 2, 2, 37), (1, 1, 1, 5)])
def test_bitpack_roundtrip_and_oracle(cm, oracle, cuda, shape):
    g = torch.Generator().manual_seed(sum(shape))
    mask = torch.rand(shape, generator=g) < 0.13
    packed, shp = cm.ops.bitpack(mask.to(cuda))
    ref, _ = oracle.bitpack(mask)
    assert torch.equal(packed.cpu(), ref)
    assert torch.equal(cm.ops.bitunpack(packed, shp).cpu(), mask)


def _check_m2i(inds, counts, rinds, rcounts):
    assert torch.equal(counts.cpu(), rcounts)
    inds, rinds = inds.cpu().reshape(-1, inds.shape[-1]), rinds.reshape(-1, rinds.shape[-1])
    for r, c in enumerate(rcounts.reshape(-1).tolist()):
        valid = rinds[r, :c] >= 0       # the oracle marks "not enough unset columns to pad" with -1
        assert torch.equal(inds[r, :c][valid], rinds[r, :c][valid]), f"row {r}"


@pytest.mark.parametrize("shape,dens,mult", [
    ((1, 2, 3, 384), 0.2, 128), ((1, 2, 2, 500), 0.5, 112), ((2, 1, 2, 1000), 0.07, 128),
    ((1, 1, 2, 256), 0.0, 128), ((1, 1, 2, 256), 1.0, 128), ((1, 2, 3, 4608), 0.165, 112),
    ((1, 1, 2, 119056), 0.07, 128), ((1, 1, 3, 333), 0.9, 16),
])
def test_mask_to_indices_bit_exact(cm, oracle, cuda, shape, dens, mult):
    g = torch.Generator().manual_seed(int(dens * 100) + shape[-1])
    mask = torch.rand(shape, generator=g) < dens
    rinds, rcounts = oracle.mask_to_indices(mask, mult, 192)
    inds, counts = cm.ops.mask_to_indices(mask.to(cuda), mult, 192)
    assert inds.shape == rinds.shape and inds.dtype == torch.int32
    _check_m2i(inds, counts, rinds, rcounts)
    # fused bit-packed variant: identical output
    packed, shp = cm.ops.bitpack(mask.to(cuda))
    inds2, counts2 = cm.ops.bitmask_to_indices(packed, shp, mult, 192)
    _check_m2i(inds2, counts2, rinds, rcounts)


def test_mask_to_indices_reference_fixture_sets(cm, cuda):
    z = np.load(os.path.join(GOLDEN, "masktoinds.npz"))
    n = len([k for k in z.files if k.startswith("mask")])
    for i in range(n):
        mask = torch.from_numpy(z[f"mask{i}"]).to(cuda)
        inds, counts = cm.ops.mask_to_indices(mask, int(z[f"mult{i}"]), 192)
        assert np.array_equal(counts.cpu().numpy(), z[f"counts{i}"])
        fi = inds.cpu().numpy().reshape(-1, inds.shape[-1])
        fr = z[f"inds{i}"].reshape(-1, z[f"inds{i}"].shape[-1])
        for r, k in enumerate(z[f"nnz{i}"].reshape(-1)):
            assert set(fi[r, :k]) == set(fr[r, :k])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("sparsity,mult", [(0.7, 256), (0.3, 128), (0.95, 256), (0.0, 256), (1.0, 256)])
def test_topk_indices_matches_oracle(cm, oracle, cuda, dtype, sparsity, mult):
    g = torch.Generator().manual_seed(11)
    B, R, C = 2, 5, 12288
    act = torch.randn(B, R, C, generator=g).abs().to(dtype)
    rcounts, kept = oracle.topk_indices(act, sparsity, mult)
    inds = torch.full((B, R, C), -7, dtype=torch.int32, device=cuda)
    counts = torch.empty(B, R, dtype=torch.int32, device=cuda)
    cm.ops.topk_indices(act.to(cuda), inds, counts, sparsity, mult, 0.0)
    assert torch.equal(counts.cpu(), rcounts)
    inds = inds.cpu().numpy()
    for b in range(B):
        for r in range(R):
            k = kept[b * R + r]
            got = inds[b, r, : k.size]
            assert np.array_equal(got, k), "kept columns must be the oracle's set, in ascending order"
            pad = inds[b, r, k.size: int(rcounts[b, r])]
            assert len(set(pad)) == pad.size and not (set(pad) & set(k)) and (pad >= 0).all() and (pad < C).all()


def test_topk_indices_random_keep_is_superset(cm, oracle, cuda):
    g = torch.Generator().manual_seed(12)
    act = torch.randn(1, 4, 4096, generator=g).abs().to(torch.bfloat16)
    _, kept = oracle.topk_indices(act, 0.8, 1)
    inds = torch.empty(1, 4, 4096, dtype=torch.int32, device=cuda)
    counts = torch.empty(1, 4, dtype=torch.int32, device=cuda)
    cm.ops.topk_indices(act.to(cuda), inds, counts, 0.8, 1, 0.05)
    for r in range(4):
        got = set(inds[0, r, : int(counts[0, r])].cpu().tolist())
        assert set(kept[r].tolist()) <= got
        extra = len(got) - kept[r].size
        assert 0.02 * 4096 * 0.8 < extra < 0.09 * 4096 * 0.8      # ~5 % of the rejected columns


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_copy_indices_matches_oracle(cm, oracle, cuda, dtype):
    g = torch.Generator().manual_seed(13)
    B, M, R, F = 2, 3, 4, 1024
    src = torch.randn(B, M * R, F, generator=g).to(dtype)
    dst = torch.randn(B, M * R, F, generator=g).to(dtype)
    inds = torch.stack([torch.randperm(F, generator=g) for _ in range(B * M)]).int().reshape(B, M, F)
    cnts = torch.tensor([[256, 0, 1024], [16, 512, 3]], dtype=torch.int32)
    ref = oracle.copy_indices(src, dst, inds, cnts)
    d = dst.to(cuda)
    cm.ops.copy_indices(src.to(cuda), d, inds.to(cuda), cnts.to(cuda))
    assert torch.equal(d.cpu(), ref)
