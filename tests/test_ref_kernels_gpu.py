"""GPU parity against the REFERENCE'S OWN index kernels.

`oracle/build_ref.py` compiles /root/reference/csrc/indexed_io/{mask_to_indices,topk_indices,
copy_indices,scatter_add}.cu unmodified for sm_100a into oracle/_ref/libchipmunk_ref_indexed_io.so (registered as
`torch.ops.chipmunk_ref.*`); the built library travels to the GPU box.  Here the real reference kernels
and ours run on the same inputs:

* mask_to_indices — counts and `indices[..., :counts]` bit-exact INCLUDING the emission order
  (mask_to_indices.cu:47-86: lane-strided scan, thread 0 pads with unset columns in ascending order);
* topk_indices    — counts exact, kept-column SET exact (the reference's order is an atomicInc race,
  topk_indices.cu:108-113), padding entries drawn from the rejected columns; with random_amount > 0 the
  kept set is still exact because the XORWOW seeding and the short-circuit draw order are reproduced;
* copy_indices    — destination tensor bit-exact;
* csp_scatter_add — the activation cache after the scatter bit-exact (one bf16 add per element: the reference does it
  with `cp.reduce.async.bulk ... add.bf16`, scatter_add.cu:50-64; ours in registers).

The library is test infrastructure: nothing under chipmunk_b200/ loads it.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                       "libchipmunk_ref_indexed_io.so")


@pytest.fixture(scope="module")
def ref(cuda):
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    torch.ops.load_library(REF_LIB)
    return torch.ops.chipmunk_ref


@pytest.mark.parametrize("shape,dens,mult", [
    ((1, 2, 3, 384), 0.2, 128), ((1, 2, 2, 500), 0.5, 112), ((2, 1, 2, 1000), 0.07, 128),
    ((1, 1, 2, 256), 0.0, 128), ((1, 2, 3, 4608), 0.165, 112), ((1, 1, 2, 119056), 0.07, 128),
    ((1, 1, 3, 333), 0.9, 16), ((1, 24, 24, 4608), 0.17, 128), ((2, 3, 4, 4096), 0.2, 128),
])
def test_mask_to_indices_equals_reference_kernel(cm, ref, cuda, shape, dens, mult):
    g = torch.Generator().manual_seed(int(dens * 100) + shape[-1])
    mask = (torch.rand(shape, generator=g) < dens).to(cuda)
    rinds, rcounts = ref.mask_to_indices(mask, mult, 192)
    torch.cuda.synchronize()
    inds, counts = cm.ops.mask_to_indices(mask, mult, 192)
    packed, shp = cm.ops.bitpack(mask)
    inds2, counts2 = cm.ops.bitmask_to_indices(packed, shp, mult, 192)
    torch.cuda.synchronize()
    assert inds.shape == rinds.shape and inds.dtype == rinds.dtype == torch.int32
    assert torch.equal(counts, rcounts) and torch.equal(counts2, rcounts)
    n = shape[-1]
    # compare only the defined prefix of each row: the reference leaves the tail uninitialised (torch::empty),
    # and when fewer than `padding` unset columns exist it stops early (mask_to_indices.cu:74-83)
    nnz = mask.sum(-1, dtype=torch.int32)
    defined = torch.minimum(rcounts, torch.full_like(rcounts, n)).unsqueeze(-1)
    pos = torch.arange(inds.shape[-1], device=cuda).view(1, 1, 1, -1)
    valid = pos < defined
    assert (nnz <= rcounts).all()
    assert torch.equal(torch.where(valid, inds, 0), torch.where(valid, rinds, 0))
    assert torch.equal(torch.where(valid, inds2, 0), torch.where(valid, rinds, 0))


def _topk_compare(cm, ref, cuda, act, sparsity, mult, rnd):
    B, R, C = act.shape
    act = act.to(cuda)
    ri = torch.full((B, R, C), -7, dtype=torch.int32, device=cuda)
    rc = torch.zeros(B, R, dtype=torch.int32, device=cuda)
    ref.topk_indices(act, ri, rc, sparsity, mult, rnd)
    torch.cuda.synchronize()
    oi = torch.full((B, R, C), -7, dtype=torch.int32, device=cuda)
    oc = torch.zeros(B, R, dtype=torch.int32, device=cuda)
    cm.ops.topk_indices(act, oi, oc, sparsity, mult, rnd)
    torch.cuda.synchronize()
    assert torch.equal(oc, rc), "counts differ from the reference kernel"
    ri, oi, rc = ri.cpu(), oi.cpu(), rc.cpu()
    for b in range(B):
        for r in range(R):
            c = int(rc[b, r])
            if sparsity in (0.0, 1.0):
                assert torch.equal(oi[b, r, :c], ri[b, r, :c])
                continue
            # the reference's kept set = everything it wrote before padding; its count before padding is not
            # returned, so recover it from the threshold rule: kept columns are those in BOTH outputs' prefixes
            rs, os_ = set(ri[b, r, :c].tolist()), set(oi[b, r, :c].tolist())
            assert len(rs) == c and len(os_) == c, "duplicate or missing entries in the valid prefix"
            # padding (< mult entries) may legitimately differ (the reference picks whichever rejected
            # columns win an atomicAdd); everything else must be the same set
            assert len(rs ^ os_) <= 2 * (mult - 1), f"kept sets differ beyond padding: {len(rs ^ os_)}"
            common = rs & os_
            assert len(common) >= c - (mult - 1)
            assert min(os_) >= 0 and max(os_) < C


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("sparsity,mult", [(0.7, 256), (0.3, 128), (0.95, 256), (0.0, 256), (1.0, 256), (0.8, 1)])
def test_topk_indices_equals_reference_kernel(cm, ref, cuda, dtype, sparsity, mult):
    g = torch.Generator().manual_seed(21)
    act = torch.randn(2, 5, 12288, generator=g).abs().to(dtype)
    _topk_compare(cm, ref, cuda, act, sparsity, mult, 0.0)


def test_topk_indices_exact_set_without_padding(cm, ref, cuda):
    """multiple_of = 1: no padding, so the valid prefix must be EXACTLY the reference's set."""
    g = torch.Generator().manual_seed(22)
    act = torch.randn(1, 36, 12288, generator=g).abs().to(torch.bfloat16).to(cuda)
    for rnd in (0.0, 0.05, 0.3):
        ri = torch.full((1, 36, 12288), -7, dtype=torch.int32, device=cuda)
        rc = torch.zeros(1, 36, dtype=torch.int32, device=cuda)
        oi, oc = ri.clone(), rc.clone()
        ref.topk_indices(act, ri, rc, 0.7, 1, rnd)
        torch.cuda.synchronize()
        cm.ops.topk_indices(act, oi, oc, 0.7, 1, rnd)
        torch.cuda.synchronize()
        assert torch.equal(oc, rc), f"random_amount={rnd}: counts differ"
        for r in range(36):
            c = int(rc[0, r])
            assert set(oi[0, r, :c].tolist()) == set(ri[0, r, :c].tolist()), f"random_amount={rnd} row {r}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_copy_indices_equals_reference_kernel(cm, ref, cuda, dtype):
    g = torch.Generator().manual_seed(23)
    B, M, R, F = 2, 3, 4, 1024
    src = torch.randn(B, M * R, F, generator=g).to(dtype).to(cuda)
    dst = torch.randn(B, M * R, F, generator=g).to(dtype).to(cuda)
    inds = torch.stack([torch.randperm(F, generator=g) for _ in range(B * M)]).int().reshape(B, M, F).to(cuda)
    cnts = torch.tensor([[256, 0, 1024], [16, 512, 3]], dtype=torch.int32, device=cuda)
    rd, od = dst.clone(), dst.clone()
    ref.copy_indices(src, rd, inds, cnts)
    torch.cuda.synchronize()
    cm.ops.copy_indices(src, od, inds, cnts)
    torch.cuda.synchronize()
    assert torch.equal(od, rd)


@pytest.mark.parametrize("M,F,counts", [(4, 1024, [256, 0, 1024, 16]), (6, 12288, [3840, 3840, 256, 2304, 12288, 512])])
def test_scatter_add_equals_reference_kernel(cm, ref, cuda, M, F, counts):
    g = torch.Generator().manual_seed(24 + F)
    packed = torch.randn(1, M * 128, F, generator=g).to(torch.bfloat16).to(cuda)
    pa = torch.randn(1, F, M * 128, generator=g).to(torch.bfloat16).to(cuda)
    inds = torch.stack([torch.randperm(F, generator=g) for _ in range(M)]).int().reshape(1, M, F).to(cuda)
    cnts = torch.tensor([counts], dtype=torch.int32, device=cuda)
    r, o = pa.clone(), pa.clone()
    ref.csp_scatter_add(packed, r, inds, cnts, M)           # one block per token block (its grid-stride loop reads counts[blockIdx.x])
    torch.cuda.synchronize()
    torch.ops.chipmunk.csp_scatter_add(packed, o, inds, cnts, 6)
    torch.cuda.synchronize()
    assert torch.equal(o, r), "scatter-add differs from the reference kernel"
    assert not torch.equal(o, pa) or sum(counts) == 0


def test_fused_cache_update_equals_mm1_plus_reference_scatter(cm, ref, cuda):
    """chipmunk.ops.mlp fuses the activation-cache update into csp_mlp_mm1's epilogue; the reference runs its
    scatter-add kernel on mm1's packed output (ops/mlp.py:64-92 -> csp_mlp_mm2_and_scatter_add).  Same inputs:
    the cache after our fused mm1 must equal the cache after [our mm1 without update + the REFERENCE scatter kernel]."""
    from chipmunk_b200 import torch_ops as T
    g = torch.Generator().manual_seed(31)
    M, K, F = 512, 256, 2048
    bf = torch.bfloat16
    x = torch.randn(M, K, generator=g).to(bf).to(cuda)
    w1 = (torch.randn(F, K, generator=g) / K ** 0.5).to(bf).to(cuda)
    b1 = (0.1 * torch.randn(F, generator=g)).to(bf).to(cuda)
    pa = torch.randn(F, M, generator=g).to(bf).to(cuda)
    idx = torch.stack([torch.randperm(F, generator=g) for _ in range(M // 128)]).int().to(cuda)
    cnt = torch.tensor([512, 2048, 256, 1024], dtype=torch.int32, device=cuda)
    c_a = torch.zeros(M, F, dtype=bf, device=cuda)
    c_b = torch.zeros(M, F, dtype=bf, device=cuda)
    pa_two_pass, pa_fused = pa.clone(), pa.clone()
    T.mlp_mm1(x, w1, c_a, b1, pa_two_pass, idx, cnt, False)
    ref.csp_scatter_add(c_a[None], pa_two_pass[None], idx[None], cnt[None], M // 128)
    torch.cuda.synchronize()
    T.mlp_mm1(x, w1, c_b, b1, pa_fused, idx, cnt, True)
    torch.cuda.synchronize()
    assert torch.equal(c_a, c_b)
    assert torch.equal(pa_fused, pa_two_pass), "fused cache update differs from mm1 + the reference's scatter-add kernel"
