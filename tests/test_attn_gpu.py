"""GPU parity: column-sparse / dense attention kernels vs the CPU oracle (oracle/chipmunk_oracle.py).

Tolerance (north_star: 1e-3 relative bf16): outputs are bf16, whose own rounding step is 2^-8
relative, so the bar is stated on the aggregate error:
    ||out - ref||_F / ||ref||_F <= 1e-3 * 4   (ref = oracle with the same rounding points, itself bf16)
    max |out - ref| <= 2 bf16 ulps of max|ref|
The oracle reduces with the exact row max, the kernel with a lazily updated one, so individual
P values can round differently; sums of hundreds of such terms agree far inside the bound.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(out, ref, rel=4e-3, ulps=2):
    out, ref = out.float().cpu(), ref.float().cpu()
    assert torch.isfinite(out).all()
    err = (out - ref).norm() / ref.norm().clamp_min(1e-12)
    amax = (out - ref).abs().max()
    bound = ulps * ref.abs().max() * 2.0 ** -8
    assert err <= rel, f"relative Frobenius error {err:.3e} > {rel}"
    assert amax <= bound, f"max abs error {amax:.3e} > {bound:.3e}"


def _rand_qkv(B, H, Nq, Nk, gen, strided=False):
    def mk(n):
        if strided:   # [n, B, H, D] storage viewed as [B,H,n,D], as in the reference's test_csp_attn.py:16-21
            return torch.randn(n, B, H, 128, generator=gen).to(torch.bfloat16).permute(1, 2, 0, 3)
        return torch.randn(B, H, n, 128, generator=gen).to(torch.bfloat16)
    return mk(Nq), mk(Nk), mk(Nk)


@pytest.mark.parametrize("B,H,N,count,strided", [
    (1, 2, 384, 128, False),
    (1, 2, 384, 256, True),
    (2, 3, 500, 112, False),      # ragged: N % 192 != 0, count multiple of 112 (fused FLUX path)
    (1, 2, 1000, 336, True),
    (1, 1, 192, 16, False),       # a single short step
    (1, 1, 777, 200, False),      # count not a multiple of 16: tail masked by position
])
def test_csp_attn_matches_oracle(cm, oracle, cuda, B, H, N, count, strided):
    gen = torch.Generator().manual_seed(1234 + N + count)
    q, k, v = _rand_qkv(B, H, N, N, gen, strided)
    G = (N + 191) // 192
    idx, cnt = oracle.random_index_sets(B, H, G, N, count, gen)
    idx_full = torch.zeros(B, H, G, N, dtype=torch.int32)
    idx_full[..., :count] = idx
    o0 = torch.randn(B, H, N, 128, generator=gen).to(torch.bfloat16)

    dq, dk, dv = q.to(cuda), k.to(cuda), v.to(cuda)
    if strided:
        dq, dk, dv = (t.permute(2, 0, 1, 3).contiguous().permute(1, 2, 0, 3) for t in (dq, dk, dv))
    di, dc = idx_full.to(cuda), cnt.to(cuda)

    # csp_128_attn: fresh output
    out = torch.ops.chipmunk.csp_128_attn(dq, dk, dv, di, dc)
    ref = oracle.csp_128_attn(q, k, v, idx_full, cnt)
    _close(out, ref)

    # csp_attn: accumulate with o_scale = +1 and -1
    for sc in (1, -1):
        o = o0.to(cuda).clone()
        torch.ops.chipmunk.csp_attn(dq, dk, dv, o, di, dc, sc)
        ref = oracle.csp_attn(q, k, v, o0, idx_full, cnt, sc)
        _close(o, ref)


@pytest.mark.parametrize("B,H,N,count,strided", [(1, 2, 384, 128, False), (2, 3, 500, 112, True), (1, 1, 777, 200, False)])
def test_csp_attn_add_equals_clone_plus_accumulate(cm, oracle, cuda, B, H, N, count, strided):
    """The fused out-of-place add-back (cm_csp_attn_add) must be BIT-identical to the reference sequence
    `o = cache.clone(); csp_attn(q, k, v, o, idx, cnt, s)` (modules/attn.py:165-190), for s = +1 and -1, must leave
    the cache untouched, must copy the cache through for groups with count 0, and must match the oracle."""
    gen = torch.Generator().manual_seed(99 + N)
    q, k, v = _rand_qkv(B, H, N, N, gen, strided)
    G = (N + 191) // 192
    idx, cnt = oracle.random_index_sets(B, H, G, N, count, gen)
    idx_full = torch.zeros(B, H, G, N, dtype=torch.int32)
    idx_full[..., :count] = idx
    cnt[0, 0, G - 1] = 0                                   # one empty group
    cache = torch.randn(B, H, N, 128, generator=gen).to(torch.bfloat16)
    dq, dk, dv = q.to(cuda), k.to(cuda), v.to(cuda)
    if strided:
        dq, dk, dv = (t.permute(2, 0, 1, 3).contiguous().permute(1, 2, 0, 3) for t in (dq, dk, dv))
    di, dc, dcache = idx_full.to(cuda), cnt.to(cuda), cache.to(cuda)
    for sc in (1, -1):
        two_pass = dcache.clone()
        torch.ops.chipmunk.csp_attn(dq, dk, dv, two_pass, di, dc, sc)
        fused = cm.ops.csp_attn_add(dq, dk, dv, dcache, di, dc, sc)
        assert torch.equal(dcache.cpu(), cache), "the cache must not be modified"
        assert torch.equal(fused, two_pass), "fused add-back differs from clone + in-place accumulate"
        _close(fused, oracle.csp_attn(q, k, v, cache, idx_full, cnt, sc))
    rows = slice((G - 1) * 192, N)
    assert torch.equal(fused[0, 0, rows].cpu(), cache[0, 0, rows])


def test_csp_attn_identity_indices_is_sdpa(cm, oracle, cuda):
    """The reference's own known-answer test (src/chipmunk/tests/test_csp_attn.py:30-38):
    indices = arange(n), counts = n, o = 0, o_scale = 1  ==>  F.scaled_dot_product_attention."""
    gen = torch.Generator().manual_seed(7)
    for n, strided in ((4480 // 8, False), (672, True)):
        B, H = 1, 3
        q, k, v = _rand_qkv(B, H, n, n, gen)
        G = (n + 191) // 192
        idx = torch.arange(n, dtype=torch.int32).repeat(B, H, G, 1).contiguous()
        cnt = torch.full((B, H, G), n, dtype=torch.int32)
        dq, dk, dv = q.to(cuda), k.to(cuda), v.to(cuda)
        if strided:
            dq, dk, dv = (t.permute(2, 0, 1, 3).contiguous().permute(1, 2, 0, 3) for t in (dq, dk, dv))
        o = torch.zeros(B, H, n, 128, dtype=torch.bfloat16, device=cuda)
        torch.ops.chipmunk.csp_attn(dq, dk, dv, o, idx.to(cuda), cnt.to(cuda), 1)
        ref = oracle.sdpa(q, k, v)
        _close(o, ref.to(torch.bfloat16), rel=6e-3, ulps=3)


def test_csp_attn_zero_count_and_varying_counts(cm, oracle, cuda):
    gen = torch.Generator().manual_seed(99)
    B, H, N = 1, 2, 576
    q, k, v = _rand_qkv(B, H, N, N, gen)
    G = 3
    idx = torch.stack([torch.randperm(N, generator=gen) for _ in range(B * H * G)]).int().reshape(B, H, G, N)
    cnt = torch.tensor([[[0, 128, 384], [256, 0, 16]]], dtype=torch.int32)
    o0 = torch.randn(B, H, N, 128, generator=gen).to(torch.bfloat16)
    o = o0.to(cuda).clone()
    torch.ops.chipmunk.csp_attn(q.to(cuda), k.to(cuda), v.to(cuda), o, idx.to(cuda), cnt.to(cuda), 1)
    ref = oracle.csp_attn(q, k, v, o0, idx, cnt, 1)
    _close(o, ref)
    # groups with count 0 must be bit-identical to the input
    assert torch.equal(o[0, 0, :192].cpu(), o0[0, 0, :192])
    out = torch.ops.chipmunk.csp_128_attn(q.to(cuda), k.to(cuda), v.to(cuda), idx.to(cuda), cnt.to(cuda))
    assert (out[0, 0, :192] == 0).all()


def test_csp_attn_argument_errors(cm, cuda):
    q = torch.zeros(1, 1, 192, 64, dtype=torch.bfloat16, device=cuda)
    idx = torch.zeros(1, 1, 1, 192, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(1, 1, 1, dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError, match="Head dimension must be 128"):
        torch.ops.chipmunk.csp_128_attn(q, q, q, idx, cnt)
    q = torch.zeros(1, 1, 192, 128, dtype=torch.bfloat16, device=cuda)
    with pytest.raises(RuntimeError, match="o_scale must be 1 or -1"):
        torch.ops.chipmunk.csp_attn(q, q, q, q.clone(), idx, cnt, 2)
    with pytest.raises(RuntimeError, match="32-bit integer"):
        torch.ops.chipmunk.csp_128_attn(q, q, q, idx.long(), cnt)


@pytest.mark.parametrize("B,H,N", [(1, 2, 384), (1, 2, 500), (2, 1, 1024)])
def test_dense_attn_matches_oracle_and_sdpa(cm, oracle, cuda, B, H, N):
    gen = torch.Generator().manual_seed(N)
    q, k, v = _rand_qkv(B, H, N, N, gen)
    o, l = torch.ops.chipmunk.dense_attn(q.to(cuda), k.to(cuda), v.to(cuda))
    ro, rl = oracle.dense_attn(q, k, v)
    _close(o, ro)
    assert l.shape == (B, H, N, 1) and l.dtype == torch.float32
    torch.testing.assert_close(l.cpu(), rl, rtol=2e-3, atol=0)
    _close(o, oracle.sdpa(q, k, v).to(torch.bfloat16), rel=6e-3, ulps=3)   # reference test_dense_attn.py:29-34


@pytest.mark.parametrize("B,H,N", [(1, 2, 384), (1, 2, 500), (2, 1, 1000)])
def test_dense_colsum_attn_matches_oracle(cm, oracle, cuda, B, H, N):
    """o, l as dense_attn; cs = per-192-row-group column sums of exp(s)*p (reference
    dense_colsum_attn.cu:267-277 accumulates them in bf16 with atomics; kernel and oracle
    accumulate in fp32 and round once, so they agree to bf16 rounding: 1 ulp = 2^-8 relative)."""
    gen = torch.Generator().manual_seed(N + 5)
    q, k, v = _rand_qkv(B, H, N, N, gen)
    # p = l of a slightly different q, as on consecutive denoising steps
    q_prev = (q.float() + 0.05 * torch.randn(q.shape, generator=gen)).to(torch.bfloat16)
    _, p = oracle.dense_attn(q_prev, k, v)
    ro, rcs, rl = oracle.dense_colsum_attn(q, k, v, p)
    o, cs, l = torch.ops.chipmunk.dense_colsum_attn(q.to(cuda), k.to(cuda), v.to(cuda), p.to(cuda))
    _close(o, ro)
    torch.testing.assert_close(l.cpu(), rl, rtol=2e-3, atol=0)
    G = (N + 191) // 192
    assert cs.shape == (B, H, G, N) and cs.dtype == torch.bfloat16
    torch.testing.assert_close(cs.float().cpu(), rcs.float(), rtol=1.2e-2, atol=1e-6)
    # wrapper with padded p (what SparseDiffAttn passes): same values
    pn = ((N + 191) // 192) * 192
    pp = torch.zeros(B, H, pn, 1); pp[:, :, :N] = p
    o2, cs2, l2 = cm.ops.dense_colsum_attn(q.to(cuda), k.to(cuda), v.to(cuda), pp.to(cuda))
    assert torch.equal(cs2, cs[..., : (N + 191) // 192, :N]) and l2.shape[-2] == pn
