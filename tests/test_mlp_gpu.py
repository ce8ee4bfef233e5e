"""GPU parity: column-sparse MLP kernels (mm1, mm2, scatter_add, ops.mlp) vs the CPU oracle.

Tolerance: the kernels and the oracle round at the same points (bf16 output of an fp32
accumulator), so elementwise differences come from fp32 summation order and tanh.approx only:
    max |out - ref| <= 2 bf16 ulps of max|ref|   and   relative Frobenius error <= 2e-3.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _close(out, ref, rel=2e-3, ulps=2):
    out, ref = out.float().cpu(), ref.float().cpu()
    assert torch.isfinite(out).all()
    err = (out - ref).norm() / ref.norm().clamp_min(1e-12)
    amax = (out - ref).abs().max()
    bound = ulps * ref.abs().max() * 2.0 ** -8
    assert err <= rel, f"relative Frobenius error {err:.3e} > {rel}"
    assert amax <= bound, f"max abs error {amax:.3e} > {bound:.3e}"


def _problem(M, K, F, N, counts, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g).to(BF)
    w1 = (torch.randn(F, K, generator=g) / K ** 0.5).to(BF)
    b1 = (0.1 * torch.randn(F, generator=g)).to(BF)
    w2t = (torch.randn(F, N, generator=g) / F ** 0.5).to(BF)
    pa = torch.randn(F, M, generator=g).to(BF)
    out = torch.randn(M, N, generator=g).to(BF)
    idx = torch.stack([torch.randperm(F, generator=g) for _ in range(M // 128)]).int()
    cnt = torch.tensor(counts, dtype=torch.int32)
    assert cnt.numel() == M // 128
    return x, w1, b1, w2t, pa, out, idx, cnt


CASES = [
    (256, 128, 512, 256, [256, 512], 0),
    (384, 192, 768, 512, [768, 0, 272], 1),       # a zero block and a count that is only a multiple of 16
    (128, 3072, 1024, 256, [512], 2),              # FLUX reduction length
    (512, 64, 2048, 768, [1024, 2048, 256, 16], 3),
]


@pytest.mark.parametrize("M,K,F,N,counts,seed", CASES)
def test_mm1_matches_oracle(cm, oracle, cuda, M, K, F, N, counts, seed):
    x, w1, b1, w2t, pa, out, idx, cnt = _problem(M, K, F, N, counts, seed)
    c0 = torch.full((M, F), 7.0, dtype=BF)
    ref = oracle.csp_mlp_mm1(x, w1, c0, b1, pa, idx, cnt)
    c = c0.to(cuda)
    dpa = pa.to(cuda)
    torch.ops.chipmunk.csp_mlp_mm1(x.to(cuda), w1.to(cuda), c, b1.to(cuda), dpa, idx.to(cuda), cnt.to(cuda))
    _close(c, ref)
    assert torch.equal(dpa.cpu(), pa), "csp_mlp_mm1 must not modify the cache"
    # untouched columns stay untouched
    for mb, n in enumerate(counts):
        assert (c[mb * 128:(mb + 1) * 128, n:] == 7.0).all()


@pytest.mark.parametrize("M,K,F,N,counts,seed", CASES)
def test_mm2_and_scatter_match_oracle(cm, oracle, cuda, M, K, F, N, counts, seed):
    x, w1, b1, w2t, pa, out, idx, cnt = _problem(M, K, F, N, counts, seed)
    packed = torch.randn(M, F, generator=torch.Generator().manual_seed(seed + 100)).to(BF)
    ref_out = oracle.csp_mlp_mm2(packed, w2t, idx, cnt, out)
    ref_pa = oracle.csp_scatter_add(packed, pa, idx, cnt)
    dp, dpa, dout = packed.to(cuda), pa.to(cuda), out.to(cuda)
    torch.ops.chipmunk.csp_mlp_mm2_and_scatter_add(dp[None], dpa[None], idx.to(cuda)[None], cnt.to(cuda)[None],
                                                   dp[None], w2t.to(cuda)[None], dout[None], 6, 0)
    _close(dout, ref_out)
    assert torch.equal(dpa.cpu(), ref_pa), "scatter-add is one bf16 add per element: must be bit-exact"
    # the stand-alone scatter op
    dpa2 = pa.to(cuda)
    torch.ops.chipmunk.csp_scatter_add(dp[None], dpa2[None], idx.to(cuda)[None], cnt.to(cuda)[None], 6)
    assert torch.equal(dpa2.cpu(), ref_pa)


@pytest.mark.parametrize("M,K,F,N,counts,seed", CASES[:2])
def test_ops_mlp_end_to_end(cm, oracle, cuda, M, K, F, N, counts, seed):
    """chipmunk.ops.mlp (run_e2e): fused cache update in mm1's epilogue + mm2."""
    x, w1, b1, w2t, pa, out, idx, cnt = _problem(M, K, F, N, counts, seed)
    ref_out, ref_pa, _ = oracle.mlp_sparse_step(x, w1, b1, w2t, idx, cnt, pa, out)
    dpa, dout = pa.to(cuda), out.to(cuda)
    cm.ops.mlp(x.to(cuda), w1.to(cuda), b1.to(cuda), w2t.to(cuda), idx.to(cuda), cnt.to(cuda), dpa, dout, 6)
    _close(dout, ref_out, rel=4e-3, ulps=3)
    _close(dpa, ref_pa, rel=4e-3, ulps=2)


def test_mm1_reference_harness_fixture(cm, oracle, cuda):
    """The reference's own seeded fixture (csrc/mlp/csp_mlp_mm1.cu:443-485,590-602): M=3840, N=12288,
    K=3072, mt19937(42) U(-0.5,0.5), reversed identity indices, pass if |gpu - cpu_gemm| <= 0.1.
    Checked here on the first 4 token blocks and 2048 neurons of that problem (same generator
    stream order: A, B, bias, pa_cache), against the harness formula evaluated in fp32."""
    M, N, K = 3840, 12288, 3072
    Ms, Ns = 512, 2048
    vals = oracle.std_mt19937_uniform(42, M * K + K * N)     # A then B, as the harness draws them
    a = torch.from_numpy(vals[: M * K].reshape(M, K)).float()[:Ms]
    b = torch.from_numpy(vals[M * K:].reshape(N, K)).float()[:Ns]
    g = torch.Generator().manual_seed(42)
    bias = torch.rand(Ns, generator=g) - 0.5
    pa = torch.rand(Ns, Ms, generator=g) - 0.5
    idx = torch.arange(Ns - 1, -1, -1, dtype=torch.int32).repeat(Ms // 128, 1)
    cnt = torch.full((Ms // 128,), Ns, dtype=torch.int32)
    ab, bb, biasb, pab = a.to(BF), b.to(BF), bias.to(BF), pa.to(BF)
    pre = ab.float() @ bb.float().flip(0).t() + biasb.float().flip(0)[None]
    ref = 0.5 * pre * (1 + torch.tanh(0.7978845608028654 * (pre + 0.044715 * pre ** 3))) - pab.float().flip(0).t()
    c = torch.zeros(Ms, Ns, dtype=BF, device=cuda)
    torch.ops.chipmunk.csp_mlp_mm1(ab.to(cuda), bb.to(cuda), c, biasb.to(cuda), pab.to(cuda), idx.to(cuda), cnt.to(cuda))
    assert (c.float().cpu() - ref).abs().max() <= 0.1


def test_mlp_argument_errors(cm, cuda):
    z = lambda *s: torch.zeros(*s, dtype=BF, device=cuda)
    idx = torch.zeros(1, 256, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(1, dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        torch.ops.chipmunk.csp_mlp_mm1(z(128, 100), z(256, 100), z(128, 256), z(256), z(256, 128), idx, cnt)
    with pytest.raises(RuntimeError, match="multiple of 128"):
        torch.ops.chipmunk.csp_mlp_mm1(z(100, 64), z(256, 64), z(100, 256), z(256), z(256, 100), idx, cnt)
    with pytest.raises(RuntimeError, match="int32"):
        torch.ops.chipmunk.csp_mlp_mm1(z(128, 64), z(256, 64), z(128, 256), z(256), z(256, 128), idx.long(), cnt)
