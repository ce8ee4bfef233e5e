"""GPU parity against the REFERENCE's own Triton kernels, run on the B200 next to ours.

The reference has exactly one mm2 implementation -- the Triton kernel `csp_mlp_mm2_kernel`
(/root/reference/src/chipmunk/triton/csp_mlp_mm2.py:24-109) -- and an fp8-capable Triton mm1
(`matmul_kernel_one_fp8`, triton/csp_mlp_mm1.py:37-164) that also runs with bf16 operands and unit scales.  Both are
arch-portable: they JIT for sm_100.  oracle/build_ref.py stages the two files byte-for-byte into the git-ignored
oracle/_ref/triton_ref/ (no copy enters the repo); here they are imported from there and run on identical inputs.
This pins `cm_csp_mlp_mm2` (and the CPU oracle's mm2) to the reference implementation itself, and cross-checks
`cm_csp_mlp_mm1` + its fused cache update against the reference's other mm1.

Rounding points.  mm2: both compute bf16(acc) and then one bf16 add onto `out` (csp_mlp_mm2.py:100-101): the only
difference is fp32 summation order, so results agree to 1 bf16 ulp on isolated elements.  mm1: the Triton kernel
rounds gelu(.) to bf16 FIRST, subtracts the cache in bf16 and stores bf16(gelu) as the new cache
(csp_mlp_mm1.py:127-140); the reference's CUDA mm1 (csp_mlp_mm1.cu:366-375), which ours follows, subtracts in fp32 and
rounds once; the cache then becomes bf16(cache + c).  Both pairs differ by at most one extra bf16 rounding.
"""
import importlib.util
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
HERE = os.path.dirname(os.path.abspath(__file__))
TRITON_REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "triton_ref")


def _load(name):
    path = os.path.join(TRITON_REF, name + ".py")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run `python oracle/build_ref.py` where /root/reference exists (it stages the "
                    "reference's Triton kernels into the git-ignored oracle/_ref/)")
    spec = importlib.util.spec_from_file_location("chipmunk_ref_triton_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)          # csp_mlp_mm2.py launches a kernel at import (its function-pointer probe)
    return mod


@pytest.fixture(scope="module")
def ref_mm2(cuda):
    return _load("csp_mlp_mm2")


@pytest.fixture(scope="module")
def ref_mm1(cuda):
    return _load("csp_mlp_mm1")


def _problem(M, K, F, counts, seed, cuda):
    g = torch.Generator(device=cuda).manual_seed(seed)
    x = torch.randn(M, K, device=cuda, generator=g).to(BF)
    w1 = (torch.randn(F, K, device=cuda, generator=g) / K ** 0.5).to(BF)
    b1 = (0.1 * torch.randn(F, device=cuda, generator=g)).to(BF)
    w2t = (torch.randn(F, K, device=cuda, generator=g) / F ** 0.5).to(BF)
    pa = torch.randn(F, M, device=cuda, generator=g).to(BF)
    out = torch.randn(M, K, device=cuda, generator=g).to(BF)
    packed = torch.randn(M, F, device=cuda, generator=g).to(BF)
    idx = torch.stack([torch.randperm(F, device=cuda, generator=g) for _ in range(M // 128)]).int()
    cnt = torch.tensor(counts, dtype=torch.int32, device=cuda)
    return x, w1, b1, w2t, pa, out, packed, idx, cnt


def _ulp_close(a, b, what, mag=None, max_ulps=2, frac_off=2e-2):
    """bf16 results of the same fp32 value up to summation order: equal or adjacent bf16 values nearly everywhere.
    `mag`: magnitude of the operands of the final bf16 add (a result that cancels is only accurate to THEIR ulp)."""
    a32, b32 = a.float(), b.float()
    assert torch.isfinite(a32).all(), what
    scale = torch.maximum(a32.abs(), b32.abs())
    if mag is not None:
        scale = torch.maximum(scale, mag.float().to(scale.device))
    tol = max_ulps * 2.0 ** -7 * scale.clamp_min(2.0 ** -6)      # one bf16 ulp of x is in (2^-8 |x|, 2^-7 |x|]
    bad = (a32 - b32).abs() > tol
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} elements differ by more than {max_ulps} bf16 ulps"
    off = float((a != b).float().mean())
    assert off <= frac_off, f"{what}: {off:.4f} of the elements are not bit-identical"
    rel = float((a32 - b32).norm() / b32.norm())
    assert rel <= 1e-3, f"{what}: relative Frobenius error {rel:.2e}"


@pytest.mark.parametrize("M,K,F,counts", [
    (512, 1024, 2048, [1024, 2048, 256, 512]),
    (4608, 3072, 12288, [3840] * 36),                        # BASELINE configs[1]: FLUX, 70 % sparse
    (4608, 3072, 12288, [3840, 256, 12288, 2304, 1024, 7680] * 6),
])
def test_mm2_matches_reference_triton_kernel(cm, oracle, cuda, ref_mm2, M, K, F, counts):
    x, w1, b1, w2t, pa, out, packed, idx, cnt = _problem(M, K, F, counts, 5 + M, cuda)
    ours, theirs = out.clone(), out.clone()
    ref_mm2.csp_mlp_mm2(packed, w2t, idx, cnt, theirs, 148)
    torch.cuda.synchronize()
    from chipmunk_b200 import torch_ops as T
    T.mlp_mm2(packed, w2t, ours, None, idx, cnt, False)
    torch.cuda.synchronize()
    assert not torch.equal(theirs, out)
    mag = out.float().abs() + (theirs.float() - out.float()).abs()       # |out| + |bf16(acc)|
    _ulp_close(ours, theirs, "mm2 vs the reference Triton kernel", mag)
    if M <= 512:      # and the CPU oracle's restatement is pinned by the same kernel
        ref = oracle.csp_mlp_mm2(packed.cpu(), w2t.cpu(), idx.cpu(), cnt.cpu(), out.cpu())
        _ulp_close(theirs.cpu(), ref, "reference Triton mm2 vs the CPU oracle", mag.cpu())


@pytest.mark.parametrize("M,K,F,counts", [
    (512, 1024, 2048, [1024, 2048, 256, 512]),
    (4608, 3072, 12288, [3840] * 36),
])
def test_mm1_and_cache_update_match_reference_triton_kernel(cm, oracle, cuda, ref_mm1, M, K, F, counts):
    """bf16 operands, unit scales: c = bf16(bf16(gelu(x w1^T + b)) - cache), cache <- bf16(gelu(...))
    (csp_mlp_mm1.py:119-140).  Ours: c = bf16(gelu(...) - cache), cache <- bf16(cache + c).  The Triton kernel
    works on 128-column tiles (:91-92), so counts are multiples of 128 here."""
    x, w1, b1, w2t, pa, out, packed, idx, cnt = _problem(M, K, F, counts, 9 + M, cuda)
    one = torch.ones(1, device=cuda, dtype=torch.float32)
    # The kernel is @triton.autotune'd over 8 configs and UPDATES the cache in place (csp_mlp_mm1.py:140), so the tuning
    # runs of a first call compound on its arguments: tune on scratch copies first (the choice is cached per M, N, K).
    ref_mm1.csp_mlp_mm1(x, w1, b1, idx, cnt, pa.clone(), torch.zeros(M, F, dtype=BF, device=cuda), one, one)
    torch.cuda.synchronize()
    c_ref = torch.zeros(M, F, dtype=BF, device=cuda)
    pa_ref = pa.clone()
    # the wrapper reads `N, K = b.shape` and strides (b.stride(1), b.stride(0)) (csp_mlp_mm1.py:145-159): b is [F, K]
    ref_mm1.csp_mlp_mm1(x, w1, b1, idx, cnt, pa_ref, c_ref, one, one)
    torch.cuda.synchronize()
    from chipmunk_b200 import torch_ops as T
    c = torch.zeros(M, F, dtype=BF, device=cuda)
    pa_new = pa.clone()
    T.mlp_mm1(x, w1, c, b1, pa_new, idx, cnt, True)
    torch.cuda.synchronize()
    for mb, n in enumerate(counts):
        rows = slice(mb * 128, (mb + 1) * 128)
        a, b = c[rows, :n].float(), c_ref[rows, :n].float()
        # one extra bf16 rounding (of gelu, magnitude <= ~4) on their side: 2^-9 * 4 absolute, plus ours
        assert float((a - b).abs().max()) <= 3 * 2.0 ** -8 * max(1.0, float(b.abs().max())), f"mm1 block {mb}"
        assert float((a - b).norm() / b.norm()) <= 4e-3, f"mm1 block {mb}"
        f = idx[mb, :n].long()
        pn, pr = pa_new[f][:, rows].float(), pa_ref[f][:, rows].float()
        assert float((pn - pr).abs().max()) <= 3 * 2.0 ** -8 * max(1.0, float(pr.abs().max())), f"cache block {mb}"
        assert float((pn - pr).norm() / pr.norm()) <= 4e-3, f"cache block {mb}"
        # rows of the cache that are not selected stay untouched in both
        if n < F:
            g = idx[mb, n:].long()
            assert torch.equal(pa_new[g][:, rows], pa[g][:, rows])
