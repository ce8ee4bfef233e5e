"""GPU tests of the sparse-delta modules (the classes the FLUX / HunyuanVideo blocks instantiate):
step scheduling, cache algebra and index bookkeeping, checked against dense PyTorch."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _fresh(cm, **attn):
    from chipmunk_b200.util.config import reset_to_defaults
    from chipmunk_b200.util import layer_counter as lc
    reset_to_defaults()
    cfg = cm.util.GLOBAL_CONFIG
    cfg["steps"] = 50
    cfg["attn"].update(first_n_dense_layers=0, **attn)
    cfg["mlp"].update(first_n_dense_layers=0, top_keys=0.3)
    lc.singleton.__init__(0, 0)
    return cfg


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("compressed,pad", [(False, False), (True, True)])
def test_sparse_diff_attn_steps(cm, oracle, cuda, compressed, pad):
    """FLUX-style (uncompressed, fused, counts%112) and Hunyuan-style (bit-packed mask, counts%128) flows."""
    _fresh(cm, top_keys=0.3, recompute_mask=compressed, should_compress_indices=compressed,
           pad_qkv_before_kernel=pad, counts_multiple_of=112, random_keys=0.0)
    layer_num, counter = cm.LayerCounter.build_for_layer(is_attn_sparse=True)
    attn = cm.SparseDiffAttn(layer_num, counter)
    g = torch.Generator(device=cuda).manual_seed(0)
    B, H, N = 1, 2, 1000
    q, k, v = (torch.randn(B, H, N, 128, device=cuda, generator=g).to(BF) for _ in range(3))
    dense = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    outs = [attn(q, k, v) for _ in range(3)]            # steps 0, 1 (full) and 2 (sparse), same inputs
    assert counter.cur_inference_step == 3
    for o in outs[:2]:
        assert _rel(o, dense) < 6e-3
    # sparse step on unchanged inputs: cache + sparse == dense up to bf16 rounding of the two adds
    assert _rel(outs[2], dense) < 1.2e-2
    # sparse step on perturbed inputs: exactly cache + sparse(q2, k, v) over the stored index sets (oracle)
    q2 = (q.float() + 0.3 * torch.randn(q.shape, device=cuda, generator=g)).to(BF)
    cache_before = attn.storage.get_out_cache().clone()
    if compressed:
        inds, counts = cm.ops.bitmask_to_indices(attn.storage.get_indices(), attn.mask_shape[0], 128, 192)
    else:
        inds, counts = attn.storage.get_indices(), attn.storage.get_counts()
    o3 = attn(q2, k, v)
    ref = oracle.csp_attn(q2.cpu(), k.cpu(), v.cpu(), cache_before.cpu(), inds.cpu(), counts.cpu(), 1)
    # two bf16 roundings (delta, then cache + delta) on both sides; the selection kernel emits the stored columns in
    # mask_to_indices order, which changes the fp32 summation order of a few rows by one bf16 ulp
    assert _rel(o3.cpu(), ref) < 5e-3
    assert torch.equal(attn.storage.get_out_cache(), cache_before), "a sparse step must not modify the cache"
    cache = attn.storage.get_out_cache()
    assert cache.shape == q.shape and cache.dtype == BF
    if compressed:
        assert attn.storage.get_indices().dtype == torch.uint8          # bit-packed mask is what is stored
    else:
        assert attn.storage.get_indices().dtype == torch.int32 and attn.storage.get_counts().min() > 0


def test_sparse_diff_mlp_steps(cm, cuda):
    cfg = _fresh(cm)
    cfg["mlp"].update(top_keys=0.3, random_keys=0.0, full_step_every=10, block_mask_cache=2)
    layer_num, counter = cm.LayerCounter.build_for_layer(is_mlp_sparse=True)
    torch.manual_seed(0)
    K, F = 256, 1024
    fc1 = torch.nn.Linear(K, F, device=cuda, dtype=BF)
    fc2 = torch.nn.Linear(F, K, device=cuda, dtype=BF)
    act = torch.nn.GELU(approximate="tanh")
    mlp = cm.SparseDiffMlp(layer_num, counter, fc1, act, fc2, 6)
    x0 = torch.randn(1, 512, K, device=cuda, dtype=BF)
    dense = lambda x: fc2(act(fc1(x)))
    with torch.no_grad():
        y0 = mlp(x0)                                   # step 0: full
        assert torch.equal(y0, dense(x0))
        y1 = mlp(x0)                                   # step 1: sparse, unchanged input -> delta ~ 0
        assert _rel(y1, dense(x0)) < 1e-2
        x1 = (x0.float() + 0.2 * torch.randn_like(x0.float())).to(BF)
        stale = _rel(dense(x0), dense(x1))
        y2 = mlp(x1)                                   # step 2: sparse, moved input
        assert _rel(y2, dense(x1)) < stale
        # every neuron active (top_keys = 1): the sparse step must reproduce the dense MLP
        cfg["mlp"]["top_keys"] = 1.0
        y3 = mlp(x1)
        assert _rel(y3, dense(x1)) < 1.5e-2
    counts = mlp.storage.get_counts()
    assert counts.shape == (1, 4) and int(counts.min()) == F and counts.dtype == torch.int32
    assert mlp.storage.get_sparse_act_T().shape == (1, F, 512)
    pa = mlp.storage.get_sparse_act_T()[0].t()
    assert _rel(pa, act(fc1(x1))[0]) < 1.5e-2          # the activation cache tracks gelu(fc1 x)
