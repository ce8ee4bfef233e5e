"""The module algebra against golden vectors produced by the REFERENCE's own module code.

tests/golden/make_golden_modules.py ran src/chipmunk/modules/{attn,mlp}.py (with the reference's op wrappers, LayerCounter,
config and layer storage) on CPU, the CUDA operators underneath replaced by the oracle; this file replays the same inputs
through chipmunk_b200's modules on the GPU and compares every step's output, the stored masks / index sets and the caches.
Index sets and packed masks are compared exactly (the inputs have a wide gap at every selection threshold); bf16 tensors
within the tolerances written at each assert (two bf16 roundings per step on both sides, different fp32 summation orders).
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from module_cases import ROW_STRIDE, attn_step_inputs, from_bits, mlp_step_input  # noqa: E402

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fresh(cm):
    from chipmunk_b200.util.config import reset_to_defaults
    from chipmunk_b200.util import layer_counter as lc
    reset_to_defaults()
    lc.singleton.__init__(0, 0)
    cfg = cm.util.GLOBAL_CONFIG
    cfg["steps"] = 50
    return cfg


def _close(got: torch.Tensor, want: torch.Tensor, rel: float, scale: torch.Tensor = None, what: str = ""):
    """||got - want||_F <= rel * ||scale||_F (scale = want unless a reference magnitude is given)."""
    got, want = got.float().cpu(), want.float()
    ref = (want if scale is None else scale.float()).norm()
    err = (got - want).norm() / ref
    assert float(err) <= rel, f"{what}: relative error {float(err):.3e} > {rel:.1e}"


@pytest.mark.parametrize("name,torch_selection,resident", [("hunyuan", False, False), ("hunyuan", True, False), ("hunyuan", False, True),
                                                          ("hunyuan", True, True), ("flux", False, False), ("flux", True, False)])
def test_sparse_diff_attn_matches_reference_modules(cm, cuda, monkeypatch, name, torch_selection, resident):
    z = np.load(os.path.join(GOLD, "modules_attn.npz"))
    compressed, pad, multiple_of, tt, th, tw, txt_len, tk, local_voxels, H, salt = (int(x) for x in z[f"{name}_cfg"])
    N = tt * th * tw + txt_len
    cfg = _fresh(cm)
    cfg["attn"].update(is_enabled=True, first_n_dense_layers=0, top_keys=tk / N, random_keys=0.0, local_voxels=local_voxels,
                       local_1d_window=0, full_step_every=10, full_step_schedule=None, recompute_mask=bool(compressed),
                       should_compress_indices=bool(compressed), counts_multiple_of=multiple_of,
                       pad_qkv_before_kernel=bool(pad), torch_selection=torch_selection, random_columns=0.0,
                       keep_indices_resident=resident)       # index lists kept in HBM between steps: same results
    if torch_selection:
        # the reference formulation draws its 1 % random columns with torch.randint: none, as in the fixture
        real = torch.randint
        monkeypatch.setattr(torch, "randint", lambda lo, hi, shape, **kw: torch.ones(shape, dtype=kw.get("dtype", torch.int64), device=kw.get("device")) if hi == 100 else real(lo, hi, shape, **kw))
    import chipmunk_b200.modules.attn as A
    for n in ("singleton_static_mask", "singleton_video_query_groups", "singleton_static_words", "singleton_group_flags"):
        monkeypatch.setattr(A, n, None)            # restored after the test: other tests expect no static mask
    layer, counter = cm.LayerCounter.build_for_layer(is_attn_sparse=True)
    attn = cm.SparseDiffAttn(layer, counter)
    if compressed:
        attn.initialize_static_mask((tt, th, tw), txt_len, H, cuda)
    q0, k0, v0 = (from_bits(z[f"{name}_{t}0"]) for t in "qkv")
    outs = []
    for s in range(4):
        q, k, v = (t.to(cuda) for t in attn_step_inputs(q0, k0, v0, s, salt))
        outs.append(attn(q, k, v))
        if s == 1:
            cache1 = attn.storage.get_out_cache().clone()
    want = [from_bits(z[f"{name}_o{s}"]) for s in range(4)]
    # full steps (0: dense, 1: dense + column sums + selection): the dense output, two kernels' bf16 rounding apart
    for s in (0, 1):
        _close(outs[s][:, :, ::ROW_STRIDE], want[s], 4e-3, what=f"{name} full step {s}")
    # what the full step stored: the selected columns, exactly
    if compressed:
        assert tuple(attn.mask_shape[0]) == tuple(int(x) for x in z[f"{name}_mask_shape"])
        assert np.array_equal(attn.storage.get_indices().cpu().numpy(), z[f"{name}_packed_mask"]), "stored bit mask differs"
    else:
        sets = z[f"{name}_index_sets"]
        inds, cnt = attn.storage.get_indices().cpu(), attn.storage.get_counts().cpu()
        assert int(cnt.min()) == int(cnt.max()) == tk
        for h in range(H):
            for g in range(inds.shape[2]):
                assert np.array_equal(np.sort(inds[0, h, g, :tk].numpy()), sets[h]), "stored index set differs"
    # the cache = dense - sparse of step 1: a difference of two nearly equal tensors, so measured against the output's norm
    _close(cache1[:, :, ::ROW_STRIDE], from_bits(z[f"{name}_cache"]), 6e-3, scale=want[1], what=f"{name} cache")
    _close(attn.storage.get_lse_constants()[:, :, :N:ROW_STRIDE], torch.from_numpy(z[f"{name}_lse"]), 3e-3, what=f"{name} lse")
    # sparse steps: cache + sparse(q, k, v) on moved inputs
    for s in (2, 3):
        _close(outs[s][:, :, ::ROW_STRIDE], want[s], 6e-3, what=f"{name} sparse step {s}")
    assert torch.equal(attn.storage.get_out_cache(), cache1), "sparse steps must leave the cache alone"
    assert bool(attn._resident) == (resident and bool(compressed))
    assert counter.cur_inference_step == 4


def test_sparse_diff_mlp_matches_reference_module(cm, cuda):
    z = np.load(os.path.join(GOLD, "modules_mlp.npz"))
    w1, b1, w2, b2, x0 = (from_bits(z[n]) for n in ("w1", "b1", "w2", "b2", "x0"))
    dirs, active = torch.from_numpy(z["dirs"]), z["active"]
    F, K = w1.shape
    cfg = _fresh(cm)
    cfg["mlp"].update(is_enabled=True, is_fp8=False, top_keys=active.shape[1] / F, random_keys=0.0, full_step_every=10,
                      block_mask_cache=2, first_n_dense_layers=0, counts_multiple_of=256, bm=128, mbm=128)
    layer, counter = cm.LayerCounter.build_for_layer(is_mlp_sparse=True)
    fc1 = torch.nn.Linear(K, F, device=cuda, dtype=BF)
    fc2 = torch.nn.Linear(F, K, device=cuda, dtype=BF)
    with torch.no_grad():
        fc1.weight.copy_(w1); fc1.bias.copy_(b1); fc2.weight.copy_(w2); fc2.bias.copy_(b2)
    mlp = cm.SparseDiffMlp(layer, counter, fc1, torch.nn.GELU(approximate="tanh"), fc2, 6)
    with torch.no_grad():
        for s in range(3):                     # step 0 full; steps 1, 2 sparse with the indices recomputed
            y = mlp(mlp_step_input(x0, dirs, s).to(cuda))
            # cuBLAS bf16 vs the CPU's bf16 linear on the full step; mm1 / mm2 vs the oracle on the sparse ones
            _close(y[:, ::ROW_STRIDE], from_bits(z[f"y{s}"]), 6e-3, what=f"mlp step {s}")
            if s > 0:
                inds, cnt = mlp.storage.get_indices().cpu(), mlp.storage.get_counts().cpu()
                assert int(cnt.min()) == int(cnt.max()) == active.shape[1]
                for b in range(active.shape[0]):
                    assert np.array_equal(np.sort(inds[0, b, : active.shape[1]].numpy()), active[b]), "selected neurons differ"
    _close(mlp.storage.get_sparse_act_T()[:, ::ROW_STRIDE], from_bits(z["sparse_act_T"]), 6e-3, what="activation cache")
    _close(mlp.storage.get_out_cache()[:, ::ROW_STRIDE], from_bits(z["out_cache"]), 6e-3, what="output cache")
    _close(mlp.storage.get_blockmean_mid_cache(), from_bits(z["blockmean_mid_cache"]), 6e-3, what="block-mean cache")
