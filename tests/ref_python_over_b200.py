"""The REFERENCE's own Python host code running on this repo's kernels (INTEGRATION.md option 2), checked against the golden vectors.

    python tests/ref_python_over_b200.py            (GPU box: run by tests/test_ref_python_gpu.py in a fresh interpreter)
    python tests/ref_python_over_b200.py --cpu-oracle    (authoring container: the same scaffolding over the CPU oracle's
                                                          operators, which must reproduce the fixtures bit for bit)

What runs unmodified, from the byte-for-byte staged copy in oracle/_ref/chipmunk_py/ (oracle/build_ref.py; git-ignored, it
travels to the GPU box with the snapshot): src/chipmunk/modules/{attn,mlp}.py (`SparseDiffAttn`, `SparseDiffMlp`),
src/chipmunk/ops/*.py (the padding / slicing wrappers, the torch bitpack / bitunpack, the voxel masks) and src/chipmunk/util/*
(LayerCounter, GLOBAL_CONFIG, AttnStorage / MlpStorage / MaybeOffloadedTensor with its CUDA streams).  Underneath,
`torch.ops.chipmunk.*` are the ten operators registered by chipmunk_b200.torch_ops -- the sm_100a kernels behind the C ABI.
The reference's top-level `chipmunk/__init__.py` only imports the compiled `cuda` module and `triton`; both are empty modules
here (`csp_mlp_mm2_function_ptr = 0`: the B200 operator ignores it).

Scaffolding, the same as in tests/golden/make_golden_modules.py which produced the expected values:
  * TORCHDYNAMO_DISABLE=1: `@torch.compile` on the reference's bitpack / bitunpack is an optimisation, the eager function runs;
  * inside the reference's ops/attn.py `torch.empty` gives zeros: the reference pads Q with uninitialised rows;
  * the 1 % random columns of `random_and_topk` (`torch.randint(0, 100, ...) == 0`) are switched off;
  * `offloading.global_disable_offloading = True` (the FLUX example's setting; no pinned 1.2 GB buffers).
The flows and tolerances are those of tests/test_modules_golden_gpu.py (which replays them through chipmunk_b200's own modules).
"""
import os
import sys

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import importlib  # noqa: E402
import types  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
STAGED = os.path.join(ROOT, "oracle", "_ref", "chipmunk_py")
sys.path.insert(0, GOLD)
sys.path.insert(0, ROOT)
from module_cases import ROW_STRIDE, attn_step_inputs, from_bits, mlp_step_input  # noqa: E402

BF = torch.bfloat16


class _TorchWithZeroedEmpty:
    """`torch` as seen by the reference's ops/attn.py: identical, except that empty() returns zeros."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def empty(*a, **k):
        return torch.zeros(*a, **k)


def install_on_b200():
    """`chipmunk` = the staged reference Python; torch.ops.chipmunk.* = this repo's CUDA operators."""
    assert "chipmunk" not in sys.modules, "needs a fresh interpreter: `chipmunk` must resolve to the staged reference code"
    import chipmunk_b200  # noqa: F401   loads libchipmunk_b200.so and registers the ten operators (CUDA dispatch key)
    pkg = types.ModuleType("chipmunk")
    pkg.__path__ = [os.path.join(STAGED, "chipmunk")]
    sys.modules["chipmunk"] = pkg
    sys.modules["chipmunk.cuda"] = types.ModuleType("chipmunk.cuda")
    tr = types.ModuleType("chipmunk.triton")
    tr.csp_mlp_mm2_function_ptr, tr.csp_mlp_mm2, tr.csp_mlp_mm1_fp8 = 0, None, None
    sys.modules["chipmunk.triton"] = tr
    mods = {n: importlib.import_module(n) for n in ("chipmunk.util", "chipmunk.ops", "chipmunk.modules.attn", "chipmunk.modules.mlp")}
    pkg.util, pkg.ops = mods["chipmunk.util"], mods["chipmunk.ops"]
    for m in mods.values():
        assert m.__file__.startswith(STAGED), m.__file__
    return mods


def install_on_cpu_oracle():
    import make_golden_modules as G
    G.REF_SRC = STAGED
    lib, mods, _ = G.install_reference_on_cpu()        # also turns torch.empty into zeros, globally
    mods["_library"] = lib                             # the operators live as long as their torch.library.Library object
    return mods


def close(got, want, rel, scale=None, what=""):
    got, want = got.float().cpu(), want.float()
    ref = (want if scale is None else scale.float()).norm()
    err = float((got - want).norm() / ref)
    print(f"  {what}: relative error {err:.3e} (bound {rel:.1e})")
    assert err <= rel, f"{what}: relative error {err:.3e} > {rel:.1e}"


def fresh(mods):
    cfgmod = importlib.import_module("chipmunk.util.config")
    import copy
    cfgmod.GLOBAL_CONFIG.clear()
    cfgmod.GLOBAL_CONFIG.update(copy.deepcopy(cfgmod.BASE_CONFIG))
    cfg = cfgmod.GLOBAL_CONFIG
    cfg["steps"] = 50
    cfg["offloading"]["global_disable_offloading"] = True
    lc = importlib.import_module("chipmunk.util.layer_counter")
    lc.singleton.__init__(0, 0)
    return cfg, lc


def attention_flow(mods, name, dev):
    z = np.load(os.path.join(GOLD, "modules_attn.npz"))
    compressed, pad, multiple_of, tt, th, tw, txt_len, tk, local_voxels, H, salt = (int(x) for x in z[f"{name}_cfg"])
    N = tt * th * tw + txt_len
    cfg, lc = fresh(mods)
    cfg["attn"].update(is_enabled=True, first_n_dense_layers=0, top_keys=tk / N, random_keys=0.0, local_voxels=local_voxels,
                       local_1d_window=0, full_step_every=10, full_step_schedule=None, recompute_mask=bool(compressed),
                       should_compress_indices=bool(compressed), counts_multiple_of=multiple_of, pad_qkv_before_kernel=bool(pad))
    A = mods["chipmunk.modules.attn"]
    A.singleton_static_mask = A.singleton_video_query_groups = None
    layer, counter = lc.LayerCounter.build_for_layer(is_attn_sparse=True)
    attn = A.SparseDiffAttn(layer, counter)
    if compressed:
        attn.initialize_static_mask((tt, th, tw), txt_len, H, dev)
    q0, k0, v0 = (from_bits(z[f"{name}_{t}0"]) for t in "qkv")
    outs = []
    for s in range(4):
        q, k, v = (t.to(dev) for t in attn_step_inputs(q0, k0, v0, s, salt))
        outs.append(attn(q, k, v))
        if s == 1:
            cache1 = attn.storage.get_out_cache().clone()
    want = [from_bits(z[f"{name}_o{s}"]) for s in range(4)]
    for s in (0, 1):
        close(outs[s][:, :, ::ROW_STRIDE], want[s], 4e-3, what=f"{name} full step {s}")
    if compressed:
        assert tuple(attn.mask_shape[0]) == tuple(int(x) for x in z[f"{name}_mask_shape"])
        assert np.array_equal(attn.storage.get_indices().cpu().numpy(), z[f"{name}_packed_mask"]), "stored bit mask differs"
        print(f"  {name}: stored bit mask identical ({z[f'{name}_packed_mask'].size} bytes)")
    else:
        sets = z[f"{name}_index_sets"]
        inds, cnt = attn.storage.get_indices().cpu(), attn.storage.get_counts().cpu()
        assert int(cnt.min()) == int(cnt.max()) == tk
        for h in range(H):
            for g in range(inds.shape[2]):
                assert np.array_equal(np.sort(inds[0, h, g, :tk].numpy()), sets[h]), "stored index set differs"
        print(f"  {name}: stored index sets identical")
    close(cache1[:, :, ::ROW_STRIDE], from_bits(z[f"{name}_cache"]), 6e-3, scale=want[1], what=f"{name} cache")
    close(attn.storage.get_lse_constants()[:, :, :N:ROW_STRIDE], torch.from_numpy(z[f"{name}_lse"]), 3e-3, what=f"{name} lse")
    for s in (2, 3):
        close(outs[s][:, :, ::ROW_STRIDE], want[s], 6e-3, what=f"{name} sparse step {s}")
    assert torch.equal(attn.storage.get_out_cache(), cache1), "sparse steps must leave the cache alone"
    assert counter.cur_inference_step == 4


def mlp_flow(mods, dev):
    z = np.load(os.path.join(GOLD, "modules_mlp.npz"))
    w1, b1, w2, b2, x0 = (from_bits(z[n]) for n in ("w1", "b1", "w2", "b2", "x0"))
    dirs, active = torch.from_numpy(z["dirs"]), z["active"]
    F, K = w1.shape
    cfg, lc = fresh(mods)
    cfg["mlp"].update(is_enabled=True, is_fp8=False, top_keys=active.shape[1] / F, random_keys=0.0, full_step_every=10,
                      block_mask_cache=2, first_n_dense_layers=0, counts_multiple_of=256, bm=128, mbm=128)
    layer, counter = lc.LayerCounter.build_for_layer(is_mlp_sparse=True)
    fc1 = torch.nn.Linear(K, F, device=dev, dtype=BF)
    fc2 = torch.nn.Linear(F, K, device=dev, dtype=BF)
    with torch.no_grad():
        fc1.weight.copy_(w1); fc1.bias.copy_(b1); fc2.weight.copy_(w2); fc2.bias.copy_(b2)
    mlp = mods["chipmunk.modules.mlp"].SparseDiffMlp(layer, counter, fc1, torch.nn.GELU(approximate="tanh"), fc2, 6)
    with torch.no_grad():
        for s in range(3):
            y = mlp(mlp_step_input(x0, dirs, s).to(dev))
            close(y[:, ::ROW_STRIDE], from_bits(z[f"y{s}"]), 6e-3, what=f"mlp step {s}")
            if s > 0:
                inds, cnt = mlp.storage.get_indices().cpu(), mlp.storage.get_counts().cpu()
                assert int(cnt.min()) == int(cnt.max()) == active.shape[1]
                for b in range(active.shape[0]):
                    assert np.array_equal(np.sort(inds[0, b, : active.shape[1]].numpy()), active[b]), "selected neurons differ"
    print("  mlp: selected neuron sets identical")
    close(mlp.storage.get_sparse_act_T()[:, ::ROW_STRIDE], from_bits(z["sparse_act_T"]), 6e-3, what="activation cache")
    close(mlp.storage.get_out_cache()[:, ::ROW_STRIDE], from_bits(z["out_cache"]), 6e-3, what="output cache")
    close(mlp.storage.get_blockmean_mid_cache(), from_bits(z["blockmean_mid_cache"]), 6e-3, what="block-mean cache")


def main():
    cpu = "--cpu-oracle" in sys.argv
    if not os.path.isdir(os.path.join(STAGED, "chipmunk", "modules")):
        sys.exit("oracle/_ref/chipmunk_py is not staged: run `python oracle/build_ref.py` where /root/reference is mounted")
    if cpu:
        mods, dev = install_on_cpu_oracle(), torch.device("cpu")
    else:
        assert torch.cuda.is_available(), "needs a GPU (or --cpu-oracle for the scaffolding check)"
        mods, dev = install_on_b200(), torch.device("cuda:0")
        sys.modules["chipmunk.ops.attn"].torch = _TorchWithZeroedEmpty()      # only that module's view of torch
    real_randint = torch.randint
    torch.randint = lambda lo, hi, shape, **kw: (torch.ones(shape, dtype=kw.get("dtype", torch.int64), device=kw.get("device"))
                                                 if hi == 100 else real_randint(lo, hi, shape, **kw))
    try:
        for name in ("hunyuan", "flux"):
            print(f"reference SparseDiffAttn, {name} flow, on {'the CPU oracle' if cpu else 'chipmunk_b200 kernels'}:")
            attention_flow(mods, name, dev)
        print(f"reference SparseDiffMlp on {'the CPU oracle' if cpu else 'chipmunk_b200 kernels'}:")
        mlp_flow(mods, dev)
    finally:
        torch.randint = real_randint
    if not cpu:
        torch.cuda.synchronize()
    print("REFERENCE-PYTHON-OVER-" + ("ORACLE" if cpu else "B200") + " OK")


if __name__ == "__main__":
    main()
