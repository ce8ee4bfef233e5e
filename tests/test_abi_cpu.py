"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/chipmunk_b200.h declares (no compute calls without a GPU), the operator schemas are
registered under torch.ops.chipmunk, and the host-side state objects behave like the reference's."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "chipmunk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cm):
    lib = ctypes.CDLL(os.path.join(ROOT, "chipmunk_b200", "libchipmunk_b200.so"))
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/chipmunk_b200.h but not exported"
    assert set(syms) == set(cm._lib.EXPORTS)
    assert lib.cm_abi_version() == 2      # round 2: + cm_select_columns, cm_dense_attn_strided
    import re
    from chipmunk_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "chipmunk_b200.h")).read()
    assert int(re.search(r"#define\s+CM_ABI_VERSION\s+(\d+)", hdr).group(1)) == _lib.ABI_VERSION == lib.cm_abi_version()
    lib.cm_strerror.restype = ctypes.c_char_p
    assert b"16-byte" in lib.cm_strerror(-2)


def test_argument_validation_needs_no_gpu(cm):
    """Entry points reject bad arguments before touching the device."""
    lib = cm._lib.lib
    assert lib.cm_bitpack(None, None, -1, None) == -1
    assert lib.cm_mask_to_indices(None, None, None, 4, 0, 0, 128, None) == -1
    assert lib.cm_topk_indices(None, 0, None, None, 1, 1, 0, 0.5, 256, 0.0, None) == -1
    assert lib.cm_csp_mlp_mm1(None, None, None, None, None, None, None, 100, 64, 256, 256, 0, None) == -1
    s3 = (ctypes.c_int64 * 3)(8, 8, 8)
    assert lib.cm_csp_attn(None, None, None, None, None, None, 1, 1, 192, 192, s3, s3, s3, s3, 192, 3, 1, None) == -1


def test_operator_schemas_registered(cm):
    want = {
        "csp_mlp_mm1": 7, "csp_mlp_mm2_and_scatter_add": 9, "csp_attn": 7, "csp_128_attn": 5, "dense_attn": 3,
        "dense_colsum_attn": 4, "copy_indices": 4, "topk_indices": 6, "csp_scatter_add": 5, "mask_to_indices": 3,
    }
    for name, nargs in want.items():
        op = getattr(torch.ops.chipmunk, name)
        schema = op.default._schema
        assert len(schema.arguments) == nargs, (name, str(schema))
    s = str(torch.ops.chipmunk.csp_attn.default._schema)
    assert "int o_scale" in s and "Tensor indices_counts" in s
    s = str(torch.ops.chipmunk.mask_to_indices.default._schema)
    assert "int multiple_of, int pad_to_multiple_of" in s and "Tensor[]" in s


def test_no_cpu_fallback(cm):
    q = torch.zeros(1, 1, 192, 128, dtype=torch.bfloat16)
    idx = torch.zeros(1, 1, 1, 192, dtype=torch.int32)
    cnt = torch.zeros(1, 1, 1, dtype=torch.int32)
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.chipmunk.csp_128_attn(q, q, q, idx, cnt)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cm.ops.bitpack(torch.zeros(8, dtype=torch.bool))


def test_product_never_imports_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "chipmunk_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"


def test_layer_counter_schedule(cm):
    from chipmunk_b200.util import GLOBAL_CONFIG
    from chipmunk_b200.util.config import reset_to_defaults
    from chipmunk_b200.util.layer_counter import LayerCounter

    reset_to_defaults()
    GLOBAL_CONFIG["steps"] = 4
    c = LayerCounter(num_layers=2, num_sparse_submodules_per_layer=2)
    seen, full_attn, full_mlp = [], [], []
    for _ in range(4 * 2 * 2):
        full_attn.append(c.should_do_full_attn_step()); full_mlp.append(c.should_do_full_mlp_step())
        seen.append(c.increment())
    assert seen[0] == (0, 0, 0) and seen[1] == (0, 0, 1) and seen[2] == (0, 1, 0) and seen[4] == (1, 0, 0)
    assert full_attn[:8] == [True] * 8 and full_attn[8:12] == [False] * 4       # steps 0,1 full, step 2 sparse
    assert full_mlp[:4] == [True] * 4 and full_mlp[4] is False
    assert c.get_cur_coord() == (0, 0, 0) or c.cur_inference_step in (0, 3)      # rewinds at the end of a generation
    GLOBAL_CONFIG["attn"]["full_step_schedule"] = {0, 3}
    c.reset(); c.cur_inference_step = 3
    assert c.should_do_full_attn_step()
    c.cur_inference_step = 1
    assert not c.should_do_full_attn_step()
    reset_to_defaults()


def test_config_deep_merge(cm, tmp_path):
    from chipmunk_b200.util import GLOBAL_CONFIG, load_from_file
    from chipmunk_b200.util.config import reset_to_defaults

    reset_to_defaults()
    p = tmp_path / "c.yml"
    p.write_text("attn:\n  top_keys: 0.165\n  counts_multiple_of: 112\n  pad_qkv_before_kernel: false\nmlp:\n  top_keys: 0.3\n")
    load_from_file(str(p))
    assert GLOBAL_CONFIG["attn"]["top_keys"] == 0.165 and GLOBAL_CONFIG["attn"]["mbm"] == 192
    assert GLOBAL_CONFIG["mlp"]["top_keys"] == 0.3 and GLOBAL_CONFIG["mlp"]["counts_multiple_of"] == 256
    reset_to_defaults()


def test_storage_resident_mode(cm):
    from chipmunk_b200.util import AttnStorage, MlpStorage
    from chipmunk_b200.util.config import reset_to_defaults

    reset_to_defaults()
    s = MlpStorage(3)
    assert s.get_indices() is None
    t = torch.arange(6).reshape(1, 2, 3)
    s.set_indices(t)
    assert s.get_indices() is t
    s.load_async(); s.load_async_wait(); s.complete_cur_layer()
    a = AttnStorage(0, init_names=["indices", "out_cache"])
    assert a.out_cache is not None and a.out_cache.is_offload_enabled is False
    a.set_out_cache(t)
    assert a.get_out_cache() is t
    with pytest.raises(ValueError):
        from chipmunk_b200.util import MaybeOffloadedTensor
        MaybeOffloadedTensor("attn.bogus", 0, torch.float32, "cpu")


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/chipmunk_b200.h must compile as C99 with nothing but <stdint.h>."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "chipmunk_b200.h")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_parallel_shards_and_fused_guard(cm):
    """Host logic of the multi-GPU layer that needs no GPU: the fused (multicast) path refuses uneven head splits."""
    import torch
    from chipmunk_b200 import parallel
    assert parallel.shard_heads(24, 8, 3) == (9, 12)
    assert parallel.shard_heads(7, 2, 0) == (0, 4) and parallel.shard_heads(7, 2, 1) == (4, 7)
    assert callable(parallel.sparse_attention_head_parallel_fused)


def test_step_cache_glue(cm):
    """StepCache reproduces the reference's step-caching glue (examples/hunyuan/hyvideo/modules/models.py:732-741,834-835):
    skipped steps return the stored output and advance the layer counter's inference step."""
    from chipmunk_b200.util import StepCache, LayerCounter
    from chipmunk_b200.util.config import reset_to_defaults, GLOBAL_CONFIG
    reset_to_defaults()
    sc = StepCache()
    counter = LayerCounter(num_layers=1, num_sparse_submodules_per_layer=1)
    assert sc.try_skip(0, counter) is None and counter.cur_inference_step == 0
    x = torch.arange(12.0).reshape(3, 4)
    sc.store(x)
    x += 1                                              # the cache holds a copy, like the reference's clone()
    assert 7 in GLOBAL_CONFIG["step_caching"]["skip_step_schedule"]
    got = sc.try_skip(7, counter)
    assert torch.equal(got, torch.arange(12.0).reshape(3, 4)) and counter.cur_inference_step == 1
    assert sc.try_skip(8, counter) is None and counter.cur_inference_step == 1
    GLOBAL_CONFIG["step_caching"]["is_enabled"] = False
    assert sc.try_skip(7, counter) is None
    sc.reset()
    GLOBAL_CONFIG["step_caching"]["is_enabled"] = True
    with pytest.raises(RuntimeError):
        sc.try_skip(7, counter)
    reset_to_defaults()


def test_operators_trace_under_fake_tensors(cm):
    """the operators must be callable inside torch.compile'd blocks (the example models compile their
    transformer blocks).  Every operator has a fake (meta) kernel with the shapes / dtypes the real one allocates, and a
    sparse step traces to a graph that holds the `chipmunk::` nodes -- checked with fake CUDA tensors, no GPU needed."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from torch.fx.experimental.proxy_tensor import make_fx

    B, H, N = 1, 2, 500
    G, padN = (N + 191) // 192, 576
    with FakeTensorMode() as mode:
        bf = dict(dtype=torch.bfloat16, device="cuda")
        q = torch.empty(B, H, N, 128, **bf)
        idx = torch.empty(B, H, G, padN, dtype=torch.int32, device="cuda")
        cnt = torch.empty(B, H, G, dtype=torch.int32, device="cuda")
        o = torch.ops.chipmunk.csp_128_attn(q, q, q, idx, cnt)
        assert o.shape == q.shape and o.dtype == torch.bfloat16 and o.device.type == "cuda"
        o, l = torch.ops.chipmunk.dense_attn(q, q, q)
        assert o.shape == q.shape and l.shape == (B, H, N, 1) and l.dtype == torch.float32
        p = torch.empty(B, H, N, 1, dtype=torch.float32, device="cuda")
        o, cs, l = torch.ops.chipmunk.dense_colsum_attn(q, q, q, p)
        assert cs.shape == (B, H, G, N) and cs.dtype == torch.bfloat16 and l.shape == (B, H, N, 1)
        inds, counts = torch.ops.chipmunk.mask_to_indices(torch.empty(B, H, G, N, dtype=torch.bool, device="cuda"), 128, 192)
        assert inds.shape == (B, H, G, padN) and counts.shape == (B, H, G) and inds.dtype == counts.dtype == torch.int32
        # in-place operators return nothing
        assert torch.ops.chipmunk.csp_attn(q, q, q, torch.empty_like(q), idx, cnt, 1) is None
        M, K, F = 256, 128, 512
        a, w1, c = torch.empty(M, K, **bf), torch.empty(F, K, **bf), torch.empty(M, F, **bf)
        mi, mc = torch.empty(M // 128, F, dtype=torch.int32, device="cuda"), torch.empty(M // 128, dtype=torch.int32, device="cuda")
        assert torch.ops.chipmunk.csp_mlp_mm1(a, w1, c, torch.empty(F, **bf), torch.empty(F, M, **bf), mi, mc) is None
        assert torch.ops.chipmunk.csp_scatter_add(c[None], torch.empty(1, F, M, **bf), mi[None], mc[None], 6) is None
        assert torch.ops.chipmunk.csp_mlp_mm2_and_scatter_add(c[None], torch.empty(1, F, M, **bf), mi[None], mc[None], c[None],
                                                              torch.empty(1, F, K, **bf), torch.empty(1, M, K, **bf), 6, 0) is None
        act = torch.empty(1, M // 128, F, **bf)
        assert torch.ops.chipmunk.topk_indices(act, mi[None], mc[None], 0.7, 256, 0.0) is None
        assert torch.ops.chipmunk.copy_indices(act, torch.empty_like(act), mi[None], mc[None]) is None

        def sparse_step(q, k, v, cache, idx, cnt):
            o = cache.clone()
            torch.ops.chipmunk.csp_attn(q, k, v, o, idx, cnt, 1)
            return o + torch.ops.chipmunk.csp_128_attn(q, k, v, idx, cnt)

        gm = make_fx(sparse_step)(q, q, q, torch.empty_like(q), idx, cnt)
    targets = [str(n.target) for n in gm.graph.nodes if n.op == "call_function"]
    assert any("chipmunk.csp_attn" in t for t in targets) and any("chipmunk.csp_128_attn" in t for t in targets), targets
