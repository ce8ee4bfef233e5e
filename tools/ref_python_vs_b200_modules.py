"""Same kernels, two host sides: the REFERENCE's own Python modules (staged in oracle/_ref/chipmunk_py, see
tests/ref_python_over_b200.py) against chipmunk_b200's modules, both calling this repo's `torch.ops.chipmunk.*` operators on the
same inputs.  Reports, per shape, the time of a full step and of a sparse step on either side and how far the outputs are apart:
what the module-level fusions (one-kernel column selection, bit mask -> indices in one kernel, the fused out-of-place add-back,
no padding copies) are worth on top of the kernels.

    python tools/ref_python_vs_b200_modules.py [--out gpurun_out/x.json]          (GPU box)
"""
import json
import os
import sys

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ref_python_over_b200 as R  # noqa: E402


def timed(fn, reps):
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    outs = None
    t0.record()
    for _ in range(reps):
        outs = fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps, outs


def run_side(make_attn, cfg, q, k, v, n_sparse):
    """steps 0, 1 full (dense; dense + column sums + selection + cache build), then n_sparse sparse steps."""
    attn = make_attn()
    res = {}
    for s in (0, 1):
        ms, o = timed(lambda: attn(q, k, v), 1)
        res[f"full_step_{s}_ms"] = round(ms, 3)
    o_full = o
    attn(q, k, v)                                   # first sparse step: allocations
    ms, o = timed(lambda: attn(q, k, v), n_sparse)
    res["sparse_step_ms"] = round(ms, 3)
    return res, o_full, o


def one_shape(name, mods, H, N, attn_cfg, static=None, n_sparse=8):
    import chipmunk_b200 as cm
    from chipmunk_b200.util.config import reset_to_defaults
    from chipmunk_b200.util import layer_counter as our_lc
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1)
    q, k, v = (torch.randn(1, H, N, 128, generator=g).to(torch.bfloat16).to(dev) for _ in range(3))
    out = {"shape": name, "heads": H, "seq": N}

    # ---- the reference's modules over the B200 operators
    cfg, lc = R.fresh(mods)
    cfg["attn"].update(attn_cfg)
    cfg["attn"]["full_step_schedule"] = {0, 1}
    A = mods["chipmunk.modules.attn"]
    A.singleton_static_mask = A.singleton_video_query_groups = None

    def make_ref():
        layer, counter = lc.LayerCounter.build_for_layer(is_attn_sparse=True)
        a = A.SparseDiffAttn(layer, counter)
        if static:
            a.initialize_static_mask(static[0], static[1], H, dev)
        return a
    torch.manual_seed(0)
    ref_res, ref_full, ref_sparse = run_side(make_ref, cfg, q, k, v, n_sparse)
    out["reference_python"] = ref_res
    A.singleton_static_mask = A.singleton_video_query_groups = None
    torch.cuda.empty_cache()

    # ---- chipmunk_b200's modules
    reset_to_defaults()
    our_lc.singleton.__init__(0, 0)
    ocfg = cm.util.GLOBAL_CONFIG
    ocfg["steps"] = 50
    ocfg["attn"].update(attn_cfg)
    ocfg["attn"]["full_step_schedule"] = {0, 1}
    ocfg["attn"]["random_columns"] = 0.0           # no random columns on either side, so that the outputs are comparable
    import chipmunk_b200.modules.attn as OA
    for n in ("singleton_static_mask", "singleton_video_query_groups", "singleton_static_words", "singleton_group_flags"):
        setattr(OA, n, None)

    def make_ours():
        layer, counter = cm.LayerCounter.build_for_layer(is_attn_sparse=True)
        a = cm.SparseDiffAttn(layer, counter)
        if static:
            a.initialize_static_mask(static[0], static[1], H, dev)
        return a
    torch.manual_seed(0)
    our_res, our_full, our_sparse = run_side(make_ours, ocfg, q, k, v, n_sparse)
    out["chipmunk_b200_modules"] = our_res

    def rel(a, b):
        return float((a.float() - b.float()).norm() / b.float().norm())
    out["full_step_output_rel_diff"] = round(rel(our_full, ref_full), 6)
    out["sparse_step_output_rel_diff"] = round(rel(our_sparse, ref_sparse), 6)
    out["sparse_step_speedup_of_the_module_fusions"] = round(ref_res["sparse_step_ms"] / our_res["sparse_step_ms"], 3)
    out["full_step_speedup_of_the_module_fusions"] = round(ref_res["full_step_1_ms"] / our_res["full_step_1_ms"], 3)
    reset_to_defaults()
    return out


def main():
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    mods = R.install_on_b200()
    sys.modules["chipmunk.ops.attn"].torch = R._TorchWithZeroedEmpty()
    # the reference draws its 1 % random columns with torch.randint(0, 100, ...) == 0: same-sized tensor, no column set
    real_randint = torch.randint
    torch.randint = lambda lo, hi, shape, **kw: (torch.ones(shape, dtype=kw.get("dtype", torch.int64), device=kw.get("device"))
                                                 if hi == 100 else real_randint(lo, hi, shape, **kw))
    results = []
    # FLUX single-stream block (BASELINE configs[1]): plain top-k lists in multiples of 112, in-place accumulation ops
    flux = dict(is_enabled=True, first_n_dense_layers=0, top_keys=0.165, random_keys=0.0, local_voxels=0, local_1d_window=0,
                recompute_mask=False, should_compress_indices=False, counts_multiple_of=112, pad_qkv_before_kernel=False)
    results.append(one_shape("flux_block_attention", mods, 24, 4608, flux))
    print(json.dumps(results[-1]), flush=True)
    # HunyuanVideo flow (bit-packed mask, padded ops) on 4 heads of a 33-frame 544x960 clip: 9 x 34 x 60 latents + 256 text tokens
    hy = dict(is_enabled=True, first_n_dense_layers=0, top_keys=0.05, random_keys=0.0, local_voxels=0, local_1d_window=0,
              recompute_mask=True, should_compress_indices=True, counts_multiple_of=128, pad_qkv_before_kernel=True)
    results.append(one_shape("hunyuan_540p_33f_4_heads", mods, 4, 9 * 34 * 60 + 256, hy, static=((9, 34, 60), 256), n_sparse=4))
    print(json.dumps(results[-1]), flush=True)
    if out_path:
        with open(out_path, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
