"""Inputs shared by make_golden_control.py (reference side) and tests/test_control_golden.py (this repo's side):
hand-written YAML configs in the styles of the reference's example files, and odometer cases."""

CONFIG_YAMLS = {
    # HunyuanVideo style: a `!!set` full-step schedule, partial sections, offloading keys without the global switch
    "video_schedule": """
patchify:
  is_enabled: true
mlp:
  is_enabled: false
attn:
  is_enabled: true
  top_keys: 0.05
  random_keys: 0.01
  local_voxels: 0
  first_n_dense_layers: 2
  recompute_mask: true
  should_compress_indices: true
  full_step_schedule: !!set
    ? 0
    ? 1
    ? 10
    ? 40
  pad_qkv_before_kernel: true
  counts_multiple_of: 128
offloading:
  attn.out_cache: true
  attn.indices: true
  text_encoders: true
step_caching:
  is_enabled: true
  skip_step_schedule: !!set
    ? 7
    ? 11
    ? 13
""",
    # Wan style: two model invocations per step, local voxels, periodic full steps, the global offloading switch
    "video_cfg_two_invocations": """
num_model_invocations_per_inference_step: 2
patchify:
  is_enabled: false
mlp:
  is_enabled: false
attn:
  top_keys: 0.1
  local_voxels: 5
  full_step_every: 10
offloading:
  global_disable_offloading: false
  attn.counts: true
""",
    # FLUX style: both modules on, plain index lists in multiples of 112, a null schedule, everything resident
    "image_block": """
mlp:
  is_enabled: true
  is_fp8: false
  top_keys: 0.3
  random_keys: 0.05
  full_step_every: 10
  block_mask_cache: 2
  counts_multiple_of: 256
  bm: 128
  mbm: 128
patchify:
  is_enabled: true
  chunk_size_1: 8
  chunk_size_2: 4
attn:
  is_enabled: true
  top_keys: 0.165
  full_step_every: 10
  full_step_schedule: ~
  recompute_mask: false
  should_compress_indices: false
  counts_multiple_of: 112
  pad_qkv_before_kernel: false
  mbm: 192
offloading:
  global_disable_offloading: true
  attn.out_cache: false
  attn.indices: false
  text_encoders: false
""",
    # scalars at the top level, an unknown key and an unknown section (kept verbatim by the deep merge), an empty section
    "odd_keys": """
steps: 28
world_size: 8
generation_index: 3
custom_top_level: hello
custom_section:
  a: 1
  b: [1, 2, 3]
attn:
  experimental_knob: 0.5
step_caching:
  is_enabled: false
""",
    "empty_file": "",
}

# (steps, invocations per step, layers, sparse submodules per layer, attn schedule or None, attn every, mlp every)
COUNTER_CASES = {
    "flux_like": dict(steps=6, invocations=1, layers=3, subs=2, schedule=None, attn_every=4, mlp_every=3, generations=2),
    "hunyuan_like": dict(steps=12, invocations=1, layers=2, subs=1, schedule=[0, 1, 10], attn_every=10, mlp_every=10, generations=2),
    "wan_like_two_invocations": dict(steps=5, invocations=2, layers=2, subs=1, schedule=None, attn_every=3, mlp_every=2, generations=3),
    "single_everything": dict(steps=3, invocations=1, layers=1, subs=1, schedule=None, attn_every=10, mlp_every=10, generations=3),
    "empty_schedule": dict(steps=4, invocations=1, layers=2, subs=2, schedule=[], attn_every=1, mlp_every=1, generations=1),
}

# build_for_layer call sequences: (is_mlp_sparse, is_attn_sparse) per transformer block
BUILD_CASES = {
    "flux_like": [(True, True)] * 4,
    "attn_only": [(False, True)] * 3,
    "mixed": [(False, False), (False, True), (True, False), (True, True)],
}


def jsonable(x):
    """GLOBAL_CONFIG with sets as sorted lists (and tuples as lists), recursively."""
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (set, frozenset)):
        return {"__set__": sorted(x)}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    return x


def run_counter_case(case, global_config, counter_cls):
    """Drive one odometer over `generations` generations' worth of increment() calls.  Each record is
    [full_attn, full_mlp, step, layer, submodule, invocation_before] (the three coordinates as increment() returns them)."""
    global_config["steps"] = case["steps"]
    global_config["num_model_invocations_per_inference_step"] = case["invocations"]
    global_config["attn"]["full_step_every"] = case["attn_every"]
    global_config["attn"]["full_step_schedule"] = None if case["schedule"] is None else set(case["schedule"])
    global_config["mlp"]["full_step_every"] = case["mlp_every"]
    c = counter_cls(case["layers"], case["subs"])
    out = []
    calls = case["generations"] * case["steps"] * case["invocations"] * case["layers"] * case["subs"]
    for _ in range(calls):
        fa, fm = bool(c.should_do_full_attn_step()), bool(c.should_do_full_mlp_step())
        inv = c.cur_model_invocation_per_step
        step, layer, sub = c.increment()
        assert (step, layer, sub) != () and c.get_cur_coord() == (c.cur_inference_step, c.cur_layer, c.cur_layer_submodule)
        out.append([fa, fm, step, layer, sub, inv])
    return out


def run_build_case(flags, layer_counter_module):
    """LayerCounter.build_for_layer over a fresh module singleton: returned layer numbers and the singleton's totals."""
    s = layer_counter_module.singleton
    s.num_layers, s.num_submodules_per_layer, s.has_mlp_sparsity, s.has_attn_sparsity = 0, 0, False, False
    s.reset()
    nums = []
    for is_mlp, is_attn in flags:
        n, got = layer_counter_module.LayerCounter.build_for_layer(is_mlp_sparse=is_mlp, is_attn_sparse=is_attn)
        assert got is s
        nums.append(n)
    res = {"layer_nums": nums, "num_layers": s.num_layers, "subs": s.num_submodules_per_layer,
           "has_mlp": s.has_mlp_sparsity, "has_attn": s.has_attn_sparsity}
    s.num_layers, s.num_submodules_per_layer, s.has_mlp_sparsity, s.has_attn_sparsity = 0, 0, False, False
    return res
