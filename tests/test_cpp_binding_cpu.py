"""INTEGRATION.md option 3, checked: bindings/chipmunk_ops_b200.cpp defines the ten `chipmunk::*` operator functions that the
reference's binding TU (csrc/chipmunk.cpp:27-43) declares `extern`, on top of include/chipmunk_b200.h.  It must compile, and --
where the reference is mounted -- the reference's own, UNMODIFIED csrc/chipmunk.cpp built together with it and linked against
libchipmunk_b200.so must load as the `cuda` extension module and register the reference's ten schemas for the CUDA backend.
CPU only (g++; no device code on that side of the C ABI); the calls themselves need a GPU and are not exercised here."""
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TU = "/root/reference/csrc/chipmunk.cpp"


def _builder():
    spec = importlib.util.spec_from_file_location("_cm_build_binding", os.path.join(ROOT, "bindings", "build_binding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_binding_tu_compiles_against_the_header(cm):
    _builder().compile_check()


def test_binding_calls_every_operator_entry_point_of_the_abi():
    """Each of the reference's ten operators maps onto its `cm_*` function (the B200-only entry points -- the fused add-back,
    the fused gathers, select_columns, gather_rows, the bit codec -- have no counterpart in the reference's binding TU)."""
    src = open(os.path.join(ROOT, "bindings", "chipmunk_ops_b200.cpp")).read()
    for fn in ("cm_csp_attn(", "cm_dense_attn_strided(", "cm_csp_mlp_mm1(", "cm_csp_mlp_mm2(", "cm_csp_scatter_add(",
               "cm_copy_indices(", "cm_topk_indices(", "cm_mask_to_indices(", "cm_strerror("):
        assert fn in src, fn
    for op in ("void csp_attn(", "at::Tensor csp_128_attn(", "std::vector<at::Tensor> dense_attn(",
               "std::vector<at::Tensor> dense_colsum_attn(", "void csp_mlp_mm1(", "void csp_mlp_mm2_and_scatter_add(",
               "void csp_scatter_add(", "void copy_indices(", "void topk_indices(", "std::vector<at::Tensor> mask_to_indices("):
        assert op in src, op


@pytest.mark.skipif(not os.path.exists(REF_TU), reason="the reference is only mounted in the authoring container")
def test_reference_binding_tu_builds_and_registers_over_the_b200_abi(cm):
    lib = _builder().build()
    assert lib and os.path.exists(lib)
    code = r'''
import importlib.machinery, importlib.util, sys, torch
loader = importlib.machinery.ExtensionFileLoader("cuda", sys.argv[1])
mod = importlib.util.module_from_spec(importlib.util.spec_from_loader("cuda", loader)); loader.exec_module(mod)
want = {
 "csp_mlp_mm1": "chipmunk::csp_mlp_mm1(Tensor a, Tensor b_colmajor, Tensor(c!) c, Tensor bias, Tensor pa_cache_colmajor, Tensor indices, Tensor indices_counts) -> ()",
 "csp_attn": "chipmunk::csp_attn(Tensor q, Tensor k, Tensor v, Tensor o, Tensor indices, Tensor indices_counts, int o_scale) -> ()",
 "csp_128_attn": "chipmunk::csp_128_attn(Tensor q, Tensor k, Tensor v, Tensor indices, Tensor indices_counts) -> Tensor",
 "dense_attn": "chipmunk::dense_attn(Tensor q, Tensor k, Tensor v) -> Tensor[]",
 "dense_colsum_attn": "chipmunk::dense_colsum_attn(Tensor q, Tensor k, Tensor v, Tensor p) -> Tensor[]",
 "mask_to_indices": "chipmunk::mask_to_indices(Tensor mask, int multiple_of, int pad_to_multiple_of) -> Tensor[]",
}
for name in ("csp_attn", "csp_128_attn", "dense_attn", "dense_colsum_attn", "csp_mlp_mm1", "csp_mlp_mm2_and_scatter_add",
             "csp_scatter_add", "copy_indices", "topk_indices", "mask_to_indices"):
    op = getattr(torch.ops.chipmunk, name)
    if name in want:
        assert str(op.default._schema) == want[name], str(op.default._schema)
    assert torch._C._dispatch_has_kernel_for_dispatch_key(f"chipmunk::{name}", "CUDA"), name
    assert not torch._C._dispatch_has_kernel_for_dispatch_key(f"chipmunk::{name}", "CPU"), name
q = torch.zeros(1, 1, 192, 128, dtype=torch.bfloat16)
try:
    torch.ops.chipmunk.csp_128_attn(q, q, q, torch.zeros(1, 1, 1, 192, dtype=torch.int32), torch.zeros(1, 1, 1, dtype=torch.int32))
    raise SystemExit("a CPU call went through")
except NotImplementedError:
    pass
print("BINDING OK")
'''
    r = subprocess.run([sys.executable, "-c", code, lib], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0 and "BINDING OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
