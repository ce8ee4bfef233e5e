"""The drop-in boundary at the Python level (SURVEY §8b "what calls it"): every `chipmunk` import the reference's FLUX,
HunyuanVideo and Wan example code makes must resolve against the alias package, to the B200 implementation.  The list
below is those import statements (examples/flux/src/flux/{model,sampling,util,cli}.py and modules/layers.py,
examples/hunyuan/{sample_video.py,hyvideo/inference.py,hyvideo/modules/models.py}, examples/wan/{generate.py,
wan/modules/model.py}, and the reference's own tests); where the reference is mounted the statements are also collected
from its sources, so a new import there cannot go unnoticed.  CPU only."""
import ast
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXAMPLE_IMPORTS = """
from chipmunk.modules import SparseDiffMlp, SparseDiffAttn
from chipmunk.modules import quantize_fp8
from chipmunk.modules.attn import SparseDiffAttn
from chipmunk.modules.mlp import SparseDiffMlp, block_mean
from chipmunk.util import LayerCounter, GLOBAL_CONFIG
from chipmunk.util import AttnStorage, MlpStorage, MaybeOffloadedTensor
from chipmunk.util.config import GLOBAL_CONFIG, load_from_file, update_global_config
from chipmunk.util.layer_counter import LayerCounter
from chipmunk.util.storage import AttnStorage, MlpStorage, LayerStorage, MaybeOffloadedTensor
from chipmunk.util.storage.offloaded_tensor import PIPELINE_DEPTH, MaybeOffloadedTensor
from chipmunk.util.storage.layer_storage import AttnStorage, MlpStorage, LayerStorage
from chipmunk.ops import patchify_rope, patchify, unpatchify
from chipmunk.ops import mlp, copy_indices, topk_indices, mask_to_indices, scatter_add, csp_attn, dense_attn, dense_colsum_attn, bitpack, bitunpack
from chipmunk.ops.voxel import voxel_chunk_no_padding, reverse_voxel_chunk_no_padding
from chipmunk.ops.voxel import get_local_voxel_indices, get_local_indices_with_text, masktoinds, merge_indices, offsets
from chipmunk.ops.mlp import mm1, mm2_fused, mm2_unfused, run_e2e
from chipmunk.ops.attn import csp_attn, dense_attn, dense_colsum_attn
from chipmunk.ops.indexed_io import copy_indices, topk_indices, scatter_add, mask_to_indices
from chipmunk.ops.bitpack import bitpack, bitunpack
from chipmunk import cuda, triton, ops, util
import chipmunk.util.config
import chipmunk.ops
import chipmunk
""".strip().splitlines()


def _run(statements):
    """Execute the import statements in a fresh interpreter whose only `chipmunk` is this repo's alias package; print
    the module each imported name comes from."""
    code = ["import sys", f"sys.path.insert(0, {ROOT!r})", "import types"]
    for s in statements:
        code.append(s)
    code.append("import chipmunk, chipmunk_b200")
    code.append("assert chipmunk.ops is chipmunk_b200.ops and chipmunk.util is chipmunk_b200.util")
    code.append("assert chipmunk.util.config.GLOBAL_CONFIG is chipmunk_b200.util.GLOBAL_CONFIG")
    code.append("import torch; assert hasattr(torch.ops.chipmunk, 'csp_attn')")
    code.append("print('OK')")
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, "-c", "\n".join(code)], capture_output=True, text=True, env=env, cwd="/tmp")


def test_example_imports_resolve_to_the_b200_package():
    r = _run(EXAMPLE_IMPORTS)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr[-2000:]


def test_fp8_preview_fails_loudly(cm):
    import torch
    from chipmunk_b200.modules import quantize_fp8
    with pytest.raises(RuntimeError, match="is_fp8"):
        quantize_fp8(torch.nn.Linear(4, 4))


def _chipmunk_imports_in(path):
    try:
        tree = ast.parse(open(path, encoding="utf-8", errors="ignore").read())
    except SyntaxError:
        return []
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.level == 0 and node.module and (node.module == "chipmunk" or node.module.startswith("chipmunk.")):
            names = ", ".join(a.name for a in node.names)
            out.append(f"from {node.module} import {names}")
        elif isinstance(node, ast.Import):
            for a in node.names:
                if a.name == "chipmunk" or a.name.startswith("chipmunk."):
                    out.append(f"import {a.name}")
    return out


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="the reference is only mounted in the authoring container")
def test_every_chipmunk_import_of_the_reference_examples_resolves():
    """Collected from the sources: every absolute `chipmunk` import in the reference's examples (FLUX, HunyuanVideo, Wan,
    Mochi where present) and in its own Python tests.  Only what is documented as out of scope may be missing: the
    Triton kernels' own modules (`chipmunk.triton.*` functions) and the fp8 preview internals."""
    files = [f for pat in ("/root/reference/examples/**/*.py", "/root/reference/src/chipmunk/tests/*.py")
             for f in glob.glob(pat, recursive=True)]
    stmts = sorted({s for f in files for s in _chipmunk_imports_in(f)})
    assert len(stmts) >= 10, stmts
    out_of_scope = ("chipmunk.triton.", "chipmunk.modules.mlp_fp8", "chipmunk.cuda.")
    stmts = [s for s in stmts if not any(o in s for o in out_of_scope)]
    # two stale test files do `from chipmunk import get_local_voxel_indices, ...`: names the reference's own
    # chipmunk/__init__.py (`__all__ = ['cuda', 'ops', 'triton', 'util']`) does not export either
    top_level = {"cuda", "ops", "triton", "util"}
    stmts = [s for s in stmts if not (s.startswith("from chipmunk import ")
                                      and not {n.strip() for n in s[len("from chipmunk import "):].split(",")} <= top_level)]
    r = _run(stmts)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stderr[-2000:], stmts)
    missing = [s for s in stmts if s not in EXAMPLE_IMPORTS and not s.startswith("import ")]
    # the committed list above must stay a superset of what the examples do (name by name)
    have = {}
    for s in EXAMPLE_IMPORTS:
        if s.startswith("from "):
            mod, names = s[5:].split(" import ")
            have.setdefault(mod, set()).update(n.strip() for n in names.split(","))
    for s in missing:
        mod, names = s[5:].split(" import ")
        assert {n.strip() for n in names.split(",")} <= have.get(mod, set()), f"EXAMPLE_IMPORTS lacks: {s}"


@pytest.mark.skipif(not os.path.exists("/root/reference/src/chipmunk/tests/test_voxel.py"),
                    reason="the reference is only mounted in the authoring container")
def test_the_references_own_voxel_tests_pass_against_this_package():
    """src/chipmunk/tests/test_voxel.py, unmodified, with `chipmunk` = this repo's alias package: its assertions (voxel order of
    the first chunks, round trips at toy and HunyuanVideo shapes incl. ragged tails) hold for the permutation-based
    implementation, and its mask / index-table calls run."""
    code = r'''
import importlib.util, sys
sys.path.insert(0, sys.argv[1])
spec = importlib.util.spec_from_file_location("ref_test_voxel", "/root/reference/src/chipmunk/tests/test_voxel.py")
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
import chipmunk_b200.ops.voxel as ours
assert mod.voxel_chunk_no_padding is ours.voxel_chunk_no_padding and mod.get_local_indices_with_text is ours.get_local_indices_with_text
ran = 0
for name in sorted(dir(mod)):
    if name.startswith("test_") and callable(getattr(mod, name)):
        getattr(mod, name)(); ran += 1
print("RAN", ran)
'''
    r = subprocess.run([sys.executable, "-c", code, ROOT], capture_output=True, text=True, cwd="/tmp", timeout=900,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stderr[-3000:]
    assert "RAN 6" in r.stdout, r.stdout[-500:]
