"""Drop-in alias: `import chipmunk` resolves to the Blackwell-native package.

The reference's examples do `import chipmunk`, `from chipmunk.modules import SparseDiffAttn`,
`from chipmunk.util import GLOBAL_CONFIG, LayerCounter`, `chipmunk.util.config.load_from_file(...)`
and `import chipmunk.ops` (examples/flux/src/flux/modules/layers.py, examples/hunyuan/hyvideo/...).
Putting this directory on PYTHONPATH ahead of (or instead of) the reference's `src/` makes those
imports land on chipmunk_b200 with no change to the model code.
"""
import sys
import types

import chipmunk_b200 as _impl
from chipmunk_b200 import modules, ops, util  # noqa: F401
from chipmunk_b200.modules import SparseDiffAttn, SparseDiffMlp, quantize_fp8  # noqa: F401
from chipmunk_b200.util import GLOBAL_CONFIG, LayerCounter  # noqa: F401
# the reference's own voxel tests (src/chipmunk/tests/test_voxel.py:4-10, examples/hunyuan/.../test_chipmunk.py) import these
# from the top level of the package
from chipmunk_b200.ops.voxel import (get_local_indices_with_text, get_local_voxel_indices, masktoinds,  # noqa: F401
                                     reverse_voxel_chunk_no_padding, voxel_chunk_no_padding)

# submodules are taken from sys.modules, not as attributes: `chipmunk_b200.ops.mlp` the ATTRIBUTE is the run_e2e function
# (`from .mlp import run_e2e as mlp`, as in the reference's ops/__init__.py), the module is sys.modules[...]
_SUBMODULES = (
    "ops", "modules", "util",
    "ops.attn", "ops.mlp", "ops.indexed_io", "ops.bitpack", "ops.patch", "ops.voxel",
    "modules.attn", "modules.mlp",
    "util.config", "util.layer_counter", "util.step_cache",
    "util.storage", "util.storage.offloaded_tensor", "util.storage.layer_storage",
)
_alias = {f"chipmunk.{_n}": sys.modules[f"chipmunk_b200.{_n}"] for _n in _SUBMODULES}
for _name, _mod in _alias.items():
    sys.modules.setdefault(_name, _mod)

# `chipmunk.cuda` was the compiled extension (importing it only registered torch.ops.chipmunk.*);
# `chipmunk.triton` exported the Triton mm2 kernel and its raw CUfunction.  Both exist here as
# empty shells so that `from chipmunk import cuda, triton` keeps working; the operators are
# already registered by chipmunk_b200 and no Triton kernel is compiled at import time.
cuda = types.ModuleType("chipmunk.cuda")
triton = types.ModuleType("chipmunk.triton")
triton.csp_mlp_mm2_function_ptr = 0
sys.modules.setdefault("chipmunk.cuda", cuda)
sys.modules.setdefault("chipmunk.triton", triton)
