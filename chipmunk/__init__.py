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
from chipmunk_b200.modules import SparseDiffAttn, SparseDiffMlp  # noqa: F401
from chipmunk_b200.util import GLOBAL_CONFIG, LayerCounter  # noqa: F401

_alias = {
    "chipmunk.ops": ops, "chipmunk.modules": modules, "chipmunk.util": util,
    "chipmunk.ops.attn": _impl.ops.attn, "chipmunk.ops.mlp": _impl.ops.mlp,
    "chipmunk.ops.indexed_io": _impl.ops.indexed_io, "chipmunk.ops.bitpack": _impl.ops.bitpack,
    "chipmunk.ops.patch": _impl.ops.patch, "chipmunk.ops.voxel": _impl.ops.voxel,
    "chipmunk.modules.attn": _impl.modules.attn, "chipmunk.modules.mlp": _impl.modules.mlp,
    "chipmunk.util.config": _impl.util.config, "chipmunk.util.layer_counter": _impl.util.layer_counter,
    "chipmunk.util.storage": _impl.util.storage,
}
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
