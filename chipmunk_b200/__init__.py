"""chipmunk_b200 — Blackwell-native column-sparse DiT inference path.

Importing the package loads libchipmunk_b200.so (built in-tree by `python chipmunk_b200/build.py`)
and registers the `torch.ops.chipmunk.*` operators.  The Python surface mirrors the reference's
`chipmunk` package: `ops`, `modules`, `util`.
"""
from . import _lib, torch_ops

torch_ops.register()

from . import util, ops, modules  # noqa: E402
from .modules import SparseDiffAttn, SparseDiffMlp  # noqa: E402
from .util import GLOBAL_CONFIG, LayerCounter  # noqa: E402

__all__ = ["ops", "modules", "util", "SparseDiffAttn", "SparseDiffMlp", "GLOBAL_CONFIG", "LayerCounter"]
__version__ = "0.1.0"
