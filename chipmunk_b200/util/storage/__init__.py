"""Per-layer cache slots and their residency (same module paths as the reference: `chipmunk.util.storage`,
`.offloaded_tensor` -- MaybeOffloadedTensor, PIPELINE_DEPTH -- and `.layer_storage` -- Attn/Mlp/LayerStorage;
src/chipmunk/util/storage/__init__.py).  The model loops import PIPELINE_DEPTH from `.offloaded_tensor`
(examples/flux/src/flux/model.py:7, examples/wan/wan/modules/model.py:13)."""
from .offloaded_tensor import MaybeOffloadedTensor, PIPELINE_DEPTH
from .layer_storage import AttnStorage, LayerStorage, MlpStorage

__all__ = ["MaybeOffloadedTensor", "LayerStorage", "MlpStorage", "AttnStorage", "PIPELINE_DEPTH"]
