"""State the sparse-delta modules read: config, step/layer coordinates, per-layer caches.
Mirrors the names exported by the reference's `chipmunk.util` (src/chipmunk/util/__init__.py)."""
from .config import GLOBAL_CONFIG, BASE_CONFIG, load_from_file, update_global_config
from .layer_counter import LayerCounter
from .storage import AttnStorage, MlpStorage, LayerStorage, MaybeOffloadedTensor, PIPELINE_DEPTH
from .step_cache import StepCache

__all__ = ["GLOBAL_CONFIG", "BASE_CONFIG", "load_from_file", "update_global_config", "LayerCounter",
           "AttnStorage", "MlpStorage", "LayerStorage", "MaybeOffloadedTensor", "PIPELINE_DEPTH", "StepCache"]
