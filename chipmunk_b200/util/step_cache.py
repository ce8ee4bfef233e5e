"""Step caching glue: on the steps of `step_caching.skip_step_schedule` the whole transformer stack is skipped and the
previous step's image-token output is reused (reference examples/hunyuan/hyvideo/modules/models.py:732-741,834-835 and
the same pattern in the FLUX example).  The reference spells this out inside each model's forward; here it is one
object a model owns, so that the FLUX / HunyuanVideo forwards reduce to

    cached = self.step_cache.try_skip(inference_step, counter)     # -> tensor or None
    if cached is not None: return finish(cached)
    ... run the blocks ...
    self.step_cache.store(img)

with the reference's side effect reproduced: a skipped step still advances the singleton layer counter's inference
step, because none of the SparseDiff modules run (and therefore none of them increments it).
"""
from __future__ import annotations

from typing import Optional

import torch

from .config import GLOBAL_CONFIG
from .layer_counter import LayerCounter


class StepCache:
    def __init__(self):
        self._value: Optional[torch.Tensor] = None

    @staticmethod
    def is_enabled() -> bool:
        return bool(GLOBAL_CONFIG["step_caching"]["is_enabled"])

    @staticmethod
    def should_skip(inference_step: int) -> bool:
        cfg = GLOBAL_CONFIG["step_caching"]
        return bool(cfg["is_enabled"]) and inference_step in cfg["skip_step_schedule"]

    def try_skip(self, inference_step: int, layer_counter: Optional[LayerCounter] = None) -> Optional[torch.Tensor]:
        """The cached output if this step is skipped (and the counter advanced), else None."""
        if not self.should_skip(inference_step):
            return None
        if self._value is None:
            raise RuntimeError(f"step {inference_step} is in skip_step_schedule but no earlier step has been stored")
        if layer_counter is not None:
            layer_counter.cur_inference_step += 1
        return self._value

    def store(self, img: torch.Tensor) -> None:
        """Keep this step's output for the skipped steps that follow (reference models.py:834-835: `img.clone()`).
        The buffer is reused from step to step: one allocation per generation instead of one per computed step."""
        if not self.is_enabled():
            return
        if self._value is None or self._value.shape != img.shape or self._value.dtype != img.dtype or self._value.device != img.device:
            self._value = torch.empty_like(img)
        self._value.copy_(img)

    def reset(self) -> None:
        self._value = None
