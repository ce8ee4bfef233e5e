"""(inference step, model invocation, layer, submodule) odometer shared by every sparse module.
Same public behaviour as src/chipmunk/util/layer_counter.py:3-70."""
from __future__ import annotations

from .config import GLOBAL_CONFIG


class LayerCounter:
    def __init__(self, num_layers: int, num_sparse_submodules_per_layer: int):
        self.num_layers = num_layers
        self.num_submodules_per_layer = num_sparse_submodules_per_layer
        self.has_mlp_sparsity = False
        self.has_attn_sparsity = False
        self.reset()

    # ---- construction helper used by the model code, one call per transformer block
    @staticmethod
    def build_for_layer(is_mlp_sparse: bool = False, is_attn_sparse: bool = False):
        s = singleton
        layer_num = s.num_layers
        s.num_layers += 1
        if is_attn_sparse and not s.has_attn_sparsity:
            s.has_attn_sparsity = True
            s.num_submodules_per_layer += 1
        if is_mlp_sparse and not s.has_mlp_sparsity:
            s.has_mlp_sparsity = True
            s.num_submodules_per_layer += 1
        return layer_num, s

    # ---- schedule queries
    def should_do_full_mlp_step(self) -> bool:
        return self.cur_inference_step % GLOBAL_CONFIG["mlp"]["full_step_every"] == 0

    def should_do_full_attn_step(self) -> bool:
        schedule = GLOBAL_CONFIG["attn"]["full_step_schedule"]
        if schedule is not None:
            return self.cur_inference_step in schedule
        step = self.cur_inference_step
        return step < 2 or step % GLOBAL_CONFIG["attn"]["full_step_every"] == 0

    # ---- odometer
    def increment(self):
        coord = (self.cur_inference_step, self.cur_layer, self.cur_layer_submodule)
        invocations = GLOBAL_CONFIG["num_model_invocations_per_inference_step"]
        self.cur_layer_submodule += 1
        if self.cur_layer_submodule == self.num_submodules_per_layer:
            self.cur_layer_submodule = 0
            self.cur_layer += 1
            if self.cur_layer == self.num_layers:
                self.cur_layer = 0
                self.cur_model_invocation_per_step += 1
                if self.cur_model_invocation_per_step == invocations:
                    self.cur_model_invocation_per_step = 0
                    self.cur_inference_step += 1
        # the reference rewinds when the NEXT coordinate is the very last one of the generation
        at_end = (self.cur_inference_step == GLOBAL_CONFIG["steps"] - 1
                  and self.cur_layer == self.num_layers - 1
                  and self.cur_layer_submodule == self.num_submodules_per_layer - 1
                  and self.cur_model_invocation_per_step == invocations - 1)
        if at_end:
            self.reset()
        return coord

    def reset(self) -> None:
        self.cur_inference_step = 0
        self.cur_model_invocation_per_step = 0
        self.cur_layer = 0
        self.cur_layer_submodule = 0

    def get_cur_coord(self):
        return (self.cur_inference_step, self.cur_layer, self.cur_layer_submodule)


singleton = LayerCounter(0, 0)
