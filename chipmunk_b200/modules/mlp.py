"""SparseDiffMlp: dense MLP on full steps, column-sparse delta MLP otherwise.

Behavioural mirror of src/chipmunk/modules/mlp.py:8-123: on a full step cache gelu(fc1 x)^T,
the output and the per-128-token block means of fc1 x; on a sparse step pick, per token block,
the neurons whose block-mean pre-activation moved most (topk_indices), and recompute only
those columns as a delta on the cached output (ops.mlp).
"""
from __future__ import annotations

import torch

from .. import ops
from ..util import GLOBAL_CONFIG, LayerCounter, MlpStorage


def block_mean(x: torch.Tensor, mbm: int) -> torch.Tensor:
    """[b, mb*mbm, c] -> [b, mb, c] mean over each block of mbm tokens."""
    b, n, c = x.shape
    return x.reshape(b, n // mbm, mbm, c).mean(dim=2)


class SparseDiffMlp:
    def __init__(self, layer_num: int, layer_counter: LayerCounter, fc1: torch.nn.Linear,
                 activation: torch.nn.Module, fc2: torch.nn.Linear, heuristic_sms_scatter_add: int = 6):
        # lists keep the Linear modules out of any parent nn.Module's parameter registry
        self.fc1 = [fc1]
        self.fc2 = [fc2]
        self.fc2w_T = [fc2.weight.data.transpose(0, 1).contiguous()]
        self.layer_counter = layer_counter
        self.activation = activation
        self.storage = MlpStorage(layer_num)
        self.num_sms_scatter_add = heuristic_sms_scatter_add    # accepted for API parity; unused

    def _dense(self, x):
        return self.fc2[0](self.activation(self.fc1[0](x)))

    def _refresh_indices(self, x, cfg):
        fc1 = self.fc1[0]
        mbm, bm = cfg["mbm"], cfg["bm"]
        bmfc1 = fc1(block_mean(x, mbm))
        cache = self.storage.get_blockmean_mid_cache()
        mdiff = (bmfc1 - cache).abs()
        r = bm // mbm
        if r != 1:
            b, n, f = mdiff.shape
            mdiff = mdiff.reshape(b, n // r, r, f).sum(dim=2)
        mdiff = mdiff.contiguous()
        inds = torch.empty_like(mdiff, dtype=torch.int32)
        counts = torch.empty(mdiff.shape[:2], dtype=torch.int32, device=x.device)
        ops.topk_indices(mdiff, inds, counts, 1 - cfg["top_keys"], cfg["counts_multiple_of"], cfg["random_keys"])
        ops.copy_indices(bmfc1, cache, inds, counts)
        self.storage.set_indices(inds)
        self.storage.set_counts(counts)

    def forward(self, x: torch.Tensor):
        cfg = GLOBAL_CONFIG["mlp"]
        if not cfg["is_enabled"]:
            return self._dense(x)
        do_full = self.layer_counter.should_do_full_mlp_step()
        step, layer, _ = self.layer_counter.increment()
        assert x.ndim == 3 and x.shape[0] == 1, "x must be (1, N, C)"
        if layer < cfg["first_n_dense_layers"]:
            return self._dense(x)

        fc1, fc2 = self.fc1[0], self.fc2[0]
        if do_full:
            mid = fc1(x)
            act = self.activation(mid)
            out = fc2(act)
            self.storage.set_sparse_act_T(act.transpose(-1, -2).contiguous())
            self.storage.set_out_cache(out)
            self.storage.set_blockmean_mid_cache(block_mean(mid, cfg["mbm"]))
            return out

        reuse = (step % cfg["block_mask_cache"] != 0 and self.storage.get_indices() is not None and step >= 10)
        if not reuse:
            self._refresh_indices(x, cfg)

        out_cache = self.storage.get_out_cache()[0]
        ops.mlp(x=x[0], fc1w=fc1.weight.data, fc1b=fc1.bias.data, fc2w_T=self.fc2w_T[0],
                indices=self.storage.get_indices()[0], counts=self.storage.get_counts()[0],
                sparse_act_T=self.storage.get_sparse_act_T()[0], cached_out=out_cache,
                num_sms_scatter_add=self.num_sms_scatter_add)
        out_cache = out_cache.unsqueeze(0)
        self.storage.set_out_cache(out_cache)
        return out_cache

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)
