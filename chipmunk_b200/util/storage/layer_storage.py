"""Named cache slots of one layer: `MlpStorage`, `AttnStorage`, `LayerStorage` with the get_/set_ accessors and the
`load_async / load_async_wait / complete_cur_layer` calls of src/chipmunk/util/storage/layer_storage.py:5-210, generated
from a table of slot names instead of written out per slot."""
from __future__ import annotations

import torch

from .offloaded_tensor import MaybeOffloadedTensor


class _Slots:
    """Attribute bag of MaybeOffloadedTensor with generated get_/set_ accessors."""

    _prefix = ""
    _names: tuple = ()

    def __init__(self, layer_num: int):
        self.layer_num = layer_num
        for n in self._names:
            setattr(self, n, None)

    def _get(self, n):
        slot = getattr(self, n)
        return None if slot is None else slot.get_loaded_value()

    def _set(self, n, value: torch.Tensor):
        slot = getattr(self, n)
        if slot is None:
            slot = MaybeOffloadedTensor(f"{self._prefix}.{n}", self.layer_num, value.dtype, value.device)
            setattr(self, n, slot)
        slot.offload(value)

    def _each(self, names=None):
        for n in (names or self._names):
            slot = getattr(self, n)
            if slot is not None:
                yield slot

    def load_async(self):
        for s in self._each(self._load_names):
            s.load_async()

    def load_async_wait(self):
        for s in self._each(self._load_names):
            s.load_async_wait()

    def complete_cur_layer(self):
        for s in self._each(self._complete_names):
            s.complete_cur_layer()


def _accessors(cls):
    for n in cls._names:
        setattr(cls, f"get_{n}", (lambda self, _n=n: self._get(_n)))
        setattr(cls, f"set_{n}", (lambda self, value, _n=n: self._set(_n, value)))
    return cls


@_accessors
class MlpStorage(_Slots):
    """sparse_act_T [1,F,M], out_cache [1,M,N], indices [1,M/128,F], counts [1,M/128],
    blockmean_mid_cache [1,M/128,F]  (reference layer_storage.py:5-99)."""
    _prefix = "mlp"
    _names = ("sparse_act_T", "out_cache", "indices", "counts", "blockmean_mid_cache")
    _load_names = ("sparse_act_T", "out_cache", "indices", "counts")
    _complete_names = ("blockmean_mid_cache", "out_cache", "indices", "counts")


@_accessors
class AttnStorage(_Slots):
    """indices (bit-packed mask or int32 indices), counts, out_cache [B,H,N,128],
    lse_constants [B,H,N,1]  (reference layer_storage.py:101-189)."""
    _prefix = "attn"
    _names = ("indices", "counts", "out_cache", "lse_constants")
    _load_names = _names
    _complete_names = _names

    def __init__(self, layer_num: int, init_names=()):
        super().__init__(layer_num)
        # the reference pre-creates these two so that `storage.out_cache.is_offload_enabled`
        # can be read before the first set (modules/attn.py:186)
        dev = torch.device("cuda")
        if "out_cache" in init_names:
            self.out_cache = MaybeOffloadedTensor("attn.out_cache", layer_num, torch.bfloat16, dev)
        if "indices" in init_names:
            self.indices = MaybeOffloadedTensor("attn.indices", layer_num, torch.uint8, dev)


class LayerStorage:
    def __init__(self, layer_num: int):
        self.layer_num = layer_num
        self.mlp = MlpStorage(layer_num)
        self.attn = AttnStorage(layer_num)

    def load_async(self):
        self.mlp.load_async()
        self.attn.load_async()

    def load_async_wait(self):
        self.mlp.load_async_wait()
        self.attn.load_async_wait()
