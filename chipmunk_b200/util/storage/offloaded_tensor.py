"""Per-layer cache slots for the sparse-delta modules.

Public surface mirrors src/chipmunk/util/storage/{layer_storage,offloaded_tensor}.py
(`AttnStorage`, `MlpStorage`, `LayerStorage`, `MaybeOffloadedTensor`, `PIPELINE_DEPTH`, and the
`load_async / load_async_wait / complete_cur_layer` calls the model loops make), but the
default residency is different: on B200 every cache stays in HBM (see util/config.py).  Offload
remains available behind the same config keys, with two backing stores: pinned host memory (the
reference's; right-sized buffers that grow on demand instead of up to 1.2 GB per tensor name) or --
`offloading.backing: peer` -- the HBM of another GPU of the NVSwitch domain, reached with peer copies
over NVLink (14x the bandwidth of the PCIe path: a 731 MB 720p cache moves in ~1 ms instead of ~13 ms, so
the one-layer-ahead prefetch of the model loop hides it completely).  Two lazily created copy streams;
importing this module never touches CUDA.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ..config import GLOBAL_CONFIG

# GPU slots kept alive per tensor name while offloading (layer L uses slot L % PIPELINE_DEPTH)
PIPELINE_DEPTH = 2

_streams: Dict[str, "torch.cuda.Stream"] = {}
_gpu_slots: Dict[str, List[Optional[torch.Tensor]]] = {}


def _stream(kind: str) -> "torch.cuda.Stream":
    if kind not in _streams:
        _streams[kind] = torch.cuda.Stream()
    return _streams[kind]


def _invocations() -> int:
    return GLOBAL_CONFIG["num_model_invocations_per_inference_step"]


class MaybeOffloadedTensor:
    """One named cache of one layer.  Resident mode: a list of GPU tensors, one per model
    invocation of a step (cond / uncond).  Offload mode: pinned host copies plus a
    PIPELINE_DEPTH-deep ring of GPU staging tensors shared by all layers of the same name."""

    # kept for callers that pass cpu_buf_size (sizes in bytes, reference offloaded_tensor.py:42-44)
    LARGE_BUF_SIZE = 1 * 32 * 150000 * 128 * 2
    MEDIUM_BUF_SIZE = 1 * 32 * 50000 * 128 * 2
    SMALL_BUF_SIZE = 1 * 32 * 15000 * 128 * 2

    def __init__(self, name: str, layer_num: int, dtype: torch.dtype, device, cpu_buf_size: int = 0):
        flags = GLOBAL_CONFIG["offloading"]
        if name not in flags:
            raise ValueError(f"Invalid tensor name: {name}. Expected one of: {list(flags.keys())}")
        self.name = name
        self.layer_num = layer_num
        self.dtype = dtype
        self.device = device
        self.is_offload_enabled = (not flags["global_disable_offloading"]) and bool(flags[name])
        # backing store of an offloaded cache: pinned host memory, or a peer GPU's HBM over NVLink
        self.backing_device = torch.device("cpu")
        if self.is_offload_enabled and flags.get("backing", "host") == "peer":
            peer = flags.get("peer_device")
            if peer is None:
                raise ValueError("offloading.backing == 'peer' needs offloading.peer_device (a CUDA device index)")
            self.backing_device = torch.device("cuda", int(peer))
        self.layer_key = layer_num % PIPELINE_DEPTH
        n = _invocations()
        self.gpu_tensor: List[Optional[torch.Tensor]] = [None] * n
        self.cpu_buf: List[Optional[torch.Tensor]] = [None] * n
        self.real_shape: List[Optional[torch.Size]] = [None] * n
        self.model_invocation_count = 0
        if self.is_offload_enabled:
            _gpu_slots.setdefault(name, [None] * PIPELINE_DEPTH)

    # ---- bookkeeping
    def complete_cur_layer(self) -> None:
        self.model_invocation_count += 1

    def get_cur_model_invocation_key(self) -> int:
        return self.model_invocation_count % _invocations()

    # ---- store
    @torch.compiler.disable
    def offload(self, gpu_tensor: torch.Tensor) -> None:
        key = self.get_cur_model_invocation_key()
        if not self.is_offload_enabled:
            self.gpu_tensor[key] = gpu_tensor
            return
        self.real_shape[key] = gpu_tensor.shape
        buf = self.cpu_buf[key]
        if buf is None or buf.numel() < gpu_tensor.numel() or buf.dtype != gpu_tensor.dtype:
            if self.backing_device.type == "cpu":
                buf = torch.empty(gpu_tensor.numel(), dtype=gpu_tensor.dtype, device="cpu", pin_memory=True)
            else:   # a neighbour GPU's HBM: cudaMemcpyPeerAsync over NVLink on the offload / load streams
                buf = torch.empty(gpu_tensor.numel(), dtype=gpu_tensor.dtype, device=self.backing_device)
            self.cpu_buf[key] = buf
        out = _stream("offload")
        out.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(out):
            buf[: gpu_tensor.numel()].view(gpu_tensor.shape).copy_(gpu_tensor, non_blocking=True)
            gpu_tensor.record_stream(out)

    def offload_cur_value(self) -> None:
        self.offload(self.get_loaded_value())

    # ---- fetch
    def get_loaded_value(self) -> Optional[torch.Tensor]:
        if not self.is_offload_enabled:
            return self.gpu_tensor[self.get_cur_model_invocation_key()]
        t = _gpu_slots[self.name][self.layer_key]
        assert t is not None, (f"Tensor {self.name} is not loaded yet for layer {self.layer_num}. "
                               "Call load_async() then load_async_wait() first")
        return t

    @torch.compiler.disable
    def load_async(self) -> Optional[torch.Tensor]:
        key = self.get_cur_model_invocation_key()
        if not self.is_offload_enabled:
            return self.gpu_tensor[key]
        shape = self.real_shape[key]
        if shape is None:
            return None
        ring = _gpu_slots[self.name]
        if ring[self.layer_key] is None or ring[self.layer_key].shape != shape:
            ring[self.layer_key] = torch.empty(shape, dtype=self.cpu_buf[key].dtype, device=self.device)
        dst = ring[self.layer_key]
        inp = _stream("load")
        inp.wait_stream(torch.cuda.current_stream())
        inp.wait_stream(_stream("offload"))      # the previous offload() may still be writing this pinned buffer (D2H)
        with torch.cuda.stream(inp):
            dst.copy_(self.cpu_buf[key][: dst.numel()].view(shape), non_blocking=True)
        dst.record_stream(inp)                    # allocated on the current stream, written on the load stream
        return dst

    def load_async_wait(self) -> None:
        if not self.is_offload_enabled:
            return
        cur = torch.cuda.current_stream()
        cur.wait_stream(_stream("load"))
        cur.wait_stream(_stream("offload"))
