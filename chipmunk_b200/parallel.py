"""Single-node multi-GPU plumbing for the sparse-delta path: one process per GPU, heads (attention)
or 128-token blocks (MLP) sharded across ranks, and ONE all-gather of the attention output per layer.

This replaces the reference's HunyuanVideo head-parallel path -- two `all_to_all_single` per attention
plus an `all_gather_into_tensor` of the text tokens (examples/hunyuan/hyvideo/modules/head_parallel.py:
36-115, attenion.py:229-292) -- for the case where every rank already holds q/k/v of its heads:
index/mask tensors are per head, so they shard with the heads and never move.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `n_units` for `rank`; the first n_units % world ranks get one more."""
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_heads(num_heads: int, world: int, rank: int) -> Tuple[int, int]:
    """The reference requires heads % world == 0 (`local_heads_num = heads // world_size`, models.py:749);
    here uneven splits are allowed and the gather pads to the largest shard."""
    return shard_range(num_heads, world, rank)


def all_gather_heads(o_local: torch.Tensor, num_heads: int, group=None) -> torch.Tensor:
    """o_local [B, h_local, N, D] on every rank  ->  [B, num_heads, N, D] on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return o_local
    B, h_local, N, D = o_local.shape
    h_max = -(-num_heads // world)
    if num_heads % world == 0:
        out = o_local.new_empty(world * B, h_local, N, D)        # rank-major concatenation along dim 0
        dist.all_gather_into_tensor(out, o_local.contiguous(), group=group)
        return out.view(world, B, h_local, N, D).permute(1, 0, 2, 3, 4).reshape(B, num_heads, N, D)
    padded = o_local.new_zeros(B, h_max, N, D)
    padded[:, :h_local] = o_local
    out = o_local.new_empty(world * B, h_max, N, D)
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view(world, B, h_max, N, D)
    parts = []
    for r in range(world):
        b, e = shard_heads(num_heads, world, r)
        parts.append(out[r, :, : e - b])
    return torch.cat(parts, dim=1)


# ---------------------------------------------------------------------------------------------------
# Fused compute + all-gather over NVSwitch multicast (NVLS).
#
# The layer output lives in a SYMMETRIC buffer [world, B, h_local, N, D] (torch symmetric memory: the same
# allocation on every GPU of the NVSwitch domain, mapped into each other's address space, plus one multicast
# address that aliases all copies).  Each rank's attention kernel stores its rows ONCE, with `multimem.st`, to the
# multicast alias of its own slice `[rank]`: the switch replicates every 16-byte store into all GPUs' buffers while
# the rank's remaining tiles are still being computed, so the gather costs no extra pass over O and no separate
# collective -- only a device-side barrier when the kernels are done.  This is what replaces the reference's two
# all_to_all + all_gather per attention (hyvideo/modules/head_parallel.py:42-115) on B200.
_SYMM = {}
_FUSED_OK = {}      # (group id, shape, dtype, device) -> bool, decided COLLECTIVELY once


class NvlsUnavailable(RuntimeError):
    """Symmetric memory / NVLS multicast cannot be used for this (group, shape) on this system."""


def _symm_buffer(shape, dtype, device, group, need_multicast: bool = True):
    import torch.distributed._symmetric_memory as symm

    key = (tuple(shape), dtype, device.index, id(group))
    if key not in _SYMM:
        try:
            buf = symm.empty(*shape, dtype=dtype, device=device)
            hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
        except Exception as e:  # noqa: BLE001  (allocator / rendezvous failures are "not available", nothing else is)
            raise NvlsUnavailable(f"symmetric memory rendezvous failed: {e!r}") from e
        _SYMM[key] = (buf, hdl)
    buf, hdl = _SYMM[key]
    if need_multicast and not getattr(hdl, "multicast_ptr", 0):
        raise NvlsUnavailable("NVLS multicast is not available on this system: use fused='peers' or fused=False")
    return buf, hdl


def fused_gather_available(shape, dtype, device, group=None, mode: str = "multicast") -> bool:
    """Whether EVERY rank of `group` can use the fused gather for this output shape: mode "multicast" = NVLS multicast
    stores, mode "peers" = plain stores into every peer's copy (symmetric memory without a multicast object).  The probe
    (symmetric allocation + rendezvous [+ multicast pointer]) runs once per (group, shape, mode); the ranks then agree on
    the outcome with an all_reduce(MIN), so that no rank can enter the device barriers of the fused path while another
    one falls back to the NCCL all-gather (mismatched collectives = hang).  The decision is cached."""
    key = (id(group), tuple(shape), dtype, device.index, mode)
    if key not in _FUSED_OK:
        try:
            _symm_buffer(shape, dtype, device, group, need_multicast=(mode == "multicast"))
            ok = torch.ones(1, device=device, dtype=torch.int32)
        except NvlsUnavailable:
            ok = torch.zeros(1, device=device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        _FUSED_OK[key] = bool(int(ok.item()))
    return _FUSED_OK[key]


def sparse_attention_head_parallel_fused(q, k, v, o_cache, indices, counts, num_heads: int, group=None, mode: str = "multicast") -> torch.Tensor:
    """One sparse attention step, heads sharded over the ranks of `group` (num_heads % world == 0), with the
    all-gather of O fused into the kernel's epilogue: NVLS multicast stores (mode "multicast"), or plain stores into every
    peer's copy over NVLink (mode "peers": no multicast object needed).  Returns the full [B, H, N, D]
    output as a view of the symmetric buffer: it stays valid until the next call with the same shape."""
    from . import torch_ops as _t

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, h_local, N, D = q.shape
    if num_heads != h_local * world:
        raise RuntimeError("fused head-parallel attention needs num_heads == world_size * local heads")
    buf, hdl = _symm_buffer((world, B, h_local, N, D), q.dtype, q.device, group, need_multicast=(mode == "multicast"))
    hdl.barrier(channel=0)                       # every rank has consumed the previous contents of the buffer
    if mode == "peers":
        mine = int(hdl.buffer_ptrs[rank])
        deltas = [int(hdl.buffer_ptrs[r]) - mine for r in range(world)]
        _t.csp_attn_add(q, k, v, o_cache, indices, counts, 1, out=buf[rank], peer_deltas=deltas)
        hdl.barrier(channel=1)
        return buf.permute(1, 0, 2, 3, 4).reshape(B, num_heads, N, D)
    mc_delta = int(hdl.multicast_ptr) - int(hdl.buffer_ptrs[rank])
    _t.csp_attn_add(q, k, v, o_cache, indices, counts, 1, out=buf[rank], multicast_delta=mc_delta)
    hdl.barrier(channel=1)                       # every rank's kernel has finished: all slices are everywhere
    return buf.permute(1, 0, 2, 3, 4).reshape(B, num_heads, N, D)


def sparse_attention_head_parallel(q, k, v, o_cache, indices, counts, num_heads: int, group=None, fused=None) -> torch.Tensor:
    """One sparse attention step with heads sharded over the ranks of `group`.
    q/k/v/o_cache/indices/counts hold this rank's heads only; returns the full [B, H, N, D] output.
    fused=None: use the multicast-fused kernel when NVLS symmetric memory is available (8 GPUs, 720p layer:
    1.83 ms vs 2.80 ms), else the peer-store fused kernel when symmetric memory is available without multicast, else the
    in-place NCCL all-gather; fused=False forces the NCCL path, fused="peers" the peer-store kernel."""
    from . import torch_ops as _t

    if fused is not False and dist.is_initialized() and dist.get_world_size(group) > 1 \
            and num_heads == q.shape[1] * dist.get_world_size(group) and q.is_cuda:
        w = dist.get_world_size(group)
        B_, hl_, N_, D_ = q.shape
        # decided once per (group, shape) by ALL ranks together; kernel / argument errors of the fused call propagate
        shape = (w, B_, hl_, N_, D_)
        if fused != "peers" and fused_gather_available(shape, q.dtype, q.device, group):
            return sparse_attention_head_parallel_fused(q, k, v, o_cache, indices, counts, num_heads, group)
        if w <= 8 and fused_gather_available(shape, q.dtype, q.device, group, mode="peers"):
            return sparse_attention_head_parallel_fused(q, k, v, o_cache, indices, counts, num_heads, group, mode="peers")
        if fused:
            raise NvlsUnavailable("a fused gather was requested but symmetric memory is not available on every rank")

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    B, h_local, N, D = q.shape
    if world > 1 and num_heads % world == 0:
        # the kernel writes cache + delta straight into this rank's slice of the gather buffer (rank-major), and
        # the all-gather runs in place on it: no clone of the cache, no staging copy of O
        rank = dist.get_rank(group)
        out = q.new_empty(world, B, h_local, N, D)
        _t.csp_attn_add(q, k, v, o_cache, indices, counts, 1, out=out[rank])
        dist.all_gather_into_tensor(out.view(world * B, h_local, N, D), out[rank], group=group)
        return out.permute(1, 0, 2, 3, 4).reshape(B, num_heads, N, D)
    o = _t.csp_attn_add(q, k, v, o_cache, indices, counts, 1)
    return all_gather_heads(o, num_heads, group)


class HeadParallelAttn:
    """`SparseDiffAttn` for one rank's heads of a head-parallel layer (the reference's multi-GPU HunyuanVideo mode,
    hyvideo/modules/head_parallel.py:42-115 + attenion.py:229-292, with ONE gather of O per step instead of two
    all_to_all + all_gather).  q, k, v hold this rank's heads [B, H / world, N, D]; the call returns all heads [B, H, N, D].

    * sparse steps: the module's stored mask -> indices, then `sparse_attention_head_parallel`: the delta-attention kernel
      writes cache + delta straight into every GPU's copy of the output (NVLS multicast epilogue; NCCL all-gather fallback);
    * full steps: dense (+ column sums, selection, cache build) run on the local heads -- every tile of the path is
      independent given its head's K/V, so nothing is exchanged -- and the dense output is gathered with one
      in-place ncclAllGather (the dense kernel's TMA store writes directly into this rank's slice of the gather buffer).
    """

    def __init__(self, attn_module, num_heads: int, group=None):
        self.attn = attn_module
        self.num_heads = num_heads
        self.group = group

    def __call__(self, q, k, v):
        from .util import GLOBAL_CONFIG
        from . import ops

        attn, cfg = self.attn, GLOBAL_CONFIG["attn"]
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1 or not cfg["is_enabled"]:
            return attn(q, k, v)
        counter = attn.layer_counter
        full = counter.should_do_full_attn_step() or attn.layer_num < cfg["first_n_dense_layers"]
        if full:
            o_local = attn(q, k, v)                                   # advances the counter itself
            return all_gather_heads(o_local, self.num_heads, self.group)
        multiple_of = 128 if cfg["pad_qkv_before_kernel"] else cfg["counts_multiple_of"]
        inds, counts = attn._stored_indices(multiple_of, cfg["mbm"])
        out = sparse_attention_head_parallel(q, k, v, attn.storage.get_out_cache(), inds, counts, self.num_heads, self.group)
        counter.increment()
        return out
