"""3-D voxel token reordering and static local masks for video DiTs
(reference: src/chipmunk/ops/voxel.py:9-304).

`voxel_chunk_no_padding` groups tokens of a [t, h, w] latent into vt x vh x vw voxels (so that a
192-token query group is spatially compact) and appends the ragged tails in raster order; like the
2-D patching it is a fixed permutation, built once per shape and applied with one gather (the library's
cm_gather_rows kernel for CUDA tensors: 256-byte token rows, coalesced both ways; torch.index_select for CPU tensors).
"""
from __future__ import annotations

from functools import lru_cache

import torch


@lru_cache(maxsize=16)
def _voxel_perm(t: int, h: int, w: int, vt: int, vh: int, vw: int, device: str):
    dev = torch.device(device)
    T, H, W = (t // vt) * vt, (h // vh) * vh, (w // vw) * vw
    tt = torch.arange(t, device=dev)[:, None, None].expand(t, h, w)
    hh = torch.arange(h, device=dev)[None, :, None].expand(t, h, w)
    ww = torch.arange(w, device=dev)[None, None, :].expand(t, h, w)
    src = (tt * h + hh) * w + ww
    main = (tt < T) & (hh < H) & (ww < W)
    nh, nw = H // vh if vh else 0, W // vw if vw else 0
    key_main = ((((tt // vt) * nh + hh // vh) * nw + ww // vw) * vt + tt % vt) * (vh * vw) + (hh % vh) * vw + ww % vw
    order = [src[main][key_main[main].argsort()]]
    # tails, each in raster order: t >= T (all h, w); then t < T, h >= H (all w); then t < T, h < H, w >= W
    order.append(src[tt >= T])
    order.append(src[(tt < T) & (hh >= H)])
    order.append(src[(tt < T) & (hh < H) & (ww >= W)])
    perm = torch.cat([o.reshape(-1) for o in order])
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), device=dev)
    if dev.type == "cuda":
        return perm.to(torch.int32), inv.to(torch.int32)
    return perm, inv


def _gather(x: torch.Tensor, dim: int, perm: torch.Tensor) -> torch.Tensor:
    if x.is_cuda:
        from .. import torch_ops as _t
        return _t.gather_rows(x, dim, perm)
    return x.index_select(dim, perm)


def voxel_chunk_no_padding(x: torch.Tensor, voxel_shape=(4, 4, 4)) -> torch.Tensor:
    """x [b, ah, t, h, w, d] -> [b, ah, t*h*w, d] in voxel order (tails appended)."""
    b, ah, t, h, w, d = x.shape
    perm, _ = _voxel_perm(t, h, w, *voxel_shape, str(x.device))
    return _gather(x.reshape(b, ah, t * h * w, d), 2, perm)


def reverse_voxel_chunk_no_padding(x_chunk_flat: torch.Tensor, original_shape, voxel_shape=(4, 4, 4)) -> torch.Tensor:
    b, ah, t, h, w, d = original_shape
    _, inv = _voxel_perm(t, h, w, *voxel_shape, str(x_chunk_flat.device))
    return _gather(x_chunk_flat, 2, inv).reshape(b, ah, t, h, w, d)


def masktoinds(mask: torch.Tensor, multiple=None):
    """Pure-torch statement of mask -> (indices, counts): set columns first (reference voxel.py:161-180)."""
    counts = mask.sum(dim=-1).to(torch.int32)
    if multiple is not None:
        counts = ((counts + multiple - 1) // multiple) * multiple
    inds = mask.to(torch.int8).argsort(dim=-1, descending=True, stable=True)
    return inds.contiguous().to(torch.int32), counts.contiguous()


def offsets(base_coord: int, full_size: int, offset_range: int):
    """The 2*offset_range+1 relative positions a coordinate looks at along one axis: offset_range to either side, and
    what does not fit on one side is made up for on the other (reference voxel.py:101-113, which only ever extends ONE
    side: the left one is checked first)."""
    left = min(offset_range, base_coord)
    right = min(offset_range, full_size - 1 - base_coord)
    if left < offset_range:
        right += offset_range - left
    elif right < offset_range:
        left += offset_range - right
    return list(range(-left, right + 1))


def get_local_voxel_indices(full_shape, local_shape) -> torch.Tensor:
    """[t*h*w, (lt+1)*(lh+1)*(lw+1)] int64 (CPU): for every voxel the flat indices of its local box, entry
    (ic, jc, kc) of the box at column ic*(lh+1)*(lw+1) + jc*(lw+1) + kc; the box has 2*(l//2)+1 entries per axis, so for
    odd extents the remaining columns keep their initial 0 (reference voxel.py:115-158, a six-deep Python loop; here
    three per-axis tables combined by broadcasting)."""
    t, h, w = full_shape
    lt, lh, lw = local_shape
    width = (lt + 1) * (lh + 1) * (lw + 1)
    inds = torch.zeros(t * h * w, width, dtype=torch.int64)
    if lt == 0 or lh == 0 or lw == 0:
        return inds

    def axis_table(n, l):
        # [n, 2*(l//2)+1] absolute coordinates + which entries exist (a coordinate's list is shorter, and may leave the grid,
        # where the axis is too short for the window: reproduced as the reference produces it)
        width_ = 2 * (l // 2) + 1
        rows = [[b + o for o in offsets(b, n, l // 2)] for b in range(n)]
        tab = torch.tensor([r + [0] * (width_ - len(r)) for r in rows], dtype=torch.int64)
        ok = torch.tensor([[True] * len(r) + [False] * (width_ - len(r)) for r in rows])
        return tab, ok

    (at, vt_), (ah, vh_), (aw, vw_) = axis_table(t, lt), axis_table(h, lh), axis_table(w, lw)
    nt, nh, nw = at.shape[1], ah.shape[1], aw.shape[1]

    def outer(x, y, z, op):
        return op(op(x[:, None, None, :, None, None], y[None, :, None, None, :, None]), z[None, None, :, None, None, :])

    flat = outer(at * (h * w), ah * w, aw, torch.add).reshape(t * h * w, -1)
    valid = outer(vt_, vh_, vw_, torch.logical_and).reshape(t * h * w, -1)
    slot = (torch.arange(nt)[:, None, None] * ((lh + 1) * (lw + 1)) + torch.arange(nh)[None, :, None] * (lw + 1)
            + torch.arange(nw)[None, None, :]).reshape(-1)
    inds[:, slot] = torch.where(valid, flat, torch.zeros_like(flat))
    return inds


def merge_indices(a: torch.Tensor, b: torch.Tensor, full_shape):
    """Union of two index lists per row as (inds, counts) of the merged mask (reference voxel.py:182-204).
    `full_shape` is the mask's shape [..., m, n] (a shape, or a tensor of that shape)."""
    shape = tuple(full_shape.shape) if isinstance(full_shape, torch.Tensor) else tuple(full_shape)
    assert a.shape[:-1] == b.shape[:-1] == shape[:-1], "a, b and full_shape must agree in every dimension but the last"
    mask = torch.zeros(shape, device=a.device, dtype=torch.bool)
    mask.scatter_(-1, a.long(), True)
    mask.scatter_(-1, b.long(), True)
    return masktoinds(mask)


def _axis_window(n: int, reach: int, device) -> torch.Tensor:
    """[n, n] bool: window[i, j] = j is one of the 2*reach+1 positions around i, shifted inward at the
    borders so that every position sees exactly 2*reach+1 neighbours (reference `offsets`, voxel.py:101-113)."""
    i = torch.arange(n, device=device)
    lo = (i - reach).clamp(min=0)
    lo = torch.minimum(lo, torch.tensor(max(n - (2 * reach + 1), 0), device=device))
    j = torch.arange(n, device=device)
    return (j[None, :] >= lo[:, None]) & (j[None, :] < (lo + 2 * reach + 1)[:, None])


def get_local_voxel_mask(full_shape, local_shape, device) -> torch.Tensor:
    """[t*h*w, t*h*w] bool: voxel b attends voxel a iff a lies in the (lt+1 x lh+1 x lw+1) box around b
    (the index list the reference builds with a 6-deep Python loop, voxel.py:115-158)."""
    t, h, w = full_shape
    lt, lh, lw = local_shape
    n = t * h * w
    if lt == 0 or lh == 0 or lw == 0:
        m = torch.zeros(n, n, dtype=torch.bool, device=device)
        m[:, 0] = True          # the reference returns all-zero index lists, i.e. every voxel "sees" voxel 0
        return m
    wt, wh, ww = _axis_window(t, lt // 2, device), _axis_window(h, lh // 2, device), _axis_window(w, lw // 2, device)
    m = (wt[:, None, None, :, None, None] & wh[None, :, None, None, :, None] & ww[None, None, :, None, None, :]).reshape(n, n)
    if (lt % 2) or (lh % 2) or (lw % 2):
        m = m.clone()
        m[:, 0] = True      # odd extents leave zero-initialised slots in the reference's index table -> voxel 0
    return m


def get_local_indices_with_text(vid_shape, txt_len, voxel_shape, local_shape, full_tail_from_attn=False,
                                full_tail_to_attn=False, rk=0, kv_tile_size=128, device=torch.device("cuda")):
    """Static attention mask for a video + text sequence in voxel order: every query group (one voxel of
    vt*vh*vw tokens) attends the text tokens and its local box of voxels (reference voxel.py:206-304).
    Returns (mask [n_groups, seq], inds, counts)."""
    tt, th, tw = vid_shape
    lt, lh, lw = local_shape
    vt, vh, vw = voxel_shape
    vid, seq = tt * th * tw, tt * th * tw + txt_len
    vsize = vt * vh * vw
    n_groups = (seq + vsize - 1) // vsize
    mask = torch.zeros(n_groups, seq, dtype=torch.bool, device=device)
    mask[:, vid:] = True
    gt, gh, gw = tt // vt, th // vh, tw // vw
    n_img = gt * gh * gw
    local = get_local_voxel_mask((gt, gh, gw), (lt, lh, lw), device)               # [n_img, n_img] voxels
    local = local[:, :, None].expand(n_img, n_img, vsize).reshape(n_img, n_img * vsize)[:n_groups, :seq]
    rows, cols = local.shape
    mask[:rows, :cols] |= local
    pad0, pad1 = n_groups - n_img, seq - cols
    if pad1 > 0 and full_tail_to_attn:
        mask[:, cols:] = True        # every query group, the ragged-tail groups included (reference voxel.py:262-279)
    local_size = vsize * lt * lh * lw
    if local_size > 0 and pad0 > 0:
        mask[n_groups - pad0:, seq - local_size:] = True
    tail_cols = (seq // kv_tile_size) * kv_tile_size
    txt_rows = txt_len // vsize + 1
    mask[n_groups - txt_rows:, seq - tail_cols:] = True
    if full_tail_from_attn and pad0 > 0:
        mask[n_groups - pad0:, seq - tail_cols:] = True
    if rk > 0:
        rand = torch.rand(mask.shape, device=device) < rk
        if full_tail_from_attn and pad0 > 0:
            rand[n_groups - pad0:] = False
        rand[n_groups - txt_rows:] = False
        mask |= rand
    inds, counts = masktoinds(mask, multiple=kv_tile_size)
    return mask, inds, counts
