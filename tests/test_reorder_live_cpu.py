"""Token reorderings and static masks against the LIVE reference code over many shapes (authoring container only: the
reference's src/chipmunk/ops/{patch,voxel}.py are imported by path from /root/reference, as tests/golden/make_golden.py does;
elsewhere the committed fixtures of tests/test_oracle.py stand in).  The reference builds these transforms from chains of
einops rearranges and Python loops, this repo from one cached permutation / vectorised windows: they must agree exactly,
ragged tails, odd local extents and every flag of `get_local_indices_with_text` included."""
import importlib.util
import itertools
import os
import sys

import pytest
import torch

REF = "/root/reference/src/chipmunk/ops"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the authoring container")


@pytest.fixture(scope="module")
def ref(cm):
    from chipmunk_b200.util.config import reset_to_defaults
    reset_to_defaults()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)          # `from chipmunk.util import GLOBAL_CONFIG` in the reference's patch.py -> the alias package
    mods = {}
    for name in ("patch", "voxel"):
        spec = importlib.util.spec_from_file_location(f"_ref_live_{name}", os.path.join(REF, f"{name}.py"))
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    return mods


@pytest.mark.parametrize("h,w", [(8, 8), (16, 24), (24, 16), (64, 64), (40, 72), (8, 128)])
def test_patchify_matches_the_reference(cm, ref, h, w):
    from chipmunk_b200.ops import patch
    g = torch.Generator().manual_seed(h * 1000 + w)
    x = torch.randn(3, h, w, generator=g)
    want = ref["patch"].patchify(x)
    got = patch.patchify(x)
    assert torch.equal(got, want)
    assert torch.equal(patch.unpatchify(got, x.shape), ref["patch"].unpatchify(want, x.shape))
    assert torch.equal(patch.unpatchify(got, x.shape), x)
    pe = torch.randn(1, 2, 7 + h * w, 4, 2, 2, generator=g)
    assert torch.equal(patch.patchify_rope((1, h * w, 8), pe.clone(), w, h), ref["patch"].patchify_rope((1, h * w, 8), pe.clone(), w, h))


@pytest.mark.parametrize("thw,vox", [((8, 12, 16), (4, 6, 8)), ((9, 13, 17), (4, 6, 8)), ((5, 4, 6), (4, 4, 4)), ((4, 6, 8), (4, 6, 8)),
                                     ((3, 5, 7), (4, 6, 8)), ((12, 7, 9), (2, 3, 4)), ((9, 34, 60), (4, 6, 8)), ((1, 6, 8), (1, 6, 8))])
def test_voxel_chunking_matches_the_reference(cm, ref, thw, vox):
    from chipmunk_b200.ops import voxel
    g = torch.Generator().manual_seed(sum(thw))
    x = torch.randn(2, 2, *thw, 3, generator=g)
    if any(a < b for a, b in zip(thw, vox)):
        # no whole voxel fits: the reference cannot express it (einops on an empty main block); ours must still round-trip
        y = voxel.voxel_chunk_no_padding(x, vox)
        assert torch.equal(voxel.reverse_voxel_chunk_no_padding(y, x.shape, vox), x)
        return
    want = ref["voxel"].voxel_chunk_no_padding(x, voxel_shape=vox)
    got = voxel.voxel_chunk_no_padding(x, vox)
    assert torch.equal(got, want)
    assert torch.equal(voxel.reverse_voxel_chunk_no_padding(got, x.shape, vox), x)
    assert torch.equal(ref["voxel"].reverse_voxel_chunk_no_padding(want, x.shape, voxel_shape=vox), x)


@pytest.mark.parametrize("vid,txt,local", [((12, 18, 24), 40, (2, 2, 2)), ((8, 12, 16), 0, (1, 1, 1)), ((13, 19, 25), 70, (2, 2, 2)),
                                           ((16, 24, 32), 256, (3, 3, 3)), ((8, 24, 32), 100, (0, 0, 0)), ((12, 12, 24), 33, (2, 1, 3)),
                                           ((9, 34, 60), 256, (1, 1, 1))])
@pytest.mark.parametrize("tail_from,tail_to", [(False, False), (True, False), (False, True), (True, True)])
def test_static_local_masks_match_the_reference(cm, ref, vid, txt, local, tail_from, tail_to):
    from chipmunk_b200.ops import voxel
    kw = dict(full_tail_from_attn=tail_from, full_tail_to_attn=tail_to, rk=0, kv_tile_size=128, device=torch.device("cpu"))
    want_mask, want_inds, want_counts = ref["voxel"].get_local_indices_with_text(vid, txt, (4, 6, 8), local, **kw)
    mask, inds, counts = voxel.get_local_indices_with_text(vid, txt, (4, 6, 8), local, **kw)
    assert torch.equal(mask, want_mask)
    assert torch.equal(counts, want_counts)
    nnz = mask.sum(-1)
    for r in range(mask.shape[0]):            # the reference's argsort is not stable: compare the index SETS
        n = int(nnz[r])
        assert torch.equal(inds[r, :n].sort().values, want_inds[r, :n].sort().values)


def test_local_voxel_tables_match_the_reference(cm, ref):
    from chipmunk_b200.ops import voxel
    for full, local in itertools.product([(2, 3, 4), (3, 3, 3), (4, 5, 2), (5, 6, 7)], [(1, 1, 1), (2, 2, 2), (3, 3, 3), (2, 1, 3), (4, 2, 2), (2, 0, 2)]):
        if any(l // 2 > 0 and f == 1 for f, l in zip(full, local)):
            continue
        try:
            want = ref["voxel"].get_local_voxel_indices(full, local)
        except IndexError:
            continue        # the reference's `offsets` indexes an empty list when an axis has no room on either side
        assert torch.equal(voxel.get_local_voxel_indices(full, local), want), (full, local)
        if int(want.max()) >= want.shape[0]:
            continue        # an axis shorter than its window: the reference's table runs off the grid (its own scatter_ would raise)
        # and the mask form used by get_local_indices_with_text scatters exactly these indices
        m = torch.zeros(want.shape[0], want.shape[0], dtype=torch.bool)
        m.scatter_(-1, want, True)
        assert torch.equal(voxel.get_local_voxel_mask(full, local, torch.device("cpu")), m), (full, local)


@pytest.mark.parametrize("seq,txt,lv,lw1d,top", [((8, 12, 16), 64, 1, 0.0, 0.05), ((8, 12, 16), 64, 0, 0.1, 0.05), ((12, 18, 24), 40, 2, 0.05, 0.1),
                                                 ((9, 13, 17), 70, 1, 0.3, 0.3), ((16, 24, 32), 256, 3, 0.0, 0.6), ((4, 6, 8), 5, 0, 0.0, 0.05),
                                                 ((16, 24, 32), 100, 1, 1.0, 0.01)])
def test_initialize_static_mask_matches_the_reference_module(cm, ref, monkeypatch, seq, txt, lv, lw1d, top):
    """`SparseDiffAttn.initialize_static_mask` (3-D local voxels + the 1-D window + which query groups stay sparse) against the
    reference's module code (src/chipmunk/modules/attn.py:23-73), which loops over the query groups in Python."""
    import numpy as np
    import chipmunk  # noqa: F401  the alias package: `chipmunk.util`, `chipmunk.ops` of the reference module resolve to this repo ...
    import chipmunk_b200.modules.attn as OurA
    from chipmunk_b200.util.config import GLOBAL_CONFIG, reset_to_defaults
    from chipmunk_b200.util.layer_counter import LayerCounter

    reset_to_defaults()
    GLOBAL_CONFIG["attn"].update(top_keys=top, random_keys=0.0, local_voxels=lv, local_1d_window=lw1d)
    monkeypatch.setitem(sys.modules, "chipmunk.ops.voxel", ref["voxel"])        # ... except the voxel masks: the reference's own
    spec = importlib.util.spec_from_file_location("_ref_live_modules_attn", "/root/reference/src/chipmunk/modules/attn.py")
    RefA = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(RefA)
    H, dev = 2, torch.device("cpu")
    RefA.SparseDiffAttn(0, LayerCounter(1, 1)).initialize_static_mask(seq, txt, H, dev)

    def pack_rows_on_cpu(mask2d):          # ops.pack_rows_to_words is a CUDA kernel; same layout with numpy for this CPU test
        R, n = mask2d.shape
        W = (n + 31) // 32
        padded = np.zeros((R, W * 32), dtype=np.uint8)
        padded[:, :n] = mask2d.numpy()
        return torch.from_numpy(np.packbits(padded, axis=1, bitorder="little").view(np.int32).reshape(R, W).copy())
    monkeypatch.setattr(OurA.ops, "pack_rows_to_words", pack_rows_on_cpu)
    for name in ("singleton_static_mask", "singleton_video_query_groups", "singleton_static_words", "singleton_group_flags"):
        monkeypatch.setattr(OurA, name, None)
    OurA.SparseDiffAttn(0, LayerCounter(1, 1)).initialize_static_mask(seq, txt, H, dev)
    assert torch.equal(OurA.singleton_static_mask, RefA.singleton_static_mask)
    assert torch.equal(OurA.singleton_video_query_groups, RefA.singleton_video_query_groups)
    # the bit-packed copy handed to select_columns is that mask, and the group flags are that column
    words = OurA.singleton_static_words.numpy().view(np.uint8)
    bits = np.unpackbits(words, axis=1, bitorder="little")[:, : RefA.singleton_static_mask.shape[-1]].astype(bool)
    assert np.array_equal(bits, RefA.singleton_static_mask[0, 0].numpy())
    assert torch.equal(OurA.singleton_group_flags, RefA.singleton_video_query_groups[0, 0, :, 0])
    reset_to_defaults()


def test_oracle_index_path_matches_the_reference_python_on_random_shapes(oracle, ref):
    """The oracle's bit codec and mask -> indices (sets + padded counts) against the reference's own torch functions
    (ops/bitpack.py run eagerly, ops/voxel.py:masktoinds) on 40 random shapes / densities / multiples -- the committed fixtures
    of tests/test_oracle.py hold five."""
    real_compile = torch.compile
    torch.compile = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    try:
        spec = importlib.util.spec_from_file_location("_ref_live_bitpack", os.path.join(REF, "bitpack.py"))
        bp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bp)
    finally:
        torch.compile = real_compile
    g = torch.Generator().manual_seed(4242)
    for i in range(40):
        b, h, m = (int(torch.randint(1, 4, (1,), generator=g)) for _ in range(3))
        n = int(torch.randint(1, 700, (1,), generator=g))
        dens = float(torch.rand(1, generator=g)) ** 2
        mult = (16, 112, 128, 192)[i % 4]
        mask = torch.rand(b, h, m, n, generator=g) < dens
        packed, shape = bp.bitpack(mask)
        opacked, oshape = oracle.bitpack(mask)
        assert torch.equal(opacked, packed) and tuple(oshape) == tuple(shape)
        assert torch.equal(oracle.bitunpack(packed, shape), mask) and torch.equal(bp.bitunpack(packed, shape), mask)
        want_inds, want_counts = ref["voxel"].masktoinds(mask, multiple=mult)
        inds, counts = oracle.mask_to_indices(mask, mult, 192)
        assert torch.equal(counts, want_counts), (i, n, mult)
        nnz = mask.sum(-1)
        flat_i, flat_w = inds.reshape(-1, inds.shape[-1]), want_inds.reshape(-1, n)
        for r, c in enumerate(nnz.reshape(-1).tolist()):
            assert torch.equal(flat_i[r, :c].sort().values, flat_w[r, :c].sort().values), (i, r)
