"""Generate tests/golden/*.npz from the REFERENCE's own Python (run in the authoring container,
where /root/reference exists; the fixtures are committed because the reference cannot travel
to the GPU box).

    python tests/golden/make_golden.py

Sources imported by path (importing the `chipmunk` package itself fails without chipmunk.cuda):
  * /root/reference/src/chipmunk/ops/bitpack.py   -> bitpack / bitunpack   (torch.compile is
    replaced by the identity so the plain eager function runs on CPU)
  * /root/reference/src/chipmunk/ops/voxel.py     -> masktoinds (the reference's pure-torch
    statement of which index SET and which padded count mask_to_indices must produce)
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src/chipmunk/ops"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(REF, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not mounted; fixtures are already committed")
    real_compile = torch.compile
    torch.compile = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    try:
        ref_bitpack = _load("bitpack")
        ref_voxel = _load("voxel")
    finally:
        torch.compile = real_compile

    g = torch.Generator().manual_seed(20260117)

    # ---- bitpack / bitunpack: shapes incl. a non-multiple-of-8 element count
    cases = {}
    for i, shape in enumerate([(1, 2, 3, 37), (1, 3, 5, 192), (2, 2, 4, 500), (1, 1, 1, 7)]):
        mask = torch.rand(shape, generator=g) < 0.3
        packed, oshape = ref_bitpack.bitpack(mask)
        back = ref_bitpack.bitunpack(packed, oshape)
        assert torch.equal(back, mask)
        cases[f"mask{i}"] = mask.numpy()
        cases[f"packed{i}"] = packed.numpy()
    np.savez_compressed(os.path.join(HERE, "bitpack.npz"), **cases)

    # ---- masktoinds: index sets + padded counts for several densities and multiples
    cases = {}
    for i, (shape, dens, mult) in enumerate([
        ((1, 2, 3, 384), 0.2, 128),
        ((1, 2, 2, 500), 0.5, 112),
        ((2, 1, 2, 1000), 0.07, 128),
        ((1, 1, 2, 256), 0.0, 128),     # empty rows
        ((1, 1, 2, 256), 1.0, 128),     # full rows
    ]):
        mask = torch.rand(shape, generator=g) < dens
        inds, counts = ref_voxel.masktoinds(mask, multiple=mult)
        nnz = mask.sum(dim=-1).to(torch.int32)
        cases[f"mask{i}"] = mask.numpy()
        cases[f"mult{i}"] = np.int32(mult)
        cases[f"inds{i}"] = inds.numpy()          # first nnz entries of each row = the set columns
        cases[f"counts{i}"] = counts.numpy()      # nnz rounded up to `mult`
        cases[f"nnz{i}"] = nnz.numpy()
    np.savez_compressed(os.path.join(HERE, "masktoinds.npz"), **cases)
    # ---- token reordering (ops/patch.py, ops/voxel.py): toy shapes incl. ragged tails
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))      # `chipmunk` alias -> chipmunk.util for patch.py
    ref_patch = _load("patch")
    cases = {}
    x = torch.randn(3, 16, 24, generator=g)
    cases["patch_x"] = x.numpy()
    cases["patch_y"] = ref_patch.patchify(x).numpy()
    assert torch.equal(ref_patch.unpatchify(ref_patch.patchify(x), (3, 16, 24)), x)
    pe = torch.randn(1, 1, 5 + 16 * 24, 4, 2, 2, generator=g)
    cases["rope_in"] = pe.numpy().copy()
    cases["rope_out"] = ref_patch.patchify_rope((1, 16 * 24, 8), pe.clone(), 24, 16).numpy()
    for i, (shape, vox) in enumerate([((1, 2, 8, 12, 16, 3), (4, 6, 8)), ((1, 1, 9, 13, 17, 2), (4, 6, 8)), ((2, 1, 5, 4, 6, 1), (4, 4, 4))]):
        xv = torch.randn(shape, generator=g)
        yv = ref_voxel.voxel_chunk_no_padding(xv, voxel_shape=vox)
        assert torch.equal(ref_voxel.reverse_voxel_chunk_no_padding(yv, shape, voxel_shape=vox), xv)
        cases[f"vox_x{i}"] = xv.numpy(); cases[f"vox_y{i}"] = yv.numpy(); cases[f"vox_shape{i}"] = np.array(vox)
    for i, (vid, txt, local) in enumerate([((12, 18, 24), 40, (2, 2, 2)), ((16, 24, 32), 200, (0, 0, 0)), ((13, 19, 25), 70, (2, 2, 2)), ((16, 24, 32), 256, (3, 3, 3))]):
        mask, inds, counts = ref_voxel.get_local_indices_with_text(vid, txt, (4, 6, 8), local, rk=0, kv_tile_size=128,
                                                                   device=torch.device("cpu"))
        cases[f"lm_args{i}"] = np.array(list(vid) + [txt] + list(local))
        cases[f"lm_mask{i}"] = mask.numpy(); cases[f"lm_counts{i}"] = counts.numpy()
    np.savez_compressed(os.path.join(HERE, "reorder.npz"), **cases)
    # ---- the helpers under the static local masks: per-axis offsets and the local-box index table (ops/voxel.py:101-158)
    cases = {}
    grid = [(b, n, r) for n in (1, 2, 3, 5, 8) for r in (0, 1, 2, 3) for b in range(n) if not (n == 1 and r > 0)]
    offs = []
    for b, n, r in grid:
        try:
            offs.append((b, n, r, ref_voxel.offsets(b, n, r)))
        except IndexError:            # the reference indexes an empty list when neither side has room (n == 1 handled above)
            pass
    cases["offsets_args"] = np.array([o[:3] for o in offs], dtype=np.int64)
    cases["offsets_flat"] = np.array([v for o in offs for v in o[3]], dtype=np.int64)
    cases["offsets_len"] = np.array([len(o[3]) for o in offs], dtype=np.int64)
    for i, (full, local) in enumerate([((3, 4, 5), (2, 2, 2)), ((4, 4, 4), (3, 3, 3)), ((2, 3, 6), (2, 0, 2)), ((5, 3, 4), (4, 2, 2)), ((6, 6, 6), (1, 1, 1))]):
        cases[f"lvi_args{i}"] = np.array(list(full) + list(local))
        cases[f"lvi{i}"] = ref_voxel.get_local_voxel_indices(full, local).numpy()
    np.savez_compressed(os.path.join(HERE, "voxel_indices.npz"), **cases)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
