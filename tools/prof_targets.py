"""One launch of each kernel under study, for `ncu --set full` (tools/r2_profiles.sh).

    python tools/prof_targets.py select|m2i|dense|dense_cs|attn_c2|mlp_c2 ...
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chipmunk_b200 as cm  # noqa: E402
from chipmunk_b200 import torch_ops as T  # noqa: E402

BF = torch.bfloat16
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)


def run(what):
    if what in ("select", "m2i"):
        H, G, N, k = 4, 621, 119056, 8320
        # column sums cluster in a few binades, as exp(s) * p does
        cs = torch.exp(1.5 * torch.randn(1, H, G, N, device=dev, generator=g)).to(BF)
        packed, shp, inds, cnts = T.select_columns(cs, k, 128, 0.01, None, None, 7, 192)
        if what == "m2i":
            T.bitmask_to_indices(packed, shp, 128, 192)
    elif what in ("dense", "dense_cs"):
        H, N = 24, 16384
        q, k, v = (torch.randn(1, H, N, 128, device=dev, generator=g).to(BF) for _ in range(3))
        p = torch.rand(1, H, N, 1, device=dev, generator=g) * 1e-3 + 1e-5
        T._launch_dense(q, k, v, p if what == "dense_cs" else None)
    elif what in ("dense_c3", "dense_cs_c3"):
        H, N = 24, 119056
        q, k, v = (torch.randn(1, H, N, 128, device=dev, generator=g).to(BF) for _ in range(3))
        p = torch.rand(1, H, N, 1, device=dev, generator=g) * 1e-3 + 1e-5
        T._launch_dense(q, k, v, p if what == "dense_cs_c3" else None)
    elif what == "c3attn":
        from bench import make_packed_mask, C3_N, C3_COUNT
        n, G = C3_N, (C3_N + 191) // 192
        q, k, v, c = (torch.randn(1, 24, n, 128, device=dev, generator=g).to(BF) for _ in range(4))
        packed, shp = make_packed_mask(24, G, n, C3_COUNT, g, dev)
        o = torch.empty_like(c)
        for _ in range(2):
            inds, cnts = T.bitmask_to_indices(packed, shp, 128, 192)
            T.csp_attn_add(q, k, v, c, inds, cnts, 1, out=o)
    elif what == "c2":
        from bench import make_indices, NSEQ, ATTN_COUNT, MLP_K, MLP_F, MLP_COUNT
        G = (NSEQ + 191) // 192
        q, k, v, c = (torch.randn(1, 24, NSEQ, 128, device=dev, generator=g).to(BF) for _ in range(4))
        o = torch.empty_like(c)
        idx = make_indices(24 * G, NSEQ, ATTN_COUNT, g, dev).view(1, 24, G, NSEQ)
        cnt = torch.full((1, 24, G), ATTN_COUNT, dtype=torch.int32, device=dev)
        x = torch.randn(NSEQ, MLP_K, device=dev, generator=g).to(BF)
        w1 = (0.02 * torch.randn(MLP_F, MLP_K, device=dev, generator=g)).to(BF)
        b1 = (0.02 * torch.randn(MLP_F, device=dev, generator=g)).to(BF)
        w2t = (0.02 * torch.randn(MLP_F, MLP_K, device=dev, generator=g)).to(BF)
        pa = torch.randn(MLP_F, NSEQ, device=dev, generator=g).to(BF)
        oc = torch.randn(NSEQ, MLP_K, device=dev, generator=g).to(BF)
        mi = torch.stack([torch.randperm(MLP_F, device=dev, generator=g) for _ in range(NSEQ // 128)]).int()
        mc = torch.full((NSEQ // 128,), MLP_COUNT, dtype=torch.int32, device=dev)
        packed = torch.empty(NSEQ, MLP_F, device=dev, dtype=BF)
        for _ in range(3):
            T.csp_attn_add(q, k, v, c, idx, cnt, 1, out=o)
            T.mlp_mm1(x, w1, packed, b1, pa, mi, mc, True)
            T.mlp_mm2(packed, w2t, oc, None, mi, mc, False)
    torch.cuda.synchronize()


if __name__ == "__main__":
    for w in sys.argv[1:]:
        run(w)
