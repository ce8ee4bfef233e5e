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
    torch.cuda.synchronize()


if __name__ == "__main__":
    for w in sys.argv[1:]:
        run(w)
