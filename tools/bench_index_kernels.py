"""HBM throughput of the index / pack kernels (north_star iii; SURVEY §8d "index kernels: pure HBM") at the
FLUX (C2) and HunyuanVideo-720p (C3) sizes: algorithmic bytes (mask / activation read + indices / bits written)
per launch / CUDA-event time, as a fraction of the measured HBM peak.

    python tools/bench_index_kernels.py [--out profiles/r02_index_kernels.json] [--quick]

Under ncu (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`) the same script gives
the per-launch DRAM traffic (tools/r2_profiles.sh).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chipmunk_b200 as cm  # noqa: E402
from chipmunk_b200 import torch_ops as T  # noqa: E402


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def timeit(fn, iters=5, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                        # > L2: the next launch starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true", help="one iteration per kernel (for ncu)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak, src = hbm_peak()
    iters = 1 if a.quick else 5
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    res = []

    def rec(name, shape, nbytes, ms):
        gbs = nbytes / ms / 1e6
        res.append({"kernel": name, "shape": shape, "algorithmic_MB": round(nbytes / 1e6, 1), "ms": round(ms, 4),
                    "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3)})
        print(json.dumps(res[-1]), flush=True)

    for tag, (H, N, topk, mult) in {"C2_flux": (24, 4608, 784, 112), "C3_hunyuan": (24, 119056, 8320, 128)}.items():
        G = (N + 191) // 192
        rows = H * G
        mask = torch.zeros(1, H, G, N, dtype=torch.bool, device=dev)
        for r0 in range(0, G, 32):
            r1 = min(G, r0 + 32)
            sel = torch.rand(H, r1 - r0, N, device=dev, generator=g).topk(topk, dim=-1).indices
            mask[0, :, r0:r1].scatter_(-1, sel, True)
        packed, shp = T.bitpack(mask)
        ms = timeit(lambda: T.bitpack(mask), iters, flush)
        rec("bitpack", f"{tag} mask [1,{H},{G},{N}]", mask.numel() + mask.numel() // 8, ms)
        ms = timeit(lambda: T.bitunpack(packed, shp), iters, flush)
        rec("bitunpack", f"{tag}", mask.numel() + mask.numel() // 8, ms)
        idx_bytes = rows * topk * 4 + rows * 4
        ms = timeit(lambda: T.mask_to_indices(mask, mult, 192), iters, flush)
        rec("mask_to_indices", f"{tag} count {topk}", mask.numel() + idx_bytes, ms)
        ms = timeit(lambda: T.bitmask_to_indices(packed, shp, mult, 192), iters, flush)
        rec("bitmask_to_indices (fused bitunpack)", f"{tag} count {topk}", mask.numel() // 8 + idx_bytes, ms)
        if hasattr(T, "select_columns"):
            cs = torch.rand(1, H, G, N, device=dev, generator=g).to(torch.bfloat16)
            ms = timeit(lambda: T.select_columns(cs, topk, mult, 0.01, None, None, 1234), iters, flush)
            rec("select_columns (top-k + random -> bit mask + indices)", f"{tag} k {topk}",
                cs.numel() * 2 + mask.numel() // 8 + idx_bytes, ms)
        del mask, packed

    # MLP-side kernels at FLUX and at the 720p token count
    for tag, M in {"C2_flux": 4608, "hunyuan_tokens": 119168}.items():
        F = 12288
        mb = M // 128
        act = torch.rand(1, mb, F, device=dev, generator=g).to(torch.bfloat16)
        inds = torch.empty(1, mb, F, dtype=torch.int32, device=dev)
        cnts = torch.empty(1, mb, dtype=torch.int32, device=dev)
        ms = timeit(lambda: T.topk_indices(act, inds, cnts, 0.7, 256, 0.0), iters, flush)
        kept = int(cnts.sum())
        rec("topk_indices", f"{tag} [1,{mb},{F}] sparsity 0.7", act.numel() * 2 + kept * 4, ms)
        ms = timeit(lambda: T.topk_indices(act, inds, cnts, 0.7, 256, 0.05), iters, flush)
        rec("topk_indices (random_amount 0.05)", f"{tag}", act.numel() * 2 + int(cnts.sum()) * 4, ms)
        src_t = torch.randn(1, mb, F, device=dev, generator=g).to(torch.bfloat16)
        dst_t = torch.zeros_like(src_t)
        ms = timeit(lambda: T.copy_indices(src_t, dst_t, inds, cnts), iters, flush)
        rec("copy_indices", f"{tag}", int(cnts.sum()) * (4 + 2 + 2), ms)
        if M <= 16384:
            packed_t = torch.randn(1, M, F, device=dev, generator=g).to(torch.bfloat16)
            pa = torch.zeros(1, F, M, dtype=torch.bfloat16, device=dev)
            ms = timeit(lambda: T.csp_scatter_add(packed_t, pa, inds, cnts, 6), iters, flush)
            rec("csp_scatter_add", f"{tag}", int(cnts.sum()) * (128 * 2 * 3 + 4), ms)

    out = {"hbm_peak_gbs": peak, "peak_source": src, "l2": "flushed (256 MB memset) before every timed launch", "kernels": res}
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
