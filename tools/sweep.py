"""Sparsity x sequence-length sweep (BASELINE.json configs[4]): column-sparse attention vs dense SDPA
and column-sparse MLP vs cuBLAS, as absolute times, dense-equivalent TFLOP/s and fraction of the
gather-bytes roofline.  One JSON object per line on stdout; `--out` also writes them to a file.

    python tools/sweep.py [--out profiles/r01_sweep.jsonl] [--quick]

Timing: CUDA events on the current stream, >= 3 warm-ups, median of the timed launches; inputs are far
larger than L2 at every point except the 4k attention cases (noted per line as "l2_resident").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chipmunk_b200 as cm  # noqa: E402
from chipmunk_b200 import torch_ops as T  # noqa: E402
from bench import attn_alg_bytes, make_indices, peaks  # noqa: E402

H, D, QG = 24, 128, 192


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def attention_sweep(dev, seqs, sparsities, hbm, emit):
    bf = torch.bfloat16
    for n in seqs:
        g = torch.Generator(device=dev).manual_seed(n)
        q, k, v, cache = (torch.randn(1, H, n, D, device=dev, generator=g).to(bf) for _ in range(4))
        out = torch.empty_like(q)
        G = (n + QG - 1) // QG
        iters = 20 if n <= 16384 else (6 if n <= 65536 else 3)
        t_dense = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), iters, warm=2)
        dense_flops = 4.0 * n * n * D * H
        for s in sparsities:
            count = max(128, 128 * round((1 - s) * n / 128))
            idx = make_indices(H * G, n, count, g, dev).view(1, H, G, n)
            cnt = torch.full((1, H, G), count, dtype=torch.int32, device=dev)
            t = timed(lambda: T.csp_attn_add(q, k, v, cache, idx, cnt, 1, out=out), iters)
            alg = attn_alg_bytes(H, n, count)
            emit({"op": "csp_attn_add", "seq": n, "sparsity": round(1 - count / n, 4), "count": count, "ms": round(t, 4),
                  "dense_sdpa_ms": round(t_dense, 4), "speedup_vs_dense": round(t_dense / t, 2),
                  "dense_equiv_tflops": round(dense_flops / t / 1e9, 1),
                  "sparse_tflops": round(4.0 * QG * count * D * H * G / t / 1e9, 1),
                  "gather_gbs": round(alg / t / 1e6, 1), "roofline_frac": round(alg / t / 1e6 / hbm, 4),
                  "l2_resident": n <= 8192})
            del idx, cnt
        del q, k, v, cache, out
        torch.cuda.empty_cache()


def mlp_sweep(dev, ms, sparsities, hbm, emit):
    bf = torch.bfloat16
    K, F, N = 3072, 12288, 3072
    Fn = torch.nn.functional
    for M in ms:
        g = torch.Generator(device=dev).manual_seed(M)
        x = torch.randn(M, K, device=dev, generator=g).to(bf)
        w1 = (0.02 * torch.randn(F, K, device=dev, generator=g)).to(bf)
        b1 = (0.02 * torch.randn(F, device=dev, generator=g)).to(bf)
        w2t = (0.02 * torch.randn(F, N, device=dev, generator=g)).to(bf)
        w2 = w2t.t().contiguous()
        pa = torch.randn(F, M, device=dev, generator=g).to(bf)
        oc = torch.randn(M, N, device=dev, generator=g).to(bf)
        packed = torch.empty(M, F, device=dev, dtype=bf)
        idx = torch.stack([torch.randperm(F, device=dev, generator=g) for _ in range(M // 128)]).int()
        t_dense = timed(lambda: Fn.linear(Fn.gelu(Fn.linear(x[None], w1, b1), approximate="tanh"), w2), 10)
        dense_flops = 4.0 * M * K * F
        for s in sparsities:
            count = 256 * -(-int((1 - s) * F) // 256)
            cnt = torch.full((M // 128,), count, dtype=torch.int32, device=dev)
            t1 = timed(lambda: T.mlp_mm1(x, w1, packed, b1, pa, idx, cnt, True), 10)
            t2 = timed(lambda: T.mlp_mm2(packed, w2t, oc, None, idx, cnt, False), 10)
            a1 = (M // 128) * count * K * 2 + M * K * 2 + 2 * M * count * 2
            a2 = (M // 128) * count * N * 2 + M * count * 2 + 2 * M * N * 2
            emit({"op": "csp_mlp", "M": M, "sparsity": round(1 - count / F, 4), "count": count,
                  "mm1_us": round(t1 * 1e3, 1), "mm2_us": round(t2 * 1e3, 1), "dense_cublas_us": round(t_dense * 1e3, 1),
                  "speedup_vs_dense": round(t_dense / (t1 + t2), 2),
                  "dense_equiv_tflops": round(dense_flops / (t1 + t2) / 1e9, 1),
                  "sparse_tflops": round(2.0 * M * count * (K + N) / (t1 + t2) / 1e9, 1),
                  "mm1_roofline_frac": round(a1 / t1 / 1e6 / hbm, 4), "mm2_roofline_frac": round(a2 / t2 / 1e6 / hbm, 4)})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true", help="4k/16k attention and M=4608 MLP only")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    hbm = peaks()[0]
    lines = []

    def emit(d):
        lines.append(d)
        print(json.dumps(d), flush=True)

    sp = [0.50, 0.70, 0.82, 0.93]
    attention_sweep(dev, [4096, 16384] if a.quick else [4096, 16384, 65536, 119056], sp, hbm, emit)
    mlp_sweep(dev, [4608] if a.quick else [4096, 4608, 16384, 119168], sp, hbm, emit)
    if a.out:
        with open(a.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
