"""Development check of the one-pass dense / dense+colsum attention kernel (csrc/dense_attn.cu): correctness against
torch fp32 at growing sizes, then timing against cuDNN SDPA.

    python tools/quick_dense.py [--time] [--big]
"""
import argparse
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chipmunk_b200 as cm  # noqa: E402
from chipmunk_b200 import torch_ops as T  # noqa: E402

BF = torch.bfloat16
dev = torch.device("cuda", 0)


def ref(q, k, v, p):
    s = (q.float() @ k.float().transpose(-1, -2)) * 128 ** -0.5
    o = torch.softmax(s, dim=-1) @ v.float()
    l = torch.exp(-torch.logsumexp(s, dim=-1, keepdim=True))
    cs = None
    if p is not None:
        B, H, N, _ = q.shape
        G = (N + 191) // 192
        e = torch.exp(s) * p.reshape(B, H, N, 1)
        pad = G * 192 - N
        if pad:
            e = torch.cat([e, e.new_zeros(B, H, pad, e.shape[-1])], dim=2)
        cs = e.view(B, H, G, 192, -1).sum(dim=3)
    return o, l, cs


def check(B, H, Nq, Nk, strided=False, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    if strided:
        qkv = torch.randn(B, max(Nq, Nk), 3, H, 128, device=dev, generator=g).to(BF)
        q, k, v = (qkv[:, :n, i].permute(0, 2, 1, 3) for i, n in ((0, Nq), (1, Nk), (2, Nk)))
    else:
        q = torch.randn(B, H, Nq, 128, device=dev, generator=g).to(BF)
        k = torch.randn(B, H, Nk, 128, device=dev, generator=g).to(BF)
        v = torch.randn(B, H, Nk, 128, device=dev, generator=g).to(BF)
    p = torch.rand(B, H, Nq, 1, device=dev, generator=g) * 1e-2 + 1e-4
    ro, rl, rcs = ref(q, k, v, p)
    msgs = []
    for with_cs in (False, True):
        o, cs, l = T._launch_dense(q, k, v, p if with_cs else None)
        torch.cuda.synchronize()
        eo = float((o.float() - ro).norm() / ro.norm())
        el = float(((l - rl).abs() / rl).max())
        m = f"cs={int(with_cs)} o_rel={eo:.2e} l_rel={el:.2e}"
        ok = eo < 6e-3 and el < 3e-3
        if with_cs:
            ec = float(((cs.float() - rcs).abs() / rcs.abs().clamp_min(1e-30)).max())
            ef = float((cs.float() - rcs).norm() / rcs.norm())
            m += f" cs_maxrel={ec:.2e} cs_fro={ef:.2e}"
            ok = ok and ec < 2e-2 and ef < 6e-3
        msgs.append(m + ("" if ok else "  <-- FAIL"))
    print(f"B={B} H={H} Nq={Nq} Nk={Nk} strided={strided}: " + " | ".join(msgs), flush=True)


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--first", type=int, default=0, help="only the first N correctness cases (for compute-sanitizer)")
    a = ap.parse_args()
    if a.first:
        for args in [(1, 1, 128, 128), (1, 1, 128, 256), (1, 1, 128, 384)][: a.first]:
            check(*args)
        return
    for args in [(1, 1, 128, 128), (1, 1, 128, 256), (1, 1, 128, 384), (1, 1, 256, 512), (1, 2, 384, 384), (1, 2, 500, 500),
                 (2, 3, 1000, 1000), (1, 2, 777, 333), (1, 4, 2048, 2048, True), (1, 24, 4608, 4608, True)]:
        check(*args)
    if a.time:
        for H, N in [(24, 4608), (24, 16384)] + ([(24, 119056)] if a.big else []):
            g = torch.Generator(device=dev).manual_seed(1)
            q, k, v = (torch.randn(1, H, N, 128, device=dev, generator=g).to(BF) for _ in range(3))
            p = torch.rand(1, H, N, 1, device=dev, generator=g) * 1e-3 + 1e-5
            fl = 4.0 * N * N * 128 * H
            n = 5 if N < 100000 else 2
            t_new = timeit(lambda: T._launch_dense(q, k, v, None), n)
            t_cs = timeit(lambda: T._launch_dense(q, k, v, p), n)
            t_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), n)
            print(f"H={H} N={N}: one-pass dense {t_new:.3f} ms ({fl / t_new / 1e9:.0f} TF/s), dense+colsum {t_cs:.3f} ms "
                  f"({fl / t_cs / 1e9:.0f} TF/s-equiv) | "
                  f"cuDNN SDPA {t_sdpa:.3f} ms ({fl / t_sdpa / 1e9:.0f} TF/s)", flush=True)


if __name__ == "__main__":
    main()
