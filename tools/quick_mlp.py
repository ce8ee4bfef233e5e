"""Quick device timing of the column-sparse MLP (development aid, not the bench)."""
import sys
import torch
sys.path.insert(0, ".")
import chipmunk_b200 as cm  # noqa
import os
SORTED = os.environ.get("SORTED", "0") == "1"

def run(M=4096, K=3072, F=12288, N=3072, count=3840, iters=10):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    bf = torch.bfloat16
    x = torch.randn(M, K, device=dev, generator=g).to(bf)
    w1 = (torch.randn(F, K, device=dev, generator=g) * 0.02).to(bf)
    b1 = (torch.randn(F, device=dev, generator=g) * 0.02).to(bf)
    w2 = (torch.randn(N, F, device=dev, generator=g) * 0.02).to(bf)
    w2t = w2.t().contiguous()
    pa = torch.randn(F, M, device=dev, generator=g).to(bf)
    out = torch.randn(M, N, device=dev, generator=g).to(bf)
    idx = torch.stack([torch.randperm(F, device=dev, generator=g) for _ in range(M // 128)]).int()
    if SORTED:
        idx[:, :count] = idx[:, :count].sort(dim=-1).values
    cnt = torch.full((M // 128,), count, dtype=torch.int32, device=dev)
    packed = torch.empty(M, F, device=dev, dtype=bf)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    def t(fn, n=iters):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t1 = t(lambda: cm.torch_ops.mlp_mm1(x, w1, packed, b1, pa, idx, cnt, True))
    t2 = t(lambda: cm.torch_ops.mlp_mm2(packed, w2t, out, None, idx, cnt, False))
    te = t(lambda: cm.ops.mlp(x, w1, b1, w2t, idx, cnt, pa, out, 6))
    td = t(lambda: torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(x, w1, b1), approximate="tanh"), w2))
    f1, f2 = 2.0 * M * count * K, 2.0 * M * count * N
    dense = 4.0 * M * K * F
    print(f"M={M} count={count}: mm1 {t1*1e3:.1f} us ({f1/t1/1e9:.0f} TF/s)  mm2 {t2*1e3:.1f} us ({f2/t2/1e9:.0f} TF/s)  "
          f"e2e {te*1e3:.1f} us  dense cuBLAS {td*1e3:.1f} us ({dense/td/1e9:.0f} TF/s)  speedup {td/te:.2f}x", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(M=int(sys.argv[1]), F=int(sys.argv[2]), count=int(sys.argv[3])); sys.exit(0)
    run()
    run(count=6144)
    run(M=16384, count=3840, iters=5)
