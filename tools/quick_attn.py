"""Quick device timing of the column-sparse attention kernel (development aid, not the bench)."""
import sys, time
import torch
sys.path.insert(0, ".")
import chipmunk_b200  # noqa

def run(B, H, N, count, iters=10):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v = (torch.randn(B, H, N, 128, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
    G = (N + 191) // 192
    # random count-subset per group: topk of random keys
    idx = torch.empty(B, H, G, N, dtype=torch.int32, device=dev)
    for h in range(H):
        r = torch.rand(B, G, N, device=dev, generator=g)
        idx[:, h] = r.topk(count, dim=-1).indices.sort(dim=-1).values.int().new_zeros(B, G, N) if False else \
            torch.cat([r.topk(count, dim=-1).indices.sort(dim=-1).values.int(),
                       torch.zeros(B, G, N - count, dtype=torch.int32, device=dev)], dim=-1)
    cnt = torch.full((B, H, G), count, dtype=torch.int32, device=dev)
    o = torch.zeros_like(q)
    for _ in range(3):
        torch.ops.chipmunk.csp_attn(q, k, v, o, idx, cnt, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        torch.ops.chipmunk.csp_attn(q, k, v, o, idx, cnt, 1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tiles = B * H * G
    flops = 4.0 * 192 * count * 128 * tiles
    dense = 4.0 * N * N * 128 * B * H
    gbytes = tiles * (count * 516 + 3 * 192 * 128 * 2) / 1e9
    print(f"N={N} H={H} B={B} count={count}: {ms*1e3:.1f} us  sparse {flops/ms/1e9:.0f} TFLOP/s  "
          f"dense-equiv {dense/ms/1e9:.0f} TFLOP/s  gather {gbytes/ms*1e3:.0f} GB/s", flush=True)
    # dense baseline (torch SDPA)
    if N <= 130000:
        for _ in range(2): torch.nn.functional.scaled_dot_product_attention(q, k, v)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): torch.nn.functional.scaled_dot_product_attention(q, k, v)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 5
        print(f"    torch SDPA dense: {ms2*1e3:.1f} us ({dense/ms2/1e9:.0f} TFLOP/s) -> speedup {ms2/ms:.2f}x", flush=True)

if __name__ == "__main__":
    if len(sys.argv) >= 3:
        N, count = int(sys.argv[1]), int(sys.argv[2])
        B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
        H = int(sys.argv[4]) if len(sys.argv) > 4 else 24
        it = int(sys.argv[5]) if len(sys.argv) > 5 else 10
        run(B, H, N, count, iters=it)
    else:
        run(2, 24, 4096, 768)
        run(1, 24, 4608, 672)
        run(1, 24, 16384, 2944)
