"""Timing of the dense full-step kernels (dense_attn, dense_colsum_attn) vs torch SDPA."""
import sys, torch
sys.path.insert(0, ".")
import chipmunk_b200  # noqa
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for N in (4608, 16384):
    q, k, v = (torch.randn(1, 24, N, 128, device="cuda").to(torch.bfloat16) for _ in range(3))
    o, l = torch.ops.chipmunk.dense_attn(q, k, v)
    td = t(lambda: torch.ops.chipmunk.dense_attn(q, k, v))
    tc = t(lambda: torch.ops.chipmunk.dense_colsum_attn(q, k, v, l))
    ts = t(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    fl = 4.0 * N * N * 128 * 24
    print(f"N={N}: dense_attn {td*1e3:.0f} us ({fl/td/1e9:.0f} TF/s)  dense_colsum_attn {tc*1e3:.0f} us  SDPA {ts*1e3:.0f} us ({fl/ts/1e9:.0f} TF/s)", flush=True)
