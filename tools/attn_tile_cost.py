"""Fixed per-tile cost vs per-step cost of the column-sparse attention kernel at FLUX sizes: T(count) over count."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chipmunk_b200 import torch_ops as T
from bench import make_indices
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
H, N = 24, 4608
G = (N + 191) // 192
q, k, v, c = (torch.randn(1, H, N, 128, device=dev, generator=g).to(torch.bfloat16) for _ in range(4))
o = torch.empty_like(c)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for count in (128, 256, 512, 768, 784, 1024, 1536, 2304, 4608):
    idx = make_indices(H * G, N, count, g, dev).view(1, H, G, N)
    cnt = torch.full((1, H, G), count, dtype=torch.int32, device=dev)
    us_add = t(lambda: T.csp_attn_add(q, k, v, c, idx, cnt, 1, out=o))
    us_128 = t(lambda: torch.ops.chipmunk.csp_128_attn(q, k, v, idx, cnt))
    print(f"count {count:5d} steps {-(-count // 128):3d}: csp_attn_add {us_add:7.1f} us   csp_128_attn {us_128:7.1f} us   per tile-round {us_add / 4:6.2f} us", flush=True)
