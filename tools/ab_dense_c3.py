import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chipmunk_b200 import torch_ops as T
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
def timeit(fn, n):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for H, N, n in ((24, 16384, 5), (24, 119056, 3)):
    q, k, v = (torch.randn(1, H, N, 128, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
    p = torch.rand(1, H, N, 1, device=dev, generator=g) * 1e-3 + 1e-5
    fl = 4.0 * N * N * 128 * H
    td = timeit(lambda: T._launch_dense(q, k, v, None), n)
    tc = timeit(lambda: T._launch_dense(q, k, v, p), n)
    print(f"{os.environ.get('CHIPMUNK_B200_LIB','current'):45s} N={N}: dense {td:8.3f} ms ({fl/td/1e9:5.0f} TF/s)   dense+colsum {tc:8.3f} ms ({fl/tc/1e9:5.0f})", flush=True)
    del q, k, v, p
