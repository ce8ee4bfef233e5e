import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chipmunk_b200 import torch_ops as T
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
def med(fn, n):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
H, N = 24, 16384
q, k, v = (torch.randn(1, H, N, 128, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
p = torch.rand(1, H, N, 1, device=dev, generator=g) * 1e-3 + 1e-5
s = (q[:, :2].float() @ k[:, :2].float().transpose(-1, -2)) * 128 ** -0.5
ro = torch.softmax(s, -1) @ v[:, :2].float()
o, cs, l = T._launch_dense(q, k, v, p)
err = float((o[:, :2].float() - ro).norm() / ro.norm())
fl = 4.0 * N * N * 128 * H
td = med(lambda: T._launch_dense(q, k, v, None), 9); tc = med(lambda: T._launch_dense(q, k, v, p), 9)
print(f"{os.environ.get('CHIPMUNK_B200_LIB','current(poly 1/4)'):40s} dense {td:.3f} ms ({fl/td/1e9:.0f} TF/s)  dense+colsum {tc:.3f} ms ({fl/tc/1e9:.0f})  o_rel {err:.2e}", flush=True)
