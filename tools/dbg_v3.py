import sys, torch
sys.path.insert(0, ".")
import chipmunk_b200  # noqa
from oracle import chipmunk_oracle as oracle
B, H, N, count = 1, 1, int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(0)
q, k, v = (torch.randn(B, H, N, 128, generator=g).to(torch.bfloat16) for _ in range(3))
G = (N + 191) // 192
idx, cnt = oracle.random_index_sets(B, H, G, N, count, g)
full = torch.zeros(B, H, G, N, dtype=torch.int32); full[..., :count] = idx
out = torch.ops.chipmunk.csp_128_attn(q.cuda(), k.cuda(), v.cuda(), full.cuda(), cnt.cuda())
torch.cuda.synchronize()
ref = oracle.csp_128_attn(q, k, v, full, cnt)
print(N, count, "relerr", ((out.float().cpu() - ref.float()).norm() / ref.float().norm()).item(), flush=True)
