"""world_size-2 `gloo` test of the multi-GPU host logic (head sharding + the single all-gather of O).
Runs on CPU: each rank computes its heads with the ORACLE (there is no CPU product path) and the
product's gather/sharding code reassembles the layer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, heads, q, k, v, o0, idx, cnt, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from chipmunk_b200 import parallel
    from oracle import chipmunk_oracle as oracle
    b, e = parallel.shard_heads(heads, world, rank)
    local = oracle.csp_attn(q[:, b:e], k[:, b:e], v[:, b:e], o0[:, b:e], idx[:, b:e], cnt[:, b:e], 1)
    full = parallel.all_gather_heads(local, heads)
    ret[rank] = full
    dist.destroy_process_group()


@pytest.mark.parametrize("heads", [4, 3])
def test_head_parallel_allgather_matches_single_process(oracle, heads):
    g = torch.Generator().manual_seed(0)
    B, N = 1, 200
    q, k, v, o0 = (torch.randn(B, heads, N, 128, generator=g).to(torch.bfloat16) for _ in range(4))
    G = 2
    idx, cnt = oracle.random_index_sets(B, heads, G, N, 64, g)
    ref = oracle.csp_attn(q, k, v, o0, idx, cnt, 1)
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, heads, q, k, v, o0, idx, cnt, ret), nprocs=world, join=True)
    for r in range(world):
        assert torch.equal(ret[r], ref), f"rank {r} reassembled a different layer"


def test_shard_ranges_cover_everything(cm):
    from chipmunk_b200.parallel import shard_range
    for n in (24, 36, 7, 1):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
