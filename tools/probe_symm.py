"""Probe (test infrastructure): does torch symmetric memory work on this box, with peer pointers and an NVLS
multicast address?  torchrun --nproc-per-node 2 tools/probe_symm.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    t = symm.empty(world, 1024, dtype=torch.bfloat16, device=torch.device("cuda", local))
    h = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; world", h.world_size, "rank", h.rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs],
          "multicast_ptr", hex(getattr(h, "multicast_ptr", 0) or 0), "signal_pad_ptrs", len(h.signal_pad_ptrs), flush=True)
    t.zero_()
    h.barrier()
    # every rank writes its row into every peer's buffer through the peer mapping
    for p in range(world):
        peer = h.get_buffer(p, (world, 1024), torch.bfloat16)
        peer[rank].fill_(float(rank + 1))
    h.barrier()
    torch.cuda.synchronize()
    print(rank, "rows seen locally:", [float(t[r, 0]) for r in range(world)], flush=True)
except Exception as e:  # noqa: BLE001
    print(rank, "symmetric memory unavailable:", repr(e)[:300], flush=True)
dist.destroy_process_group()
