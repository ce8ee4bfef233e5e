"""One attention layer through a whole 50-step sample (BASELINE.json configs[3], single GPU part): the
SparseDiffAttn module driven by the reference's HunyuanVideo schedule (examples/hunyuan/chipmunk-config.yml:
top_keys 0.05 + 1 % random keys, full steps {0, 1, 10, 40}, mask recomputed on every full step, bit-packed indices),
every step timed with CUDA events, against 50 dense SDPA calls.

    python tools/sample_schedule.py [--seq 119056] [--heads 24] [--out profiles/r01_sample_schedule.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chipmunk_b200 as cm  # noqa: E402
from chipmunk_b200.util import GLOBAL_CONFIG, LayerCounter  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, default=119056)
    ap.add_argument("--heads", type=int, default=24)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    # under torchrun: BASELINE.json configs[3], the schedule with the layer head-parallel over the GPUs of one box
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        assert a.heads % world == 0
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    total_heads = a.heads
    a.heads = a.heads // world
    GLOBAL_CONFIG["steps"] = a.steps
    GLOBAL_CONFIG["attn"].update({"is_enabled": True, "top_keys": 0.05, "random_keys": 0.01, "local_voxels": 0,
                                  "local_1d_window": 0, "first_n_dense_layers": 2, "recompute_mask": True,
                                  "should_compress_indices": True, "full_step_schedule": {0, 1, 10, 40},
                                  "pad_qkv_before_kernel": True, "counts_multiple_of": 128})
    GLOBAL_CONFIG["mlp"]["is_enabled"] = False
    counter = LayerCounter(num_layers=1, num_sparse_submodules_per_layer=1)
    layer = cm.SparseDiffAttn(layer_num=2, layer_counter=counter)        # layer_num >= first_n_dense_layers
    if world > 1:
        from chipmunk_b200.parallel import HeadParallelAttn
        layer = HeadParallelAttn(layer, total_heads)

    if world > 1:
        # one-time initialisation outside the timed steps: NCCL communicator, symmetric-memory rendezvous of the output buffer
        from chipmunk_b200 import parallel
        dist.all_reduce(torch.ones(1, device=dev))
        parallel.all_gather_heads(torch.zeros(1, a.heads, 8, 128, device=dev, dtype=torch.bfloat16), total_heads)
        parallel.fused_gather_available((world, 1, a.heads, a.seq, 128), torch.bfloat16, dev, None)
        torch.cuda.synchronize()
    g = torch.Generator(device=dev).manual_seed(rank)
    q, k, v = (torch.randn(1, a.heads, a.seq, 128, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
    times, kinds = [], []
    for step in range(a.steps - 1):          # the reference's odometer rewinds before the last coordinate
        full = counter.should_do_full_attn_step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = layer(q, k, v)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        times.append(ms)
        kinds.append("full" if full else "sparse")
        del o
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    t_dense = e0.elapsed_time(e1) / 3
    full_ms = [t for t, kd in zip(times, kinds) if kd == "full"]
    sparse_ms = [t for t, kd in zip(times, kinds) if kd == "sparse"]
    if rank != 0:
        dist.destroy_process_group()
        return
    res = {"n_gpus": world, "workload": "one attention layer through a 49-step HunyuanVideo schedule (full steps 0,1,10,40; recompute_mask; packed indices)",
           "seq": a.seq, "heads": total_heads, "heads_per_gpu": a.heads, "n_full": len(full_ms), "n_sparse": len(sparse_ms),
           "full_step_ms": [round(t, 2) for t in full_ms],
           "sparse_step_ms_median": round(sorted(sparse_ms)[len(sparse_ms) // 2], 3),
           "sparse_step_ms_minmax": [round(min(sparse_ms), 3), round(max(sparse_ms), 3)],
           "layer_total_ms": round(sum(times), 1), "dense_sdpa_ms_per_step": round(t_dense, 2),   # cuDNN SDPA on the same (per-GPU) heads
           "dense_total_ms": round(t_dense * len(times), 1), "speedup_vs_dense_sdpa": round(t_dense * len(times) / sum(times), 2)}
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            f.write(json.dumps(res) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
