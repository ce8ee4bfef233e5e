#!/usr/bin/env python
"""Benchmark of the column-sparse DiT hot path (BASELINE.json metric / configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]

Workload ("flux_single_stream_block"): one SPARSE denoising step of one FLUX.1-dev single-stream
block on synthetic tensors: column-sparse delta attention over N = 4096 image + 512 text tokens,
H = 24 heads, d = 128, 83.5 % column sparsity (count = 112*round(0.165*N/112) keys per 192-query
group, the fused `csp_attn` path), then the column-sparse MLP (d = 3072, F = 12288, 70 % sparse:
3840 active neurons per 128-token block).  A step = csp_attn_add (cache + sparse delta,
one kernel) + csp_mlp_mm1 + csp_mlp_mm2, the three launches of the modules' sparse step.

metric = dense-equivalent TFLOP/s: FLOPs the DENSE block would need (4*N^2*d*H + 4*M*K*F)
divided by the sparse step time.  `value` has inputs resident in HBM; `e2e` feeds the same step
through the public ops from pinned HOST buffers with the H2D/D2H copies inside the timed region.

N > 1 GPUs (torchrun, one rank per GPU): every rank runs the step on its own sample (the path
shards over independent samples without any collective: weak scaling).  The head-parallel mode of
the reference's multi-GPU HunyuanVideo path is measured on the 720p attention shape and reported in the
`c3_hunyuan_attn` object, twice: heads sharded + ONE in-place NCCL all-gather of O, and the same layer with the
gather fused into the attention epilogue (NVLS multicast stores into a symmetric buffer, chipmunk_b200/parallel.py);
the two results are asserted bit-identical on every run.

--impl reference times the reference's own CPU implementation of the path -- its pure-PyTorch
dense branch (modules/attn.py:194, modules/mlp.py:34), restated in oracle/ -- on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ---------------------------------------------------------------------------------- workload
H, D = 24, 128
N_IMG, N_TXT = 4096, 512
NSEQ = N_IMG + N_TXT                      # 4608
ATTN_TOP_KEYS = 0.165                      # examples/flux chipmunk config: 83.5 % attention sparsity
ATTN_MULT = 112
ATTN_COUNT = ATTN_MULT * round(ATTN_TOP_KEYS * NSEQ / ATTN_MULT)   # 784
MLP_K, MLP_F = 3072, 12288
MLP_TOP = 0.30
MLP_COUNT = 256 * -(-int(MLP_TOP * MLP_F) // 256)                   # 3840
QG = 192

DENSE_FLOPS_ATTN = 4.0 * NSEQ * NSEQ * D * H
DENSE_FLOPS_MLP = 4.0 * NSEQ * MLP_K * MLP_F
DENSE_FLOPS = DENSE_FLOPS_ATTN + DENSE_FLOPS_MLP


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ---------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------- inputs
def make_indices(rows, n, count, gen, device):
    """`rows` independent uniformly random count-subsets of [0,n), ascending, padded to n columns."""
    out = torch.zeros(rows, n, dtype=torch.int32, device=device)
    step = 64
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        sel = torch.rand(r1 - r0, n, device=device, generator=gen).topk(count, dim=-1).indices
        out[r0:r1, :count] = sel.sort(dim=-1).values.int()
    return out


class Block:
    """Synthetic tensors of one FLUX single-stream block's sparse step."""

    def __init__(self, device, seed=0, heads=H, nseq=NSEQ, attn_count=ATTN_COUNT, with_mlp=True):
        g = torch.Generator(device=device).manual_seed(seed)
        bf = torch.bfloat16
        self.heads, self.nseq = heads, nseq
        G = (nseq + QG - 1) // QG
        self.q, self.k, self.v = (torch.randn(1, heads, nseq, D, device=device, generator=g).to(bf) for _ in range(3))
        self.o_cache = torch.randn(1, heads, nseq, D, device=device, generator=g).to(bf)
        self.o = self.o_cache.clone()
        self.a_idx = make_indices(heads * G, nseq, attn_count, g, device).view(1, heads, G, nseq)
        self.a_cnt = torch.full((1, heads, G), attn_count, dtype=torch.int32, device=device)
        self.with_mlp = with_mlp
        if with_mlp:
            M = nseq
            self.x = torch.randn(M, MLP_K, device=device, generator=g).to(bf)
            self.w1 = (0.02 * torch.randn(MLP_F, MLP_K, device=device, generator=g)).to(bf)
            self.b1 = (0.02 * torch.randn(MLP_F, device=device, generator=g)).to(bf)
            self.w2t = (0.02 * torch.randn(MLP_F, MLP_K, device=device, generator=g)).to(bf)
            self.pa_T = torch.randn(MLP_F, M, device=device, generator=g).to(bf)
            self.out_cache = torch.randn(M, MLP_K, device=device, generator=g).to(bf)
            self.m_idx = torch.stack([torch.randperm(MLP_F, device=device, generator=g) for _ in range(M // 128)]).int()
            self.m_cnt = torch.full((M // 128,), MLP_COUNT, dtype=torch.int32, device=device)
            self.packed = torch.empty(M, MLP_F, device=device, dtype=bf)


def attn_alg_bytes(heads, nseq, count):
    tiles = heads * ((nseq + QG - 1) // QG)
    return tiles * (count * (2 * D * 2 + 4) + 3 * QG * D * 2)


def mm1_alg_bytes(M, count):
    return (M // 128) * count * MLP_K * 2 + M * MLP_K * 2 + 2 * M * count * 2


def mm2_alg_bytes(M, count):
    return (M // 128) * count * MLP_K * 2 + M * count * 2 + 2 * M * MLP_K * 2


# ---------------------------------------------------------------------------------- our arm
def run_ours(args):
    import chipmunk_b200 as cm
    from chipmunk_b200 import torch_ops as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    hbm_gbs, tf_burst, tf_sus, peak_src = peaks()

    blk = Block(dev, seed=rank)
    M = NSEQ
    stream = torch.cuda.current_stream()

    def step(ev=None):
        if ev is not None: ev[0].record()
        T.csp_attn_add(blk.q, blk.k, blk.v, blk.o_cache, blk.a_idx, blk.a_cnt, 1, out=blk.o)   # SparseDiffAttn's sparse step
        if ev is not None: ev[1].record()
        T.mlp_mm1(blk.x, blk.w1, blk.packed, blk.b1, blk.pa_T, blk.m_idx, blk.m_cnt, True)
        if ev is not None: ev[2].record()
        T.mlp_mm2(blk.packed, blk.w2t, blk.out_cache, None, blk.m_idx, blk.m_cnt, False)
        if ev is not None: ev[3].record()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for i in range(args.steps):
            step(evs[i])
        e1.record()
        barrier()
        # keep the sampler alive long enough to see the load even for very short runs
        t_end = time.time() + 0.35
        while time.time() < t_end:
            step()
        torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * DENSE_FLOPS / (ms_step * 1e-3) / 1e12

    t_attn = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    t_mm1 = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    t_mm2 = statistics.mean(e[2].elapsed_time(e[3]) for e in evs)

    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_db = json.load(f)

    def roof(alg_bytes, ms, traffic=None):
        a = alg_bytes / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": round(a, 1), "peak": hbm_gbs, "unit": "GB/s", "frac": round(a / hbm_gbs, 4),
                "traffic": traffic, "peak_source": peak_src}

    kernels = {
        "csp_attn_add": (t_attn, roof(attn_alg_bytes(H, NSEQ, ATTN_COUNT), t_attn, traffic_db.get("csp_attn_add"))),
        "csp_mlp_mm1": (t_mm1, roof(mm1_alg_bytes(M, MLP_COUNT), t_mm1, traffic_db.get("csp_mlp_mm1"))),
        "csp_mlp_mm2": (t_mm2, roof(mm2_alg_bytes(M, MLP_COUNT), t_mm2, traffic_db.get("csp_mlp_mm2"))),
    }
    dominant = max(kernels, key=lambda k: kernels[k][0])
    roofline = dict(kernels[dominant][1])
    roofline["kernel"] = dominant
    roofline["launch_us"] = round(kernels[dominant][0] * 1e3, 1)
    if dominant.startswith("csp_mlp"):
        # The gathered weight rows are L2-resident, so `frac` (algorithmic bytes / HBM peak) can exceed 1: the limit these
        # kernels actually sit on is the SM's L2->SM ingress, ~43 B/clk/SM: 48 KB per 128x256x64 stage.
        stages = (M // 128) * (MLP_COUNT // 256) * (MLP_K // 64)
        roofline["sm_ingress"] = {"bytes_per_launch": stages * 49152, "achieved_gbs": round(stages * 49152 / (kernels[dominant][0] * 1e-3) / 1e9, 1),
                                  "peak_gbs_at_sm_clock": "43 B/clk x 148 SMs x clocks.sm_mhz", "B_per_clk_per_sm_peak": 43}
    sparse_flops = {"csp_attn_add": 4.0 * QG * ATTN_COUNT * D * H * ((NSEQ + QG - 1) // QG),
                    "csp_mlp_mm1": 2.0 * M * MLP_COUNT * MLP_K, "csp_mlp_mm2": 2.0 * M * MLP_COUNT * MLP_K}
    per_kernel = {k: {"us": round(v[0] * 1e3, 1), "gather_roofline_frac": v[1]["frac"],
                      "tensor_tflops": round(sparse_flops[k] / (v[0] * 1e-3) / 1e12, 1),
                      "tensor_frac_of_burst": round(sparse_flops[k] / (v[0] * 1e-3) / 1e12 / tf_burst, 4)}
                  for k, v in kernels.items()}

    # ---- e2e: same step through the public ops with pinned HOST inputs/outputs
    host_in = [t.cpu().pin_memory() for t in (blk.q, blk.k, blk.v, blk.x)]
    host_out = [torch.empty_like(blk.o, device="cpu").pin_memory(), torch.empty_like(blk.out_cache, device="cpu").pin_memory()]
    h2d = sum(t.numel() * t.element_size() for t in host_in)
    d2h = sum(t.numel() * t.element_size() for t in host_out)

    # Pipelined over three streams (H2D | compute | D2H) with double-buffered device inputs/outputs, the way a
    # serving loop would drive the ops: step i+1's inputs upload and step i-1's results download while step i
    # computes.  Every step still uploads ITS inputs and downloads ITS results inside the timed region.
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    inbuf = [[torch.empty_like(t) for t in (blk.q, blk.k, blk.v, blk.x)] for _ in range(2)]
    obuf = [torch.empty_like(blk.o) for _ in range(2)]
    ostage = [torch.empty_like(blk.out_cache) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_comp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(nsteps):
        for i in range(nsteps):
            j = i & 1
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_comp[j])              # step i-2 no longer reads inbuf[j]
                for dst, src in zip(inbuf[j], host_in):
                    dst.copy_(src, non_blocking=True)
                ev_in[j].record(s_in)
            stream.wait_event(ev_in[j])
            if i >= 2:
                stream.wait_event(ev_out[j])                 # step i-2's results have left obuf[j] / ostage[j]
            q_, k_, v_, x_ = inbuf[j]
            T.csp_attn_add(q_, k_, v_, blk.o_cache, blk.a_idx, blk.a_cnt, 1, out=obuf[j])
            cm.ops.mlp(x_, blk.w1, blk.b1, blk.w2t, blk.m_idx, blk.m_cnt, blk.pa_T, blk.out_cache, 6)
            ostage[j].copy_(blk.out_cache)
            ev_comp[j].record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[j])
                host_out[0].copy_(obuf[j], non_blocking=True)
                host_out[1].copy_(ostage[j], non_blocking=True)
                ev_out[j].record(s_out)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)

    e2e_run(4)
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item()) / args.steps
    e2e = {"value": round(world * DENSE_FLOPS / (ms_e2e * 1e-3) / 1e12, 2), "unit": "TFLOP/s-equiv",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e, 4),
           "pipelining": "3 streams, double-buffered; every step uploads its inputs and downloads its results"}

    extras = {}
    if not args.no_extras:
        extras.update(dense_gpu_baselines(blk, dev))
        extras["c3_hunyuan_attn"] = c3_attention(dev, world, rank)
    cpu = None
    if rank == 0 and world == 1 and not args.no_extras:
        cpu = cpu_reference(sample_seconds=12.0)

    clocks = clk.summary()
    if "sm_ingress" in roofline and clocks.get("sm_mhz"):
        pk = 43 * 148 * clocks["sm_mhz"] * 1e6 / 1e9
        roofline["sm_ingress"]["peak_gbs_at_sm_clock"] = round(pk, 1)
        roofline["sm_ingress"]["frac"] = round(roofline["sm_ingress"]["achieved_gbs"] / pk, 4)
    if rank == 0:
        line = {
            "metric": "dense-equivalent TFLOP/s of one column-sparse FLUX single-stream block step (attn + MLP)",
            "value": round(value, 2), "unit": "TFLOP/s-equiv", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "flux_single_stream_block", "seq": NSEQ, "heads": H, "head_dim": D, "d_model": MLP_K,
                       "mlp_hidden": MLP_F, "attn_sparsity": round(1 - ATTN_COUNT / NSEQ, 4), "attn_count": ATTN_COUNT,
                       "mlp_sparsity": round(1 - MLP_COUNT / MLP_F, 4), "mlp_count": MLP_COUNT,
                       "parallelism": f"dp{world} (independent samples, no collective)",
                       "l2": "per-step working set ~0.6 GB > 126 MB L2; no explicit flush"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 3 * args.steps,
            "clocks": clocks, "us_per_layer": {"attn": round(t_attn * 1e3, 1), "mlp": round((t_mm1 + t_mm2) * 1e3, 1)},
            "kernels": per_kernel,
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def _time(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def dense_gpu_baselines(blk, dev):
    """Dense library baselines on the same box: cuDNN/flash SDPA and cuBLAS MLP."""
    F = torch.nn.functional
    w2 = blk.w2t.t().contiguous()
    t_sdpa = _time(lambda: F.scaled_dot_product_attention(blk.q, blk.k, blk.v), 10)
    x3 = blk.x[None]
    t_mlp = _time(lambda: F.linear(F.gelu(F.linear(x3, blk.w1, blk.b1), approximate="tanh"), w2), 10)
    return {"dense_gpu_baseline": {"sdpa_us": round(t_sdpa * 1e3, 1), "sdpa_tflops": round(DENSE_FLOPS_ATTN / t_sdpa / 1e9, 1),
                                   "cublas_mlp_us": round(t_mlp * 1e3, 1), "cublas_mlp_tflops": round(DENSE_FLOPS_MLP / t_mlp / 1e9, 1)}}


def c3_attention(dev, world, rank):
    """HunyuanVideo 720p attention layer (N = 118800 + 256, 93 % sparse, count 8320).
    world == 1: whole layer on this GPU + dense SDPA baseline.
    world > 1: heads sharded across ranks + ONE NCCL all-gather of O (reference head_parallel.py:42-115)."""
    n, count = 119056, 8320
    hl = H // world if H % world == 0 else -(-H // world)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    bf = torch.bfloat16
    G = (n + QG - 1) // QG
    q, k, v = (torch.randn(1, hl, n, D, device=dev, generator=g).to(bf) for _ in range(3))
    o = torch.zeros(1, hl, n, D, device=dev, dtype=bf)
    idx = make_indices(hl * G, n, count, g, dev).view(1, hl, G, n)
    cnt = torch.full((1, hl, G), count, dtype=torch.int32, device=dev)
    res = {"seq": n, "count": count, "heads_per_gpu": hl}
    if world == 1:
        ts = sorted(_time(lambda: torch.ops.chipmunk.csp_attn(q, k, v, o, idx, cnt, 1), 1, warm=1 if i == 0 else 0) for i in range(5))
        t = ts[2]                                                         # median of 5 launches
        td = _time(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), 2, warm=1)
        dense = 4.0 * n * n * D * H
        res.update({"sparse_ms": round(t, 3), "sparse_ms_min_max": [round(ts[0], 3), round(ts[-1], 3)], "dense_sdpa_ms": round(td, 3), "speedup_vs_dense_sdpa": round(td / t, 2),
                    "dense_equiv_tflops": round(dense / t / 1e9, 1),
                    "gather_gbs": round(attn_alg_bytes(H, n, count) / t / 1e6, 1)})
        return res
    import torch.distributed as dist
    from chipmunk_b200 import parallel

    def layer():
        # csp_attn_add on this rank's heads, written into the gather buffer + ONE (in-place) all-gather of O
        return parallel.sparse_attention_head_parallel(q, k, v, o, idx, cnt, hl * world, fused=False)

    for _ in range(2):
        full = layer()
    # self-check outside the timed region: this rank's slice of the gathered layer is what its kernel wrote,
    # and every other slice arrived (finite, non-zero)
    from chipmunk_b200 import torch_ops as T
    mine = T.csp_attn_add(q, k, v, o, idx, cnt, 1)
    assert torch.equal(full[:, rank * hl:(rank + 1) * hl], mine), "head-parallel gather: own slice differs"
    assert bool(torch.isfinite(full.float()).all()) and all(
        float(full[:, r * hl:(r + 1) * hl].float().abs().sum()) > 0 for r in range(world)), "head-parallel gather: missing slice"
    del mine, full
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        layer()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_attn = _time(lambda: torch.ops.chipmunk.csp_attn(q, k, v, o, idx, cnt, 1), 3, warm=1)
    res.update({"layer_ms_max_over_ranks": round(float(t.item()), 3), "attn_only_ms_this_rank": round(t_attn, 3),
                "allgather_bytes_per_rank": o.numel() * 2, "collective": "1x ncclAllGather of O per layer"})
    # the same layer with the all-gather fused into the kernel epilogue (NVLS multicast stores, parallel.py)
    try:
        def layer_fused():
            return parallel.sparse_attention_head_parallel_fused(q, k, v, o, idx, cnt, hl * world)

        # all ranks agree (all_reduce MIN inside) that the symmetric buffer + NVLS multicast exist before anyone enters a device barrier
        if not parallel.fused_gather_available((world, 1, hl, n, D), bf, dev, None):
            raise RuntimeError("symmetric memory / NVLS multicast not available on every rank")
        full_f = layer_fused()
        full_n = layer()
        dist.barrier(); torch.cuda.synchronize()
        assert torch.equal(full_f, full_n), "fused multicast gather differs from the NCCL all-gather"
        del full_n
        for _ in range(2):
            layer_fused()
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            layer_fused()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        tf = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        res["fused_multicast_layer_ms_max_over_ranks"] = round(float(tf.item()), 3)
    except Exception as e:  # noqa: BLE001  (no NVLS on this box: the NCCL number above stands)
        res["fused_multicast_layer_ms_max_over_ranks"] = None
        res["fused_multicast_error"] = repr(e)[:200]
    return res


# ---------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(sample_seconds=12.0, heads=6, rows=1152):
    """The reference's dense PyTorch branch on the host cores, on a bounded sample of the workload:
    `heads` of the 24 heads (full 4608x4608 attention each) and `rows` of the 4608 MLP token rows."""
    from oracle import chipmunk_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(1, heads, NSEQ, D, generator=g) for _ in range(3))
    x = torch.randn(1, rows, MLP_K, generator=g)
    w1 = 0.02 * torch.randn(MLP_F, MLP_K, generator=g); b1 = torch.zeros(MLP_F)
    w2 = 0.02 * torch.randn(MLP_K, MLP_F, generator=g); b2 = torch.zeros(MLP_K)
    flops = 4.0 * NSEQ * NSEQ * D * heads + 4.0 * rows * MLP_K * MLP_F
    oracle.dense_block_cpu(q, k, v, x, w1, b1, w2, b2)          # warm-up
    t0, n = time.time(), 0
    while True:
        oracle.dense_block_cpu(q, k, v, x, w1, b1, w2, b2)
        n += 1
        if time.time() - t0 > sample_seconds or n >= 20:
            break
    dt = (time.time() - t0) / n
    return {"value": round(flops / dt / 1e12, 4), "unit": "TFLOP/s-equiv", "cores": cores, "kind": "port",
            "sample": f"{heads}/24 heads of dense SDPA + {rows}/4608 token rows of dense MLP, fp32, {n} reps "
                      f"({dt:.2f} s each); the reference's pure-PyTorch dense branch (modules/attn.py:194, mlp.py:34)",
            "seconds_per_full_block_extrapolated": round(dt * DENSE_FLOPS / flops, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import chipmunk_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    heads, rows = 6, 1152
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(1, heads, NSEQ, D, generator=g) for _ in range(3))
    x = torch.randn(1, rows, MLP_K, generator=g)
    w1 = 0.02 * torch.randn(MLP_F, MLP_K, generator=g); b1 = torch.zeros(MLP_F)
    w2 = 0.02 * torch.randn(MLP_K, MLP_F, generator=g); b2 = torch.zeros(MLP_K)
    flops = 4.0 * NSEQ * NSEQ * D * heads + 4.0 * rows * MLP_K * MLP_F
    for _ in range(max(1, min(args.warmup, 2))):
        oracle.dense_block_cpu(q, k, v, x, w1, b1, w2, b2)
    t0 = time.time()
    for _ in range(args.steps):
        oracle.dense_block_cpu(q, k, v, x, w1, b1, w2, b2)
    dt = (time.time() - t0) / args.steps
    val = round(flops / dt / 1e12, 4)
    sample = (f"each step = {heads}/24 heads of dense SDPA + {rows}/4608 token rows of the dense MLP (1/4 of the block), "
              "fp32 on all host threads")
    print(json.dumps({
        "impl": "reference",
        "metric": "dense-equivalent TFLOP/s of one column-sparse FLUX single-stream block step (attn + MLP)",
        "value": val, "unit": "TFLOP/s-equiv", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "flux_single_stream_block", "seq": NSEQ, "heads": H, "head_dim": D, "d_model": MLP_K,
                   "mlp_hidden": MLP_F, "note": "reference's dense PyTorch branch on CPU; bounded sample per step"},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s-equiv", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s-equiv", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip dense GPU baselines, C3 attention and the CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
