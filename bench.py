#!/usr/bin/env python
"""Benchmark of the column-sparse DiT hot path (BASELINE.json metric; headline = configs[2], the shape its targets are
quoted on and the largest configuration that fits one GPU; configs[1] is measured beside it in the same run).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]

Workload ("hunyuan_720p_attention_layer", BASELINE.json configs[2]/[3]): one SPARSE denoising step of one HunyuanVideo
720x1280x129 attention layer on synthetic tensors -- N = 118 800 video + 256 text tokens = 119 056, H = 24 heads,
d = 128, 93 % column sparsity (count = 8320 = 65 x 128 keys per 192-query group, reference config top_keys 0.05 + local
/ random columns), bit-packed mask storage (`should_compress_indices`).  A step is what `SparseDiffAttn` runs on a
sparse step (reference src/chipmunk/modules/attn.py:172-190):
        indices, counts = bitmask_to_indices(packed mask)      (reference: bitunpack + mask_to_indices)
        o = csp_attn_add(q, k, v, o_cache, indices, counts)    (reference: clone + csp_attn, delta added in the epilogue)
two kernel launches.  The MLP of this model is not sparsified by the reference (`mlp.is_enabled: false`).

metric = dense-equivalent TFLOP/s: FLOPs of the DENSE layer (4 N^2 d H = 174 TFLOP) / sparse step time.
`value`: inputs resident in HBM.  `e2e`: the same step from pinned HOST buffers (q, k, v uploaded, o downloaded inside
the timed region, three streams).  `roofline`: the attention kernel against the indexed-KV gather roofline.

N > 1 GPUs (torchrun, one rank per GPU): STRONG scaling of the same layer -- heads sharded 24/N per rank, every rank's
kernel stores its rows straight into all GPUs' copies of the output through the NVSwitch (multicast stores fused into the
epilogue; `parallel.sparse_attention_head_parallel`), i.e. the reference's head-parallel HunyuanVideo path with ONE
gather of O instead of two all_to_all + all_gather (hyvideo/modules/head_parallel.py:42-115).  The NCCL variant of the
same layer (kernel + in-place ncclAllGather) is timed beside it and asserted bit-identical.

Extras at N = 1 (objects in the same JSON line): the FLUX single-stream block of configs[1] (`c2_flux_block`:
csp_attn_add + csp_mlp_mm1 + csp_mlp_mm2 with per-kernel rooflines, cuDNN-SDPA / cuBLAS baselines and the reference's
own Triton mm1/mm2 kernels raced on the same box), the full step of this layer (`c3_full_step`: one-pass dense +
column sums, column selection, cache build), dense SDPA at this shape, and an oracle check of sampled tiles.

--impl reference times the reference's own CPU implementation of the path -- its pure-PyTorch dense branch
(modules/attn.py:194), restated in oracle/ -- on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib.util
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

# ---------------------------------------------------------------------------------- workloads
H, D, QG = 24, 128, 192
# configs[2]: HunyuanVideo 720p attention layer
C3_N = 118800 + 256
C3_COUNT = 8320
C3_DENSE_FLOPS = 4.0 * C3_N * C3_N * D * H
# configs[1]: FLUX.1-dev single-stream block
N_IMG, N_TXT = 4096, 512
NSEQ = N_IMG + N_TXT                      # 4608
ATTN_TOP_KEYS = 0.165                      # examples/flux chipmunk config: 83.5 % attention sparsity
ATTN_MULT = 112
ATTN_COUNT = ATTN_MULT * round(ATTN_TOP_KEYS * NSEQ / ATTN_MULT)   # 784
MLP_K, MLP_F = 3072, 12288
MLP_TOP = 0.30
MLP_COUNT = 256 * -(-int(MLP_TOP * MLP_F) // 256)                   # 3840
C2_DENSE_FLOPS_ATTN = 4.0 * NSEQ * NSEQ * D * H
C2_DENSE_FLOPS_MLP = 4.0 * NSEQ * MLP_K * MLP_F
METRIC = "dense-equivalent TFLOP/s of one column-sparse HunyuanVideo-720p attention layer step (bitmask->indices + delta attention)"


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


def make_packed_mask(heads, groups, n, count, gen, device):
    """Bit-packed bool mask [1, heads, groups, n] with exactly `count` uniformly random columns per row: what
    SparseDiffAttn stores after a full step (reference modules/attn.py:135-138)."""
    from chipmunk_b200 import torch_ops as T
    mask = torch.zeros(1, heads, groups, n, dtype=torch.bool, device=device)
    flat = mask.view(heads * groups, n)
    for r0 in range(0, heads * groups, 64):
        r1 = min(heads * groups, r0 + 64)
        sel = torch.rand(r1 - r0, n, device=device, generator=gen).topk(count, dim=-1).indices
        flat[r0:r1].scatter_(-1, sel, True)
    packed, shape = T.bitpack(mask)
    return packed, shape


def attn_alg_bytes(heads, nseq, count):
    tiles = heads * ((nseq + QG - 1) // QG)
    return tiles * (count * (2 * D * 2 + 4) + 3 * QG * D * 2)          # K+V rows + index, Q read, cache read, O write


def mm1_alg_bytes(M, count):
    return (M // 128) * count * MLP_K * 2 + M * MLP_K * 2 + 2 * M * count * 2


def mm2_alg_bytes(M, count):
    return (M // 128) * count * MLP_K * 2 + M * count * 2 + 2 * M * MLP_K * 2


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


def _roof(alg_bytes, ms, hbm_gbs, peak_src, traffic=None, bound="hbm"):
    a = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": bound, "achieved": round(a, 1), "peak": hbm_gbs, "unit": "GB/s", "frac": round(a / hbm_gbs, 4),
            "traffic": traffic, "peak_source": peak_src}


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, BEFORE the pinned host buffers of the e2e
    leg are allocated (first touch places them on that node).  With eight ranks on one box, staging buffers that all sit
    on one socket make every H2D / D2H copy of the other socket's GPUs cross the inter-socket link (round 1: e2e got
    slower with more GPUs).  Returns a short description for the JSON line; never fails the bench."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": None, "note": "no NUMA affinity reported for the GPU"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus_bound": len(allowed)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "note": repr(e)[:80]}


# ---------------------------------------------------------------------------------- our arm
def run_ours(args):
    import chipmunk_b200 as cm  # noqa: F401  (registers torch.ops.chipmunk, loads the library: no fallback)
    from chipmunk_b200 import parallel
    from chipmunk_b200 import torch_ops as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    assert H % world == 0, "the head-parallel layer needs 24 % n_gpus == 0"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    hbm_gbs, tf_burst, tf_sus, peak_src = peaks()
    bf = torch.bfloat16
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- this rank's heads of the layer (same seed everywhere: rank r owns heads [r hl, (r+1) hl))
    hl = H // world
    n, count = C3_N, C3_COUNT
    G = (n + QG - 1) // QG
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    q, k, v = (torch.randn(1, hl, n, D, device=dev, generator=g).to(bf) for _ in range(3))
    o_cache = torch.randn(1, hl, n, D, device=dev, generator=g).to(bf)
    packed, mshape = make_packed_mask(hl, G, n, count, g, dev)
    o_local = torch.empty(1, hl, n, D, device=dev, dtype=bf)
    stream = torch.cuda.current_stream()
    use_fused = world > 1 and parallel.fused_gather_available((world, 1, hl, n, D), bf, dev, None)

    def step(ev=None, fused=None):
        if ev is not None: ev[0].record()
        inds, cnts = T.bitmask_to_indices(packed, mshape, 128, QG)
        if ev is not None: ev[1].record()
        if world == 1:
            out = T.csp_attn_add(q, k, v, o_cache, inds, cnts, 1, out=o_local)
        else:
            out = parallel.sparse_attention_head_parallel(q, k, v, o_cache, inds, cnts, H, fused=use_fused if fused is None else fused)
        if ev is not None: ev[2].record()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, fn, record=False):
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(nsteps)] if record else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(nsteps):
            fn(evs[i]) if record else fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / nsteps, evs

    for _ in range(warmup):
        step()
    with ClockSampler(local) as clk:
        ms_step, evs = timed(steps, step, record=True)
        t_end = time.time() + 0.3            # keep the sampler on a loaded GPU for very short runs
        while time.time() < t_end:
            step()
        torch.cuda.synchronize()
    value = C3_DENSE_FLOPS / (ms_step * 1e-3) / 1e12
    t_m2i = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    t_attn = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)

    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_db = json.load(f)
    # the attention launch of THIS rank (hl heads); at N > 1 its time includes the multicast / NCCL gather of O
    roofline = _roof(attn_alg_bytes(hl, n, count), t_attn, hbm_gbs, peak_src, traffic_db.get("c3_csp_attn_add") if world == 1 else None)
    roofline.update({"note": "SURVEY 8d gather roofline: algorithmic indexed-KV bytes / HBM peak.  `traffic` (ncu DRAM bytes per launch) is ~7x "
                             "smaller: a head's K/V (61 MB) stays in the 126 MB L2 under head-major tile order, so the gather is served by L2 "
                             "and the kernel is bound by its S -> softmax -> P.V chain (tensor pipe 63 % active) and by the 1 kW power "
                             "cap (it runs at ~1.6 of 1.965 GHz: profiles/r02_probe_mma_power.txt), not by HBM",
                     "kernel": "attn::attn_kernel<QUAD=true> (csp_attn_add)" + (" + fused multicast gather of O" if use_fused else (" + ncclAllGather" if world > 1 else "")),
                     "launch_us": round(t_attn * 1e3, 1), "algorithmic_bytes_per_launch": attn_alg_bytes(hl, n, count),
                     "tensor_tflops": round(4.0 * QG * count * D * hl * G / (t_attn * 1e-3) / 1e12, 1),
                     "tensor_frac_of_burst_peak": round(4.0 * QG * count * D * hl * G / (t_attn * 1e-3) / 1e12 / tf_burst, 4)})
    m2i_bytes = packed.numel() + hl * G * (count * 4 + 4)
    kernels = {"bitmask_to_indices": {"us": round(t_m2i * 1e3, 1), "algorithmic_MB": round(m2i_bytes / 1e6, 1),
                                      "GB/s": round(m2i_bytes / t_m2i / 1e6, 1), "frac_of_hbm_peak": round(m2i_bytes / t_m2i / 1e6 / hbm_gbs, 4)},
               "csp_attn_add": {"us": round(t_attn * 1e3, 1), "gather_roofline_frac": roofline["frac"]}}
    # B200 option `attn.keep_indices_resident` (off by default = the reference's behaviour): the index lists of the last full
    # step stay in HBM (0.5 GB per layer at 720p), a sparse step is the attention launch alone.  Reported beside the headline.
    resident = {"ms_per_step": round(t_attn, 4), "value": round(C3_DENSE_FLOPS / (t_attn * 1e-3) / 1e12, 2), "unit": "TFLOP/s-equiv",
                "index_bytes_per_layer": int(hl * G * (n + 1) * 4),
                "note": "attn.keep_indices_resident: index lists kept in HBM between steps instead of re-derived from the bit mask"}

    # ---- e2e: the same step from pinned HOST buffers (q, k, v up; this rank's o down), H2D | compute | D2H streams
    numa = bind_to_gpu_numa_node(local)
    host_in = [torch.empty(t.shape, dtype=bf).pin_memory() for t in (q, k, v)]
    for hb, t in zip(host_in, (q, k, v)):
        hb.copy_(t)
    host_out = torch.empty(o_local.shape, dtype=bf).pin_memory()
    h2d = world * sum(t.numel() * 2 for t in host_in)
    d2h = world * host_out.numel() * 2
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    inbuf = [[torch.empty_like(t) for t in (q, k, v)] for _ in range(2)]
    ostage = [torch.empty_like(o_local) for _ in range(2)]
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
                stream.wait_event(ev_out[j])                 # step i-2's result has left ostage[j]
            q_, k_, v_ = inbuf[j]
            inds, cnts = T.bitmask_to_indices(packed, mshape, 128, QG)
            if world == 1:
                T.csp_attn_add(q_, k_, v_, o_cache, inds, cnts, 1, out=ostage[j])
            else:
                full = parallel.sparse_attention_head_parallel(q_, k_, v_, o_cache, inds, cnts, H, fused=use_fused)
                ostage[j].copy_(full[:, rank * hl:(rank + 1) * hl])      # every rank returns its heads of the gathered layer
            ev_comp[j].record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[j])
                host_out.copy_(ostage[j], non_blocking=True)
                ev_out[j].record(s_out)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)

    e2e_run(2)
    e2e_steps = max(4, min(steps, 10))
    ms_e2e, _ = timed(1, lambda: e2e_run(e2e_steps))
    ms_e2e /= e2e_steps
    e2e = {"value": round(C3_DENSE_FLOPS / (ms_e2e * 1e-3) / 1e12, 2), "unit": "TFLOP/s-equiv",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e, 3), "steps": e2e_steps, "host_numa": numa,
           "note": "q, k, v of every step uploaded from pinned host memory and o downloaded inside the timed region "
                   "(3 streams, double-buffered); PCIe-bound: %.2f GB per step and GPU" % ((h2d + d2h) / world / 1e9)}
    del inbuf, ostage, host_in, host_out

    extras = {}
    if world > 1:
        # the same layer through NCCL (kernel -> in-place ncclAllGather), beside the fused multicast path
        full_n = step(fused=False)
        if use_fused:
            full_f = step(fused=True)
            barrier()
            assert torch.equal(full_f, full_n), "fused multicast gather differs from the NCCL all-gather"
            del full_f
        mine = T.csp_attn_add(q, k, v, o_cache, *T.bitmask_to_indices(packed, mshape, 128, QG), 1)
        assert torch.equal(full_n[:, rank * hl:(rank + 1) * hl], mine), "head-parallel gather: own slice differs"
        del full_n, mine
        for _ in range(2):
            step(fused=False)
        ms_nccl, _ = timed(max(3, min(steps, 10)), lambda: step(fused=False))
        t_alone = _time(lambda: T.csp_attn_add(q, k, v, o_cache, *T.bitmask_to_indices(packed, mshape, 128, QG), 1, out=o_local), 3, warm=1)
        ms_peers = None
        if parallel.fused_gather_available((world, 1, hl, n, D), bf, dev, None, mode="peers"):
            full_p = step(fused="peers")
            full_n = step(fused=False)
            barrier()
            assert torch.equal(full_p, full_n), "peer-store fused gather differs from the NCCL all-gather"
            del full_p, full_n
            ms_peers, _ = timed(max(3, min(steps, 10)), lambda: step(fused="peers"))
        extras["head_parallel"] = {"layer_ms_fused_peer_stores": None if ms_peers is None else round(ms_peers, 3), "default_path": "fused multicast epilogue (NVLS)" if use_fused else "NCCL all-gather (no NVLS multicast on this box)",
                                   "layer_ms_default": round(ms_step, 3), "layer_ms_nccl_allgather": round(ms_nccl, 3),
                                   "kernels_only_ms_this_rank": round(t_alone, 3),
                                   "allgather_bytes_per_rank": o_local.numel() * 2,
                                   "bit_identical_fused_vs_nccl": bool(use_fused)}
    if rank == 0 and world == 1 and not args.no_extras:
        extras["oracle_check"] = oracle_check(T, q, k, v, o_cache, packed, mshape, o_local, dev)
        td = _time(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), 2, warm=1)
        extras["dense_gpu_baseline"] = {"sdpa_ms": round(td, 2), "sdpa_tflops": round(C3_DENSE_FLOPS / td / 1e9, 1),
                                        "speedup_sparse_step_vs_sdpa": round(td / ms_step, 2),
                                        "speedup_attention_kernel_vs_sdpa": round(td / t_attn, 2)}
        extras["c3_full_step"] = c3_full_step(T, q, k, v, dev, td)
        del q, k, v, o_cache, packed, o_local
        torch.cuda.empty_cache()
        extras["c2_flux_block"] = c2_flux_block(T, dev, hbm_gbs, tf_burst, peak_src, traffic_db)
        extras["cpu_baseline"] = cpu_reference(sample_seconds=12.0)

    clocks = clk.summary()
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "TFLOP/s-equiv", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "hunyuan_720p_attention_layer", "seq": n, "heads": H, "head_dim": D, "attn_count": count,
                       "attn_sparsity": round(1 - count / n, 4), "mask_storage": "bit-packed (bitmask_to_indices every step)",
                       "parallelism": (f"head-parallel hp{world}: {hl} heads per GPU, one gather of O "
                                       + ("fused into the kernel epilogue (NVLS multicast)" if use_fused else "(ncclAllGather)")) if world > 1 else "1 GPU",
                       "l2": "per-step working set (q,k,v,o,cache 3.7 GB + 7.1 GB of indices) >> 126 MB L2: every step streams from HBM"},
            "roofline": roofline, "cpu_baseline": extras.pop("cpu_baseline", None), "e2e": e2e, "gpu_launches": 2 * steps,
            "clocks": clocks, "us_per_layer": {"bitmask_to_indices": round(t_m2i * 1e3, 1), "attn": round(t_attn * 1e3, 1)},
            "kernels": kernels, "sparse_step_indices_resident": resident,
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def oracle_check(T, q, k, v, o_cache, packed, mshape, o_local, dev):
    """Outside the timed region: sampled tiles of the benchmarked launch against the CPU oracle (exact rounding points)."""
    from oracle import chipmunk_oracle as oracle
    inds, cnts = T.bitmask_to_indices(packed, mshape, 128, QG)
    T.csp_attn_add(q, k, v, o_cache, inds, cnts, 1, out=o_local)
    torch.cuda.synchronize()
    n = q.shape[2]
    G = inds.shape[2]
    worst, checked = 0.0, 0
    for h, gi in ((0, 0), (3, 147), (11, 310), (17, 555), (23, G - 1)):
        h = min(h, q.shape[1] - 1)
        r0, r1 = gi * QG, min((gi + 1) * QG, n)
        qg = torch.zeros(1, 1, QG, D, dtype=torch.bfloat16)
        qg[0, 0, : r1 - r0] = q[0, h, r0:r1].cpu()
        delta = oracle.csp_128_attn(qg, k[:, h:h + 1].cpu(), v[:, h:h + 1].cpu(), inds[:, h:h + 1, gi:gi + 1].cpu(), cnts[:, h:h + 1, gi:gi + 1].cpu())
        ref = (o_cache[0, h, r0:r1].cpu().float() + delta[0, 0, : r1 - r0].float()).to(torch.bfloat16).float()
        got = o_local[0, h, r0:r1].cpu().float()
        err = float((got - ref).norm() / ref.norm())
        assert err <= 4e-3, f"bench self-check against the oracle failed at head {h} group {gi}: {err:.3e}"
        worst, checked = max(worst, err), checked + 1
    return {"tiles_checked_vs_cpu_oracle": checked, "worst_rel_fro_err": round(worst, 6), "tolerance": 4e-3}


def c3_full_step(T, q, k, v, dev, t_sdpa):
    """The FULL step of the same layer (reference modules/attn.py:111-170): one-pass dense attention + column sums,
    column selection (top-k + 1 % random -> packed mask + indices), cache = dense - sparse."""
    n = q.shape[2]
    o, l = torch.ops.chipmunk.dense_attn(q, k, v)
    res = {}
    res["dense_attn_ms"] = round(_time(lambda: torch.ops.chipmunk.dense_attn(q, k, v), 2, warm=0), 2)
    res["dense_colsum_attn_ms"] = round(_time(lambda: torch.ops.chipmunk.dense_colsum_attn(q, k, v, l), 2, warm=1), 2)
    o, cs, l = torch.ops.chipmunk.dense_colsum_attn(q, k, v, l)
    tk = 128 * round(0.05 * n / 128)
    res["select_columns_ms"] = round(_time(lambda: T.select_columns(cs, tk, 128, 0.01, None, None, 7, QG), 2, warm=1), 3)
    _, _, inds, cnts = T.select_columns(cs, tk, 128, 0.01, None, None, 7, QG)
    del cs
    res["cache_build_csp_attn_add_ms"] = round(_time(lambda: T.csp_attn_add(q, k, v, o, inds, cnts, -1), 2, warm=1), 3)
    tot = res["dense_colsum_attn_ms"] + res["select_columns_ms"] + res["cache_build_csp_attn_add_ms"]
    res.update({"full_step_ms": round(tot, 2), "dense_sdpa_ms": round(t_sdpa, 2), "full_step_over_sdpa": round(tot / t_sdpa, 3),
                "dense_tflops": round(C3_DENSE_FLOPS / res["dense_attn_ms"] / 1e9, 1),
                "dense_colsum_tflops_equiv": round(C3_DENSE_FLOPS / res["dense_colsum_attn_ms"] / 1e9, 1), "top_keys": tk})
    return res


def _load_triton_ref(name):
    path = os.path.join(ROOT, "oracle", "_ref", "triton_ref", name + ".py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("chipmunk_ref_triton_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def c2_flux_block(T, dev, hbm_gbs, tf_burst, peak_src, traffic_db):
    """BASELINE.json configs[1]: one sparse step of a FLUX.1-dev single-stream block (N = 4608, 84 % attention / 70 % MLP
    sparsity): csp_attn_add + csp_mlp_mm1 (fused cache update) + csp_mlp_mm2, against cuDNN SDPA, cuBLAS and the
    reference's own Triton kernels on the same box."""
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    G = (NSEQ + QG - 1) // QG
    M = NSEQ
    q, k, v = (torch.randn(1, H, NSEQ, D, device=dev, generator=g).to(bf) for _ in range(3))
    o_cache = torch.randn(1, H, NSEQ, D, device=dev, generator=g).to(bf)
    o = torch.empty_like(o_cache)
    a_idx = make_indices(H * G, NSEQ, ATTN_COUNT, g, dev).view(1, H, G, NSEQ)
    a_cnt = torch.full((1, H, G), ATTN_COUNT, dtype=torch.int32, device=dev)
    x = torch.randn(M, MLP_K, device=dev, generator=g).to(bf)
    w1 = (0.02 * torch.randn(MLP_F, MLP_K, device=dev, generator=g)).to(bf)
    b1 = (0.02 * torch.randn(MLP_F, device=dev, generator=g)).to(bf)
    w2t = (0.02 * torch.randn(MLP_F, MLP_K, device=dev, generator=g)).to(bf)
    pa_T = torch.randn(MLP_F, M, device=dev, generator=g).to(bf)
    out_cache = torch.randn(M, MLP_K, device=dev, generator=g).to(bf)
    m_idx = torch.stack([torch.randperm(MLP_F, device=dev, generator=g) for _ in range(M // 128)]).int()
    m_cnt = torch.full((M // 128,), MLP_COUNT, dtype=torch.int32, device=dev)
    packed = torch.empty(M, MLP_F, device=dev, dtype=bf)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def blk(ev=None):
        if ev: ev[0].record()
        T.csp_attn_add(q, k, v, o_cache, a_idx, a_cnt, 1, out=o)
        if ev: ev[1].record()
        T.mlp_mm1(x, w1, packed, b1, pa_T, m_idx, m_cnt, True)
        if ev: ev[2].record()
        T.mlp_mm2(packed, w2t, out_cache, None, m_idx, m_cnt, False)
        if ev: ev[3].record()

    for _ in range(5):
        blk()
    iters = 200                                                        # ~65 ms of back-to-back steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(iters)]
    torch.cuda.synchronize()
    for i in range(iters):
        blk(evs[i])
    torch.cuda.synchronize()
    ms_blk = evs[0][0].elapsed_time(evs[-1][3]) / iters
    t = [statistics.mean(e[j].elapsed_time(e[j + 1]) for e in evs) for j in range(3)]
    # cold-L2 variant: 192 MB written before every launch
    cold = []
    for fn in (lambda: T.csp_attn_add(q, k, v, o_cache, a_idx, a_cnt, 1, out=o),
               lambda: T.mlp_mm1(x, w1, packed, b1, pa_T, m_idx, m_cnt, True),
               lambda: T.mlp_mm2(packed, w2t, out_cache, None, m_idx, m_cnt, False)):
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        cold.append(sorted(ts)[2])
    names = ("csp_attn_add", "csp_mlp_mm1", "csp_mlp_mm2")
    algs = (attn_alg_bytes(H, NSEQ, ATTN_COUNT), mm1_alg_bytes(M, MLP_COUNT), mm2_alg_bytes(M, MLP_COUNT))
    flops = (4.0 * QG * ATTN_COUNT * D * H * G, 2.0 * M * MLP_COUNT * MLP_K, 2.0 * M * MLP_COUNT * MLP_K)
    kern = {}
    for nm, ms, mc, ab, fl in zip(names, t, cold, algs, flops):
        tr = traffic_db.get(nm)
        r = _roof(ab, ms, hbm_gbs, peak_src, tr)
        # the gathered weight rows of the MLP are L2 hits (measured DRAM traffic is a fraction of the algorithmic bytes):
        # the HBM figure is the SURVEY 8d fraction, the bound that actually binds is L2 -> SM ingress / the tensor pipe
        r["bound"] = "hbm" if nm == "csp_attn_add" else "l2-ingress (frac is the SURVEY 8d algorithmic-bytes / HBM-peak figure and exceeds 1: the gathered weights are L2-resident)"
        kern[nm] = {"us": round(ms * 1e3, 1), "us_cold_l2": round(mc * 1e3, 1), "roofline": r,
                    "tensor_tflops": round(fl / (ms * 1e-3) / 1e12, 1), "tensor_frac_of_burst_peak": round(fl / (ms * 1e-3) / 1e12 / tf_burst, 4)}
    F = torch.nn.functional
    w2 = w2t.t().contiguous()
    t_sdpa = _time(lambda: F.scaled_dot_product_attention(q, k, v), 20)
    x3 = x[None]
    t_mlp = _time(lambda: F.linear(F.gelu(F.linear(x3, w1, b1), approximate="tanh"), w2), 20)
    res = {"ms_per_block_step": round(ms_blk, 4), "dense_equiv_tflops": round((C2_DENSE_FLOPS_ATTN + C2_DENSE_FLOPS_MLP) / ms_blk / 1e9, 1),
           "kernels": kern, "timed_region_ms": round(ms_blk * iters, 1),
           "dense_gpu_baseline": {"sdpa_us": round(t_sdpa * 1e3, 1), "cublas_mlp_us": round(t_mlp * 1e3, 1),
                                  "attn_speedup_vs_sdpa": round(t_sdpa / t[0], 2), "mlp_speedup_vs_cublas": round(t_mlp / (t[1] + t[2]), 2)}}
    # the reference's own Triton kernels on this box ("beat the Triton kernel on the same box")
    try:
        mm2 = _load_triton_ref("csp_mlp_mm2")
        mm1 = _load_triton_ref("csp_mlp_mm1")
        if mm1 is not None and mm2 is not None:
            one = torch.ones(1, device=dev, dtype=torch.float32)
            pa2, pk2, oc2 = pa_T.clone(), torch.zeros_like(packed), out_cache.clone()
            mm1.csp_mlp_mm1(x, w1, b1, m_idx, m_cnt, pa2, pk2, one, one)              # autotune
            t_r1 = _time(lambda: mm1.csp_mlp_mm1(x, w1, b1, m_idx, m_cnt, pa2, pk2, one, one), 10)
            t_r2 = _time(lambda: mm2.csp_mlp_mm2(packed, w2t, m_idx, m_cnt, oc2, 148), 10)
            res["reference_gpu"] = {"triton_mm1_us": round(t_r1 * 1e3, 1), "triton_mm2_us": round(t_r2 * 1e3, 1),
                                    "mm1_speedup_vs_reference_triton": round(t_r1 / t[1], 2),
                                    "mm2_speedup_vs_reference_triton": round(t_r2 / t[2], 2),
                                    "what": "the reference's csp_mlp_mm1 (Triton, bf16 operands, unit scales) and csp_mlp_mm2 "
                                            "(Triton, its only mm2) from oracle/_ref/triton_ref, same inputs, same box"}
    except Exception as e:  # noqa: BLE001   (no Triton / no staged reference: the cuBLAS baseline stands)
        res["reference_gpu"] = {"unavailable": repr(e)[:160]}
    return res


# ---------------------------------------------------------------------------------- CPU reference arm
def _cpu_sample(heads, rows, gen):
    q = torch.randn(1, heads, rows, D, generator=gen)
    k, v = (torch.randn(1, heads, C3_N, D, generator=gen) for _ in range(2))
    return q, k, v


def cpu_reference(sample_seconds=12.0, heads=1, rows=4096):
    """The reference's dense PyTorch branch (F.scaled_dot_product_attention, modules/attn.py:194) on the host cores, on a
    bounded sample of the workload: `rows` of the 119 056 query rows of `heads` of the 24 heads, against ALL keys."""
    from oracle import chipmunk_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    q, k, v = _cpu_sample(heads, rows, torch.Generator().manual_seed(0))
    flops = 4.0 * rows * C3_N * D * heads
    oracle.sdpa(q, k, v)
    t0, reps = time.time(), 0
    while True:
        oracle.sdpa(q, k, v)
        reps += 1
        if time.time() - t0 > sample_seconds or reps >= 20:
            break
    dt = (time.time() - t0) / reps
    return {"value": round(flops / dt / 1e12, 4), "unit": "TFLOP/s-equiv", "cores": cores, "kind": "port",
            "sample": f"{rows} of 119056 query rows x {heads}/24 heads against all 119056 keys (dense SDPA, fp32), {reps} reps "
                      f"({dt:.2f} s each); the reference's pure-PyTorch dense branch (modules/attn.py:194)",
            "seconds_per_full_layer_extrapolated": round(dt * C3_DENSE_FLOPS / flops, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import chipmunk_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    heads, rows = 1, 4096
    q, k, v = _cpu_sample(heads, rows, torch.Generator().manual_seed(0))
    flops = 4.0 * rows * C3_N * D * heads
    for _ in range(max(1, min(args.warmup, 2))):
        oracle.sdpa(q, k, v)
    t0 = time.time()
    for _ in range(args.steps):
        oracle.sdpa(q, k, v)
    dt = (time.time() - t0) / args.steps
    val = round(flops / dt / 1e12, 4)
    sample = (f"each step = {rows} of 119056 query rows x {heads}/24 heads against all keys (1/{int(C3_DENSE_FLOPS / flops)} of the layer), "
              "dense SDPA in fp32 on all host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC,
        "value": val, "unit": "TFLOP/s-equiv", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "hunyuan_720p_attention_layer", "seq": C3_N, "heads": H, "head_dim": D,
                   "note": "reference's dense PyTorch branch on CPU; bounded sample per step"},
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
    ap.add_argument("--no-extras", action="store_true", help="skip the dense GPU baselines, the full step, the FLUX block and the CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
