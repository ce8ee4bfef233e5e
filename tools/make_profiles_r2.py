"""Turn the scratch outputs of tools/r2_profiles.sh (gpurun_out/q_*) into the tracked round-2 evidence under profiles/:
ncu summaries per kernel (with the L2->SM ingress counters), the launch-share list of the bench command, the measured DRAM
traffic per launch (profiles/traffic.json, read by bench.py for roofline.traffic), bench lines, sweep, schedule, SASS census.
usage: python tools/make_profiles_r2.py"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
TAG = "r02"
EXTRA = ["l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
         "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum.pct_of_peak_sustained_elapsed",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
         "smsp__inst_executed.sum", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def ncu_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return [], [], []
    return rows[0], rows[1], rows[2:]


def summary(rep, out, header=""):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ncu_summary
    hdr, units, rows = ncu_rows(rep)
    if not rows:
        return {}
    name_col = hdr.index("Kernel Name")
    res = {}
    with open(out, "w") as f:
        if header:
            f.write(header + "\n")
        for r in rows:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            f.write(f"== {r[name_col][:110]}\n")
            for k in ncu_summary.KEYS + EXTRA:
                if k in d:
                    f.write(f"   {k:92s} {d[k]:>18s} {u[k]}\n")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            try:
                tot = sum(float(d[k]) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                f.write(f"   {'dram traffic (read+write)':92s} {tot / 1e6:18.3f} Mbyte\n")
                res[r[name_col]] = tot
            except Exception:
                pass
    return res


def main():
    os.makedirs(P, exist_ok=True)
    traffic = {}
    t = summary(os.path.join(G, "q_c3.ncu-rep"), os.path.join(P, f"{TAG}_ncu_c3_hunyuan_step.txt"),
                "ncu --set full --clock-control none, one launch each: the two kernels of the headline step (tools/prof_targets.py c3attn: 24 heads, N = 119056, count 8320)")
    for k, v in t.items():
        if "attn_kernel" in k:
            traffic["c3_csp_attn_add"] = v
        if "mask_to_indices" in k:
            traffic["c3_bitmask_to_indices"] = v
    t = summary(os.path.join(G, "q_c2.ncu-rep"), os.path.join(P, f"{TAG}_ncu_c2_flux_block.txt"),
                "ncu --set full --clock-control none, one launch each: the FLUX block kernels (tools/prof_targets.py c2)")
    for k, v in t.items():
        key = "csp_attn_add" if "attn_kernel" in k else ("csp_mlp_mm2" if ("mlp_kernel<1>" in k or "(bool)1" in k) else "csp_mlp_mm1")
        traffic[key] = v
    summary(os.path.join(G, "q_dense_cs.ncu-rep"), os.path.join(P, f"{TAG}_ncu_dense_colsum.txt"), "one-pass dense attention + column sums, H = 24, N = 16384")
    summary(os.path.join(G, "q_dense.ncu-rep"), os.path.join(P, f"{TAG}_ncu_dense.txt"), "one-pass dense attention, H = 24, N = 16384")
    summary(os.path.join(G, "q_select.ncu-rep"), os.path.join(P, f"{TAG}_ncu_select_columns.txt"), "select_columns, 4 heads x 621 groups x 119056 columns, k = 8320")
    if traffic:
        traffic["source"] = f"profiles/{TAG}_ncu_c3_hunyuan_step.txt and {TAG}_ncu_c2_flux_block.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
        with open(os.path.join(P, "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    lp = os.path.join(G, "q_launches.csv")
    if os.path.exists(lp):
        shutil.copy(lp, os.path.join(P, f"{TAG}_launches_bench.csv"))
        with open(lp) as f:
            rows = list(csv.reader(l for l in f if l.startswith('"')))
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.defaultdict(list)
        for r in rows[1:]:
            v = float(r[vi].replace(",", ""))
            v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
            agg[r[ki]].append(v)
        with open(os.path.join(P, f"{TAG}_launch_shares.txt"), "w") as f:
            f.write("every launch of `python bench.py --steps 4 --warmup 3 --no-extras` with its device time (ncu --metrics gpu__time_duration.sum\n"
                    "--clock-control none; cold-cache, serialised: compare SHARES with bench.py's CUDA-event times, not absolutes)\n")
            mine = {k: v for k, v in agg.items() if "cm::" in k or "attn_kernel" in k or "mask_to_indices" in k}
            tot = sum(sum(v) for v in mine.values())
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                share = f"share of our kernels' time {100 * sum(v) / tot:5.1f}%" if k in mine else "(input generation / copies, outside the timed region)"
                f.write(f"{k[:80]:80s} n={len(v):4d} avg={sum(v) / len(v):10.1f} us  {share}\n")
            try:
                bench = json.load(open(os.path.join(G, "q_bench_n1.json")))
                ks = bench["us_per_layer"]
                tb = sum(ks.values())
                f.write("bench.py CUDA-event shares of the step: " + ", ".join(f"{n} {100 * u / tb:.1f}%" for n, u in ks.items()) + "\n")
            except Exception:
                pass
    for src, dst in (("q_bench_n1.json", f"{TAG}_bench_n1.json"), ("q_bench_ref.json", f"{TAG}_bench_reference_arm.json"),
                     ("q_sweep.jsonl", f"{TAG}_sweep.jsonl"), ("q_sample_schedule.json", f"{TAG}_sample_schedule.json"),
                     ("q_index.json", f"{TAG}_index_kernels.json"), ("q_pytest.txt", f"{TAG}_pytest_gpu.txt"),
                     ("r2_bench_n2.json", f"{TAG}_bench_n2.json"), ("r2_bench_n4.json", f"{TAG}_bench_n4.json"),
                     ("r2_bench_n8.json", f"{TAG}_bench_n8.json"), ("r2_schedule_n8.json", f"{TAG}_sample_schedule_n8.json"),
                     ("r2_schedule_n2.json", f"{TAG}_sample_schedule_n2.json")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    with open(os.path.join(P, f"{TAG}_sass_census.txt"), "w") as f:
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_census.py")], capture_output=True, text=True).stdout)


if __name__ == "__main__":
    main()
