"""Turn the scratch outputs of tools/r1_profiles.sh (gpurun_out/p_*) into the tracked evidence under profiles/:
ncu summaries per kernel, the launch-share list, the measured DRAM traffic per launch (profiles/traffic.json,
read by bench.py for roofline.traffic) and the bench lines.   usage: python tools/make_profiles.py [round_tag]"""
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
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def summary(rep, out):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(out, "w") as f:
        f.write(txt)
    return txt


def traffic_from(txt):
    res, name = {}, None
    for line in txt.splitlines():
        if line.startswith("== "):
            name = line[3:]
        elif "dram traffic (read+write)" in line and name:
            mb = float(line.split()[-2])
            key = "csp_attn_add" if "attn_kernel" in name else ("csp_mlp_mm2" if "mlp_kernel<1>" in name or "(bool)1" in name else "csp_mlp_mm1")
            res[key] = mb * 1e6
    return res


def main():
    os.makedirs(P, exist_ok=True)
    t2 = summary(os.path.join(G, "p_c2.ncu-rep"), os.path.join(P, f"{tag}_ncu_c2_flux_block.txt"))
    summary(os.path.join(G, "p_c3.ncu-rep"), os.path.join(P, f"{tag}_ncu_c3_hunyuan_attn.txt"))
    tr = traffic_from(t2)
    tr["source"] = f"profiles/{tag}_ncu_c2_flux_block.txt (ncu --set full, one launch each, bench.py workload)"
    with open(os.path.join(P, "traffic.json"), "w") as f:
        json.dump(tr, f, indent=1)
    # launch shares
    lp = os.path.join(G, "p_launches.csv")
    if os.path.exists(lp):
        shutil.copy(lp, os.path.join(P, f"{tag}_launches_bench.csv"))
        with open(lp) as f:
            rows = list(csv.reader(l for l in f if l.startswith('"')))
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.defaultdict(list)
        for r in rows[1:]:
            v = float(r[vi].replace(",", ""))
            v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
            agg[r[ki]].append(v)
        tot = sum(sum(v) / len(v) for v in agg.values())
        bench = json.load(open(os.path.join(G, "p_bench_n1.json")))
        with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as f:
            f.write("share of the step's GPU time per kernel (ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_kernel|mlp_kernel,\n"
                    "bench.py --steps 4 --warmup 3 --no-extras; cold-cache serialised launches: compare SHARES with bench.py's CUDA-event times, not absolutes)\n")
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
                a = sum(v) / len(v)
                f.write(f"{k[:70]:70s} n={len(v):4d} avg={a:8.1f} us share={100 * a / tot:5.1f}%\n")
            ks = bench["kernels"]
            tb = sum(k["us"] for k in ks.values())
            f.write("bench.py CUDA-event shares: " + ", ".join(f"{n} {100 * k['us'] / tb:.1f}%" for n, k in ks.items()) + "\n")
    for src, dst in (("p_bench_n1.json", f"{tag}_bench_n1.json"), ("p_bench_ref.json", f"{tag}_bench_reference_arm.json"),
                     ("p_sweep.jsonl", f"{tag}_sweep.jsonl"), ("p_sample_schedule.json", f"{tag}_sample_schedule.json")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))


if __name__ == "__main__":
    main()
