"""Top stall-sample SASS lines of each kernel in an .ncu-rep (source page).  usage: ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)
for b in blocks:
    rows = b["rows"]
    tot = sum(int(r[2]) for r in rows)
    print(f"== {b['name'][:90]}  samples={tot}")
    idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][2]))[:topn]
    for i in sorted(idx):
        r = rows[i]
        print(f"  {i:5d} {int(r[2]):7d} ({100*int(r[2])/max(tot,1):4.1f}%) exec={r[5]:>9s}  {r[1].strip()[:80]}")
