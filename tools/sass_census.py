"""SASS census of libchipmunk_b200.so: per kernel, how often the Blackwell-native instructions occur (the PTX names never
appear in SASS: tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UTMAREDG/UBLKCP; cp.async = LDGSTS;
legacy tensor path = HMMA).   usage: python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chipmunk_b200", "libchipmunk_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UBLKRED", "LDGSTS", "MUFU.EX2", "MUFU.TANH",
        "REDG", "RED.", "ATOMS", "HMMA", "SYNCS", "MULTIMEM", "ELECT", "USETMAXREG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        mm = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if mm:
            op = mm.group(1)
            per[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k) or (k.endswith(".") and op.startswith(k[:-1] + ".")):
                    per[cur][k] += 1
    print("SASS census of chipmunk_b200/libchipmunk_b200.so (cuobjdump -sass; sm_100a).  Columns: instruction count per kernel.")
    tot = collections.Counter()
    for fn, c in per.items():
        cols = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        print(f"{fn[:90]:90s} total={c['_total']:6d}  {cols}")
        tot.update(c)
    print("\nwhole library: " + " ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
    print("HMMA (legacy mma.sync path) = %d: every tensor-core instruction is a tcgen05 UTC*MMA." % tot["HMMA"])


if __name__ == "__main__":
    main()
