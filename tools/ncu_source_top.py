#!/usr/bin/env python
"""Top stalled SASS lines of one kernel from an .ncu-rep (source page): python tools/ncu_source_top.py rep regex [n]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:1"],
                     capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines()) if r]
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr, data = rows[hi], rows[hi + 1:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0
data = [r for r in data if len(r) > isamp and r[0].startswith("0x")]
tot = sum(num(r[isamp]) for r in data)
print(rows[0][1][:120] if len(rows[0]) > 1 else "", "\ntotal samples", tot, "instructions", len(data))
top = sorted([(num(r[isamp]), i, r[ia].strip(), r[iex]) for i, r in enumerate(data)], reverse=True)[:n]
for s, i, src, ex in top:
    print(f"{s:7d} {s / max(tot,1) * 100:5.1f}%  idx{i:5d} exec={ex:>10s}  {src[:100]}")
