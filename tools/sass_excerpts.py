#!/usr/bin/env python
"""profiles/r02_sass_excerpts.txt: per kernel of libvsb200.so, the count of the SASS mnemonics that prove the Blackwell-native
path (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit, HMMA = mma.sync,
peer LD/ST in the DSP-fused kernels) plus the first line of each kind.  python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "videosys_b200", "csrc", "libvsb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "MUFU.EX2", "SYNCS", "LDG.E.128", "STG.E.128", "ST.E.128",
      "LD.E.128", "MEMBAR", "UCGABAR", "ELECT")
cur, stats, first = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        stats[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    stats[cur]["_total"] += 1
    for k in MN:
        if op.startswith(k):
            stats[cur][k] += 1
            first.setdefault((cur, k), line.strip()[:150])
print(f"# SASS evidence, libvsb200.so ({os.path.getsize(so)} bytes), cuobjdump -sass; {len(stats)} kernels (bf16 + fp16 twins)")
print("# PTX -> SASS: tcgen05.mma -> UTCHMMA(.2CTA), tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, cp.async.bulk.tensor ->")
print("# UTMALDG/UTMASTG, mma.sync -> HMMA, mbarrier -> SYNCS, elect.sync -> ELECT\n")
for k, c in stats.items():
    if c["_total"] < 40:
        continue
    keys = " ".join(f"{m}={c[m]}" for m in MN if c[m])
    print(f"{k}\n    instructions={c['_total']} {keys}")
    for m in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA", "ST.E.128", "LD.E.128", "LDG.E.128", "STG.E.128"):
        if (k, m) in first and (m in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA") or "dsp" in k.lower() or "true" in k):
            print(f"      {first[(k, m)]}")
