#!/usr/bin/env python
"""Turns gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and *.ncu-rep captures into small text
summaries under profiles/ (tracked)."""
import csv
import re
import subprocess
import sys
from collections import OrderedDict


def launches(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    agg = OrderedDict()
    for name, us in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path}); {len(rows)} launches, {total/1e3:.2f} ms of kernel time\n")
        f.write("# per-launch times are cold-cache / serialised: compare SHARES, not absolutes\n\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:70]:70s} {n:8d} {us:12.1f} {us/n:10.1f} {us/total*100:6.1f}%\n")
    print(open(out).read())


METRICS = ("gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__bytes_read.sum,|dram__bytes_write.sum,|"
           "gpu__dram_throughput|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct|"
           "launch__registers_per_thread|launch__grid_size|launch__block_size|sm__throughput.avg.pct|"
           "l1tex__data_pipe_lsu_wavefronts_mem_shared|lts__t_bytes.sum |lts__t_sector_hit_rate|sm__cycles_elapsed.avg |"
           "smsp__inst_executed.sum |sm__pipe_xu|smsp__inst_executed_pipe_xu|launch__shared_mem_per_block|sm__cycles_active.avg ")


def rep(path, out):
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if l and not l.startswith("==")]
    rd = list(csv.reader(lines))
    if not rd:
        print("empty report", r.stderr[:500])
        return
    hdr = rd[0]
    pat = re.compile(METRICS.replace(" ", ""))
    cols = [i for i, h in enumerate(hdr) if pat.search(h) or h in ("Kernel Name", "ID")]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {path}\n")
        units = rd[1] if len(rd) > 1 else []
        for row in rd[2:]:
            f.write("\n")
            for i in cols:
                if i < len(row):
                    f.write(f"{hdr[i]} [{units[i] if i < len(units) else ''}] = {row[i][:120]}\n")
    print(open(out).read()[:6000])


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    (launches if kind == "launches" else rep)(src, dst)
