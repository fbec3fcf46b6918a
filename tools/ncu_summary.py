#!/usr/bin/env python
"""Summarise ncu output for profiles/ (runs here, no GPU needed).
  python tools/ncu_summary.py rep  <file.ncu-rep> [...]      key metrics per captured kernel launch (markdown table)
  python tools/ncu_summary.py list <launches.csv>             per-kernel totals / shares of an ncu launch list
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("dram__bytes_read.sum.per_second", "dram read/s"),
    ("dram__bytes_write.sum.per_second", "dram write/s"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def short(name):
    name = name.replace("void glu_b200::<unnamed>::", "").replace("glu_b200::<unnamed>::", "")
    return name.split("(")[0][:90]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"### {path}\n")
    for r in rows[2:]:
        print(f"**{short(r[col['Kernel Name']])}**  (launch id {r[col['ID']]})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m, label in METRICS:
            if m in col:
                print(f"| {label} (`{m}`) | {r[col[m]]} | {units[col[m]]} |")
        print()


def launch_list(path):
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    total = sum(a[1] for a in agg.values())
    print(f"### {path}: {sum(a[0] for a in agg.values())} launches, {total:.3f} ms total (ncu: cold-cache, serialised)\n")
    print("| kernel | launches | total ms | ms/launch | share |\n|---|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {ms:.3f} | {ms / n:.4f} | {100 * ms / total:.1f}% |")
    print()


if __name__ == "__main__":
    mode, files = sys.argv[1], sys.argv[2:]
    for f in files:
        (rep if mode == "rep" else launch_list)(f)
