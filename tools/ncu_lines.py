#!/usr/bin/env python
"""Per-CUDA-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [min-percent]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))


def num(x):
    try:
        return int(x)
    except (ValueError, TypeError):
        return 0


i = 0
done = set()
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        path, fn, h = r[1], rows[i + 1][1], rows[i + 2]
        j = i + 3
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            if rows[j]:
                data.append(rows[j])
            j += 1
        i = j
        if want not in fn or (fn, path) in done or "Instructions Executed" not in h:
            continue
        done.add((fn, path))
        cl, ci, cp = h.index("Line No"), h.index("Instructions Executed"), h.index("# Samples")
        cs = h.index("Source")
        stall = [(k, n) for k, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        lines = [d for d in data if d[cl].strip()]
        tot = sum(num(d[cp]) for d in lines)
        toti = sum(num(d[ci]) for d in lines)
        if not toti:
            continue
        print(f"== {path} [{fn[:70]}] samples={tot} warp-inst={toti}")
        for d in lines:
            sp, ins = num(d[cp]), num(d[ci])
            if sp * 100 >= minpct * max(tot, 1) or ins * 100 >= minpct * toti:
                st = sorted(((num(d[k]), n[6:]) for k, n in stall), reverse=True)[:2]
                print(f"{d[cl]:>5s} samp {100 * sp / max(tot, 1):5.1f}% inst {100 * ins / toti:5.1f}% "
                      f"{st[0][1]:>13s}:{st[0][0]:<6d} {st[1][1]:>13s}:{st[1][0]:<6d}| {d[cs].strip()[:80]}")
    else:
        i += 1
