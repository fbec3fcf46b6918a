#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_source_summary.py [kernel-substr]"""
import csv
import sys
from collections import defaultdict

want = sys.argv[1] if len(sys.argv) > 1 else ""
rows = list(csv.reader(sys.stdin))
# The export is a sequence of blocks: "File Path", "Function Name", header row, then data rows.
blocks = []
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        path = r[1]
        fn = rows[i + 1][1]
        header = rows[i + 2]
        j = i + 3
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            if rows[j]:
                data.append(rows[j])
            j += 1
        blocks.append((path, fn, header, data))
        i = j
    else:
        i += 1
seen_fn = set()
for path, fn, header, data in blocks:
    if want not in fn:
        continue
    key = (fn, path)
    if key in seen_fn:
        continue
    seen_fn.add(key)
    try:
        c_line, c_src = header.index("Line No"), header.index("Source")
        c_inst = header.index("Instructions Executed")
        c_samp = header.index("# Samples")
    except ValueError:
        continue
    c_wave = header.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in header else None
    tot_inst = 0
    tot_samp = 0
    per = []
    for d in data:
        try:
            inst = int(d[c_inst] or 0)
            samp = int(d[c_samp] or 0)
        except (ValueError, IndexError):
            continue
        wave = 0
        if c_wave is not None:
            try:
                wave = int(d[c_wave] or 0)
            except ValueError:
                wave = 0
        if inst or samp:
            per.append((d[c_line], d[c_src].strip()[:100], inst, samp, wave))
        tot_inst += inst
        tot_samp += samp
    if not tot_inst:
        continue
    print(f"== {path}  [{fn[:80]}]  inst={tot_inst} samples={tot_samp}")
    for line, src, inst, samp, wave in per:
        if inst * 100 >= tot_inst or samp * 100 >= max(1, tot_samp):
            print(f"  {line:>5s} inst {100.0 * inst / tot_inst:5.1f}%  stall-samples {100.0 * samp / max(1, tot_samp):5.1f}%  smem-wavefronts {wave:>10d}  | {src}")
