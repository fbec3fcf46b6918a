#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: one line per kernel (registers, spills, smem). Reads stdin."""
import re
import subprocess
import sys

text = sys.stdin.read()
rows = []
cur = None
for line in text.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        cur = {"name": m.group(1), "spill": "0/0"}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["stack"] = m.group(1)
        cur["spill"] = f"{m.group(2)}/{m.group(3)}"
    m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
    if m:
        cur["regs"] = m.group(1)
        cur["smem"] = m.group(3) or "0"
names = [r["name"] for r in rows]
try:
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
except Exception:
    dem = names
for r, d in zip(rows, dem):
    d = re.sub(r"glu_b200::\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    d = re.sub(r"\((int|bool)\)", "", d)
    d = re.sub(r"<unnamed>::", "", d)
    d = re.sub(r"glu_b200::", "", d)
    d = re.sub(r">\(.*$", ">", d)
    d = re.sub(r"^(\w+)\(.*$", r"\1", d)
    print(f"{d:70s} regs={r.get('regs','?'):>3s} stack={r.get('stack','0'):>4s} spill(st/ld)={r['spill']:>7s} smem={r.get('smem','0')}")
