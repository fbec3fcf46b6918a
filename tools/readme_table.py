#!/usr/bin/env python
"""Regenerates the reference's README benchmark table (README.md:99-134, the 34 rows of BASELINE.md) on this GPU:
runs `cpp_tests/glu_test [benchmark]` (same case names, same sizes, same "<Name>; Num elements: N, Elapsed: T" lines,
zero-filled input as the reference does, plus uniform-random keys for the sort) and writes the rows side by side with
the reference's published RTX 2060 SUPER figures.

  python tools/readme_table.py > profiles/r02_glu_test_benchmark.md        (on a GPU box)
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "cpp_tests", "glu_test")


def published():
    rows = {}
    for line in open(os.path.join(ROOT, "BASELINE.md")):
        m = re.match(r"\|\s*\**(Reduce\(Uint,Sum\)|BlellochScan\(Uint\)|RadixSort u32 key\+val)\**\s*\|\s*\**([\d,]+)\**\s*\|"
                     r"\s*\**([\d.]+) (ms|s)\**\s*\|.*`(README\.md:\d+)`", line)
        if m:
            name = {"Reduce(Uint,Sum)": "Reduce", "BlellochScan(Uint)": "BlellochScan",
                    "RadixSort u32 key+val": "Radix sort"}[m.group(1)]
            ms = float(m.group(3)) * (1000.0 if m.group(4) == "s" else 1.0)
            rows[(name, int(m.group(2).replace(",", "")))] = (ms, m.group(5))
    return rows


def to_ms(text):
    m = re.match(r"([\d.]+)\s*(ns|us|µs|ms|s)", text.strip())
    v, u = float(m.group(1)), m.group(2)
    return v * {"ns": 1e-6, "us": 1e-3, "µs": 1e-3, "ms": 1.0, "s": 1e3}[u]


def main():
    r = subprocess.run([EXE, "[benchmark]"], capture_output=True, text=True, timeout=1200)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-2000:] + r.stderr[-2000:])
        raise SystemExit(r.returncode)
    pub = published()
    rows = []
    for line in r.stdout.splitlines():
        m = re.match(r"(Reduce|BlellochScan|Radix sort); Num elements: (\d+), Elapsed: ([^()]+?)(?: \(uniform keys: ([^)]+)\))?$", line.strip())
        if m:
            rows.append((m.group(1), int(m.group(2)), to_ms(m.group(3)), to_ms(m.group(4)) if m.group(4) else None))
    name = subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip().splitlines()
    print(f"# `./glu_test [benchmark]` on {name[0] if name else 'this GPU'} next to the reference's published table\n")
    print("Reference column: `/root/reference/README.md:99-134` (RTX 2060 SUPER, GL_TIME_ELAPSED, one run, zero-filled "
          "input). B200 columns: `cpp_tests/glu_test [benchmark]` (CUDA events around the class call operator, one "
          "warm-up, zero-filled input like the reference; for the sort also uniform-random mt19937 keys).\n")
    print("| Primitive | N | reference (2060S) | B200 zero-filled | B200 uniform keys | speed-up (zero-filled) | B200 throughput |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    per = {"Reduce": (4, "GB/s"), "BlellochScan": (8, "GB/s"), "Radix sort": (1, "Mpairs/s")}
    for prim, n, ms, ms_u in rows:
        ref = pub.get((prim, n))
        bpe, unit = per[prim]
        best = ms_u if ms_u is not None else ms
        thr = bpe * n / best / (1e6 if unit == "GB/s" else 1e3)
        print(f"| {prim} | {n:,} | {('%.3f ms (`%s`)' % ref) if ref else '—'} | {ms:.4f} ms | "
              f"{('%.4f ms' % ms_u) if ms_u is not None else '—'} | {('%.0fx' % (ref[0] / ms)) if ref else '—'} | {thr:,.1f} {unit} |")
    missing = [k for k in pub if k not in {(p, n) for p, n, _, _ in rows}]
    print(f"\n{len(rows)} rows measured, {len(pub)} rows published, missing: {missing if missing else 'none'}.")
    print("\nRaw output:\n\n```\n" + r.stdout.strip() + "\n```")


if __name__ == "__main__":
    main()
