#!/usr/bin/env python
"""Per-kernel SASS opcode evidence (profiles/rNN_sass_opcodes.md): which kernels of libglu_b200.so really contain
TMA bulk copies (UBLKCP), L2 bulk prefetches (UBLKPF), mbarrier waits (SYNCS), warp reductions (REDUX), ballots (VOTE),
shared atomics (ATOMS) ... and that none contains tensor-core instructions (nothing here is a contraction).
Runs on the CPU box: cuobjdump disassembles the sm_100a cubins embedded in the in-tree .so.

  python tools/sass_opcodes.py [path/to/lib.so] > profiles/r02_sass_opcodes.md
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gl-radix-sort_b200", "libglu_b200.so")
OPS = ["UBLKCP", "UBLKPF", "SYNCS", "REDUX", "VOTE", "MATCH", "ATOMS", "ATOMG", "SHFL", "LDG", "STG", "LDS", "STS",
       "BAR", "HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
kernels = []
cur = None
arch = set()
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = {"name": m.group(1), "ops": dict.fromkeys(OPS, 0), "n": 0}
        kernels.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["n"] += 1
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o):
                cur["ops"][o] += 1
                break
names = [k["name"] for k in kernels]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode counts per kernel of `{os.path.relpath(lib, ROOT)}` (static instruction counts; cuobjdump -sass)\n")
print(f"arch: {', '.join(sorted(arch))}.  UBLKCP = cp.async.bulk (TMA 1-D bulk copy), UBLKPF = cp.async.bulk.prefetch.L2, "
      f"SYNCS = mbarrier try_wait / arrive, REDUX = redux.sync, VOTE = vote.ballot, ATOMS = shared atomics.  "
      f"HMMA/IMMA/UTC*MMA (tensor cores) are expected to be 0 everywhere.\n")
shown = [o for o in OPS if any(k["ops"][o] for k in kernels)] + ["HMMA", "UTCHMMA"]
shown = list(dict.fromkeys(shown))
print("| kernel | SASS instr | " + " | ".join(shown) + " |")
print("|---|---|" + "---|" * len(shown))
def short(d):
    d = re.sub(r"glu_b200::\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    d = re.sub(r"\(.*$", "", d)
    return d
agg = {}
for k, d in zip(kernels, dem):
    agg[short(d)] = k
for name in sorted(agg):
    k = agg[name]
    print(f"| `{name}` | {k['n']} | " + " | ".join(str(k['ops'][o]) for o in shown) + " |")
tot = {o: sum(k["ops"][o] for k in kernels) for o in OPS}
print(f"\nTotals over {len(kernels)} kernels: " + ", ".join(f"{o} {tot[o]}" for o in OPS))
