#!/usr/bin/env python
"""SASS digest of libafb200.so (sm_100a): per kernel the register count and the memory-instruction mix that the design claims
(LDG.E.*.256 / .128 gathers, REDG F32x4 vector reductions, shared-memory atomics, TMA / tcgen05 - none expected).
    python profiles/sass_digest.py > profiles/r2_sass_digest.txt        (cuobjdump only, no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "acquisition_focus_b200", "lib", "libafb200.so")
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
regs = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and name:
        regs[name] = (int(m.group(1)), int(m.group(2)))
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = set(re.findall(r"arch = (sm_\w+)", sass))
counts = collections.defaultdict(collections.Counter)
cur = None
PAT = [("LDG.256", r"\bLDG\.E[\w.]*\.256\b"), ("LDG.128", r"\bLDG\.E[\w.]*\.128\b"), ("LDG.other", r"\bLDG\.E(?![\w.]*\.(128|256))"),
       ("STG.128", r"\bSTG\.E[\w.]*\.128\b"), ("STG.other", r"\bSTG\.E(?![\w.]*\.128)"), ("REDG.F32x4", r"\bREDG\.E\.ADD\.F32x4"),
       ("REDG.other", r"\bREDG\.E\.ADD(?!\.F32x4)"), ("RED/ATOMG.f64", r"\b(RED|ATOMG)\.E\.ADD\.F64"), ("ATOMS", r"\bATOMS\."),
       ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("SHFL", r"\bSHFL\."), ("TMA(UBLKCP/UTMA)", r"\b(UBLKCP|UTMALDG|UTMASTG|UTMAREDG)"),
       ("tcgen05(UTC*)", r"\bUTC\w+"), ("BAR", r"\bBAR\.SYNC")]
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    for key, pat in PAT:
        if re.search(pat, line):
            counts[cur][key] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
names = dict(zip(list(counts), dem)) if len(dem) == len(counts) else {k: k for k in counts}
print(f"# {os.path.relpath(LIB, ROOT)}: SASS architectures {sorted(arch)}; {len(counts)} kernels")
print("# columns: registers, static smem bytes, then instruction counts (static occurrences in the SASS)")
tot = collections.Counter()
for k in sorted(counts, key=lambda k: names[k]):
    short = re.sub(r"\(.*", "", names[k]).replace("void ", "").replace("afb::", "")
    r, sm = regs.get(k, (None, None))
    print(f"{short:70s} regs={r} smem={sm} " + " ".join(f"{a}={b}" for a, b in counts[k].items()))
    tot.update(counts[k])
print("# totals:", dict(tot))
print("# TMA instructions:", tot.get("TMA(UBLKCP/UTMA)", 0), " tcgen05 instructions:", tot.get("tcgen05(UTC*)", 0),
      "(the path is gather / scatter / stream: see DESIGN.md section 5 for why neither is used)")
