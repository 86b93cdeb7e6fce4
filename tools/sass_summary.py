#!/usr/bin/env python
"""tools/sass_summary.py — opcode counts per kernel family from `cuobjdump -sass scalaopt_b200/libscalaopt_cuda.so`:
what proves which hardware path a kernel takes (UBLKCP / UTMALDG = TMA, SYNCS = mbarrier, DMMA = FP64 tensor path,
LDGSTS = cp.async, DFMA, LDS.128, peer/system-scope stores of so_finish).  python tools/sass_summary.py > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "scalaopt_b200", "libscalaopt_cuda.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fam = collections.defaultdict(collections.Counter)
kcount = collections.Counter()
cur = None
WANT = ["UBLKCP", "UTMALDG", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "LDGSTS", "LDG.E.128", "LDG.E.EF.128", "LDG.E.64", "LDS.128", "LDS.64", "SHFL", "MUFU", "ST.E.64.STRONG.SYS", "LD.E.64.STRONG.SYS", "MEMBAR.SC.SYS", "BAR.SYNC", "ATOMG", "RED"]
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        mm = re.search(r"(k_[a-z_0-9]+)", d)
        cur = mm.group(1) if mm else d[:40]
        kcount[cur] += 1
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m:
        continue
    op = m.group(1)
    fam[cur]["_all"] += 1
    for w in WANT:
        if op.startswith(w):
            fam[cur][w] += 1
cols = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "DMMA", "DFMA", "LDS.128", "LDS.64", "LDG.E.EF.128", "SHFL", "ST.E.64.STRONG.SYS", "LD.E.64.STRONG.SYS"]
print("# SASS opcode summary of `scalaopt_b200/libscalaopt_cuda.so` (sm_100a), per kernel family, summed over its instantiations\n")
print("`cuobjdump -sass` of the shipped build; produced by `tools/sass_summary.py`.  UBLKCP = `cp.async.bulk` (TMA bulk copy), UTMALDG ="
      " `cp.async.bulk.tensor` (TMA tensor-map copy), SYNCS = mbarrier operations, LDGSTS = `cp.async`, DMMA = `mma.sync.m8n8k4.f64`;"
      " `ST/LD.E.64.STRONG.SYS` are the system-scope peer stores / flag loads of `so_finish` (the in-kernel cross-GPU sum)."
      " No `UTC*MMA` / `LDTM`: FP64 has no tcgen05 kind.\n")
print("| kernel | instantiations | instructions | " + " | ".join(cols) + " |")
print("|---|---|---|" + "---|" * len(cols))
for k in sorted(fam, key=lambda k: -fam[k]["_all"]):
    print(f"| `{k}` | {kcount[k]} | {fam[k]['_all']} | " + " | ".join(str(fam[k][c]) for c in cols) + " |")
