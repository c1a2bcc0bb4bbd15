#!/usr/bin/env python
"""Instruction-level view of one kernel launch from an .ncu-rep: executed warp instructions, stall samples and
active lanes per chunk of SASS, with the memory / call instructions of each chunk as landmarks (development aid).
usage: tools/ncu_chunks.py report.ncu-rep kernel-regex launch-skip [chunk]"""
import csv
import subprocess
import sys
from collections import Counter

rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
print(rows[0][:2])
hdr = rows[1]
ie, si, te, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[ie] == "Instructions Executed":
        break
    data.append(r)
tot_i = sum(float(r[ie]) for r in data)
tot_s = sum(float(r[si]) for r in data)
tot_t = sum(float(r[te]) for r in data)
print(f"{len(data)} SASS instructions, {tot_i:.4g} warp instructions executed, {tot_t / tot_i:.2f} lanes / instruction, {tot_s:.0f} samples")
for c in range(0, len(data), chunk):
    seg = data[c:c + chunk]
    i = sum(float(r[ie]) for r in seg)
    s = sum(float(r[si]) for r in seg)
    t = sum(float(r[te]) for r in seg)
    ops = [(r[src].split()[1] if r[src].strip().startswith("@") else r[src].split()[0]) for r in seg]
    keys = [m for m in ops if m.startswith(("LDG", "STG", "LDS", "STS", "LDL", "STL", "MUFU", "ATOM", "DADD", "DMUL", "VOTE", "SHFL", "CALL", "RET", "EXIT"))]
    kc = Counter(k.split(".")[0] + (".128" if ".128" in k else "") for k in keys)
    print(f"{c:5d} inst {100 * i / tot_i:5.1f}%  smp {100 * s / tot_s:5.1f}%  lanes {t / max(i, 1):5.1f}  {dict(kc)}")
