#!/usr/bin/env python
"""Per-kernel totals of an `ncu --csv` launch list (gpu__time_duration plus whatever other metrics were collected):
time, launches, warp instructions, lanes per instruction, stall_no_inst per issue. Development aid."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ID, K, MN, MV = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
L = {}
for r in rows[start + 1:]:
    if len(r) > MV:
        L.setdefault(int(r[ID]), {"k": r[K]})[r[MN]] = float(r[MV].replace(",", ""))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = defaultdict(lambda: defaultdict(float))
for i in sorted(L):
    if i < skip:
        continue
    d = L[i]
    name = d["k"]
    for key in ("trace_kernel", "shade_kernel", "nee_resolve_kernel", "raygen_kernel", "accumulate_kernel", "adaptive_kernel", "film_kernel", "classify_count", "classify_scan", "classify_scatter"):
        if key in name:
            name = key + (name[name.index(key) + len(key):].split("(")[0] if "<" in name[name.index(key):name.index(key) + len(key) + 1] else "")
            break
    a = agg[name]
    ms = d.get("gpu__time_duration.sum", 0.0) / 1e6
    a["ms"] += ms
    a["n"] += 1
    inst = d.get("smsp__inst_executed.sum", 0.0)
    a["inst"] += inst
    a["tinst"] += inst * d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
    a["noinst_w"] += ms * d.get("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", 0.0)
    a["issue_w"] += ms * d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
tot = sum(a["ms"] for a in agg.values())
print(f"total {tot:.2f} ms over {int(sum(a['n'] for a in agg.values()))} launches")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{a['ms']:9.3f} ms {100 * a['ms'] / tot:5.1f}%  {int(a['n']):5d} launches  {a['inst'] / 1e6:10.1f} Minst  lanes {a['tinst'] / max(a['inst'], 1):5.1f}  "
          f"no_inst/issue {a['noinst_w'] / max(a['ms'], 1e-9):5.2f}  issue_active {a['issue_w'] / max(a['ms'], 1e-9):5.1f}%  {name[:70]}")
