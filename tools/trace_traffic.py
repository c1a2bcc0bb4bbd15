#!/usr/bin/env python
"""Regenerates profiles/trace_traffic.json — DRAM bytes, L1/L2 behaviour and SIMT efficiency of the five timed traversal
launches of one bench.py step — from an `ncu --set full` capture of the bench command:

    ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "crt_timed_steps/" -k regex:trace_kernel \
        -o /tmp/bench_trace python bench.py --steps 1 --warmup 1 --no-render --no-cpu-baseline --no-e2e
    python tools/trace_traffic.py /tmp/bench_trace.ncu-rep

(bench.py wraps its timed steps in the NVTX range `crt_timed_steps`, so the capture holds exactly the five timed launches,
however many launches the set-up made before them; the launch durations are checked to be in the order of the sets' sizes.)

The json is stamped with a hash of the traversal sources' code, comments and blank lines left out (`source_sha`); bench.py only
quotes `roofline.traffic` from it when the hash matches the sources it is running, so the figure cannot silently go stale.
"""
import argparse
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETS = ["closest_lcg", "closest_camera", "closest_bounce", "occluded_lcg", "occluded_shadow"]
KERNEL_SOURCES = ["traverse.cu", "traverse.cuh", "launch.h", "crt_layout.h"]


def _code_only(text: str) -> str:
    """The source without comments and without blank lines / indentation: a comment edit does not make the capture stale."""
    import re
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    lines = (re.sub(r"//.*$", "", ln).strip() for ln in text.splitlines())  # no string literal of these files holds "//"
    return "\n".join(ln for ln in lines if ln)


def source_sha(read=None) -> str:
    """Hash of the code (not the comments) of the traversal sources. `read(name)` overrides where a file comes from
    (tools use it to hash the sources of another commit)."""
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        text = read(name) if read else open(os.path.join(ROOT, "crust-render_b200", "csrc", name), encoding="utf-8").read()
        h.update(_code_only(text).encode())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--last", type=int, default=5, help="use the last N trace_kernel launches of the report (one bench step)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "trace_traffic.json"))
    ap.add_argument("--raw-out", default=None, help="also keep the raw csv page here")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    if a.raw_out:
        open(a.raw_out, "w").write(txt)
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[0], [r for r in rows[2:] if len(r) == len(rows[0])]
    import re
    # the plain build of the kernel only: the counting build (last template argument 1) runs once after the timed region
    plain = re.compile(r"trace_kernel<[^>]*,\s*(?:\(bool\))?(?:0|false)>")
    data = [r for r in data if plain.search(r[hdr.index("Kernel Name")])][-a.last:]
    if len(data) != len(SETS):
        sys.exit(f"expected {len(SETS)} trace_kernel launches, found {len(data)}")

    def col(r, name):
        return float(r[hdr.index(name)].replace(",", ""))

    def in_bytes(r, name):
        unit = rows[1][hdr.index(name)].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        return col(r, name) * scale

    per = {}
    for name, r in zip(SETS, data):
        ms = col(r, "gpu__time_duration.sum")
        unit = rows[1][hdr.index("gpu__time_duration.sum")]
        ms *= {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit, 1.0)
        rd, wr = in_bytes(r, "dram__bytes_read.sum"), in_bytes(r, "dram__bytes_write.sum")
        l2 = in_bytes(r, "lts__t_bytes.sum") if "lts__t_bytes.sum" in hdr else None
        per[name] = {"kernel": r[hdr.index("Kernel Name")], "dram_read": rd, "dram_write": wr, "traffic": rd + wr, "kernel_ms": ms,
                     "threads_per_inst": col(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                     "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in hdr else None,
                     "l1_hit_pct": col(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": col(r, "lts__t_sector_hit_rate.pct"),
                     "warp_inst": col(r, "smsp__inst_executed.sum"), "registers": int(col(r, "launch__registers_per_thread")),
                     "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                     "dram_gb_per_s": (rd + wr) / ms / 1e6, "l2_gb_per_s": None if l2 is None else l2 / ms / 1e6}
    closest = [n for n in SETS if n.startswith("closest")]
    # the five launches must be the five sets in order: closest-hit kernels first (their name has <false / (bool)0 first), then any-hit
    kinds = [bool(re.search(r"trace_kernel<(?:\(bool\))?(?:1|true)\b", per[n]["kernel"])) for n in SETS]
    if kinds != [False, False, False, True, True]:
        sys.exit(f"the captured launches are not one bench step in order (any-hit flags {kinds}): check the capture window")
    out = {"source": f"{os.path.relpath(a.report, ROOT)} (ncu --set full --clock-control none, the five timed launches of one bench step); regenerate with tools/trace_traffic.py",
           "source_sha": source_sha(), "kernel_sources": KERNEL_SOURCES, "per_launch": per,
           "mean_traffic_per_closest_launch": sum(per[n]["traffic"] for n in closest) / len(closest)}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps({k: (round(v["traffic"] / 1e9, 3), round(v["kernel_ms"], 3), v["threads_per_inst"]) for k, v in per.items()}))


if __name__ == "__main__":
    main()
