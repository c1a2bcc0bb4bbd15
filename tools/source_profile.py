#!/usr/bin/env python
"""Coarse per-region profile of one kernel from `ncu -i rep --page source --csv --print-source sass`: share of executed
warp instructions, threads per instruction and stall samples per block of consecutive SASS instructions (development aid).
usage: tools/source_profile.py page.csv [instructions per block, default 24] [--top N]"""
import csv
import sys

path = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 24
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if r and r[0].startswith("0x") and len(r) == len(hdr)]
ia, it, iss, isrc = (hdr.index(n) for n in ("Instructions Executed", "Thread Instructions Executed", "Warp Stall Sampling (All Samples)", "Source"))
base = int(data[0][0], 16)
tot = sum(int(r[ia]) for r in data)
tott = sum(int(r[it]) for r in data)
tots = sum(int(r[iss]) for r in data)
print(f"{len(data)} instructions, {tot} warp instructions executed, {tott / tot:.2f} threads/instruction, {tots} stall samples")
for i in range(0, len(data), step):
    ch = data[i:i + step]
    a = sum(int(r[ia]) for r in ch)
    t = sum(int(r[it]) for r in ch)
    s = sum(int(r[iss]) for r in ch)
    print(f"{int(ch[0][0], 16) - base:05x} inst {100 * a / tot:5.2f}% thr/inst {t / max(a, 1):5.1f} stall {100 * s / tots:5.2f}%  {ch[0][isrc].strip()[:60]}")
if top:
    print("-- top stall instructions")
    for r in sorted(data, key=lambda r: -int(r[iss]))[:top]:
        print(f"{int(r[0], 16) - base:05x} stall {100 * int(r[iss]) / tots:5.2f}% exec {100 * int(r[ia]) / tot:5.2f}% thr {int(r[it]) / max(int(r[ia]), 1):5.1f}  {r[isrc].strip()[:70]}")
