#!/usr/bin/env python
"""Registers / spills / stack / shared memory per kernel from an `nvcc -Xptxas -v` log (development aid; the summaries
committed under profiles/ come from this)."""
import re
import sys

for path in sys.argv[1:]:
    txt = open(path).read()
    print(f"# {path}")
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:.*?(\d+) bytes smem)?", txt):
        n = m.group(1)
        for key in ("trace_kernel", "shade_kernel", "nee_resolve_kernel"):
            if key in n:
                n = n[n.index(key):]
                n = re.sub(r"EEv.*$", "", n)
                break
        else:
            n = re.sub(r"^_ZN.*?\d+([a-z_]+kernel).*$", r"\1", n)
        print(f"  {n:44s} regs {m.group(5):>3s}  stack {m.group(2):>6s}  spill st/ld {m.group(3)}/{m.group(4)}  smem {m.group(6) or 0}")
