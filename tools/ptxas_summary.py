#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` output: one line per kernel (registers, spills, shared memory)."""
import re, subprocess, sys

log = open(sys.argv[1]).read().splitlines()
cur = None
rows = []
for ln in log:
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"Used (\d+) registers", ln)
    if m and cur:
        sm = re.search(r"(\d+) bytes smem", ln)
        rows.append([cur, int(m.group(1)), int(sm.group(1)) if sm else 0, 0])
        cur = None
    m = re.search(r"(\d+) bytes spill stores", ln)
    if m and rows is not None and cur:
        pass
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for (mangled, regs, smem, _), name in sorted(zip(rows, names), key=lambda t: t[1]):
    print(f"{regs:4d} regs {smem:6d} B smem  {name.split('(')[0]}")
