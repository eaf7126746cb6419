#!/usr/bin/env python3
"""Executed-instruction accounting from `ncu --page source --csv` output.
usage: ncu_hot.py src.csv   -> executed warp instructions by opcode, and the stall-sample leaders"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ops = collections.Counter()
tot = 0
hot = []
kernels_seen = 0
for r in rows[1:]:
    if r and r[0] == "Address":
        kernels_seen += 1  # a report with several launches repeats the header: keep the first launch only
        continue
    if kernels_seen > 1 or len(r) <= iex:
        continue
    src = r[isrc].strip()
    m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", src)
    n = int(float(r[iex] or 0))
    tot += n
    if m:
        ops[m.group(1)] += n
    hot.append((int(float(r[ismp] or 0)), n, src))
print(f"total executed warp instructions: {tot}")
base = max(n for _, n, _ in hot)
print("by opcode (count, per hottest-line execution):")
for k, v in ops.most_common(32):
    print(f"  {k:12s} {v:12d}  {v / base:7.1f}")
print("top stall-sample lines:")
for s, n, src in sorted(hot, reverse=True)[:25]:
    print(f"  {s:6d} samples  {n:10d} exec  {src[:90]}")
