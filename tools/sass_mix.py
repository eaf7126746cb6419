#!/usr/bin/env python3
"""Instruction mix per kernel from `cuobjdump -sass <lib>` output.  usage: sass_mix.py all.sass <substring> [top]"""
import collections, re, sys

text = open(sys.argv[1]).read()
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
for blk in text.split("Function : ")[1:]:
    name = blk.split("\n", 1)[0].strip()
    if want not in name:
        continue
    ops = collections.Counter()
    n = 0
    for ln in blk.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            ops[m.group(1)] += 1
            n += 1
    print(f"== {name}: {n} instructions")
    print("   " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(top)))
