#!/usr/bin/env python3
"""Print value / roofline fraction / e2e of bench.py JSON lines: tools/bench_line.py file.json [more.json ...]"""
import json, sys
for path in sys.argv[1:]:
    with open(path) as fh:
        lines = [ln for ln in fh.read().strip().splitlines() if ln.startswith("{")]
    for ln in lines:
        d = json.loads(ln)
        print(path, "n_gpus", d.get("n_gpus"), "value", round(d["value"]), "frac", round(d["roofline"]["frac"], 3) if "roofline" in d else None,
              "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None)
