#!/usr/bin/env python3
"""Print value / roofline fraction of a bench.py JSON line read from stdin, prefixed by argv[1:]."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(*sys.argv[1:], round(d["value"]), round(d["roofline"]["frac"], 3))
