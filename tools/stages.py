#!/usr/bin/env python
"""Print the per-stage numbers of a bench.py JSON line (stdin)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(d["config"]["workload"], "ms/step %.3f" % d["ms_per_step"], "fps %.1f" % d.get("fps", 0), "value %.3e" % d["value"], "e2e ms %.3f" % d["e2e"].get("ms_per_step", 0))
    for k, v in d["roofline"]["stages"].items():
        print("   %-22s %8.3f ms  %8.1f GB/s  frac %.3f" % (k, v["ms"], v["GBps"], v["frac"]))
