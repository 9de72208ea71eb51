#!/usr/bin/env python
"""one line per bench.py JSON file: step kernel, step, value, e2e (development aid)"""
import json
import sys

for f in sys.argv[1:]:
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        r = l["roofline"]
        print("%s: kernel %.2f us (%.4f)  step %.2f us  value %.1f G  e2e %.1f G" % (
            f, r["us_per_launch"], r["frac"], l["ms_per_step"] * 1e3, l["value"] / 1e9, l["e2e"]["value"] / 1e9))
    except Exception as e:
        print(f, "FAILED", e)
