#!/bin/bash
# step-kernel time per member for several ensemble sizes per GPU (development aid)
mkdir -p gpurun_out
for m in "$@"; do
  python bench.py --members $m --steps 200 --warmup 20 --no-cpu-baseline --no-single-run --e2e-steps 3 > gpurun_out/sweep_m$m.json 2> gpurun_out/sweep_m$m.err
  python - $m gpurun_out/sweep_m$m.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r = l["roofline"]
    print("members %3s kernel %.2f us (%.3f us/member) frac %.4f  step %.2f us value %.1f G" % (
        sys.argv[1], r["us_per_launch"], r["us_per_launch"] / int(sys.argv[1]), r["frac"], l["ms_per_step"] * 1e3, l["value"] / 1e9))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
