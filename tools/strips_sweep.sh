#!/bin/bash
# step time of a workload for pinned numbers of strips per member (development aid):
#   tools/strips_sweep.sh <workload> <members> P1 P2 ...
wl=$1; m=$2; shift 2
mkdir -p gpurun_out
for P in "$@"; do
  python bench.py --workload $wl --members $m --strips $P --particles 0 --steps 200 --warmup 20 \
      --no-cpu-baseline --no-single-run --no-scan --e2e-steps 3 > gpurun_out/strips_$P.json 2> gpurun_out/strips_$P.err
  python - $P gpurun_out/strips_$P.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r = l["roofline"]
    print("  strips %3s: %.2f us/step  frac %.4f  %.1f G" % (sys.argv[1], r["us_per_launch"], r["frac"], l["value"] / 1e9))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
