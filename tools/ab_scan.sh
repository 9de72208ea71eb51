#!/bin/bash
# A/B timing of library variants (inovesa_b200/lib/variants/) on the 256 x 256^2 scan (BASELINE configs[3]) and on
# single 512^2 .. 2048^2 runs with one launch per step (development aid, see tools/ab.sh)
mkdir -p gpurun_out
for so in inovesa_b200/lib/libivs.so inovesa_b200/lib/variants/libivs_*.so; do
  name=$(basename $so .so)
  IVS_LIB=$PWD/$so python bench.py --steps 20 --warmup 5 --repeats 1 --no-cpu-baseline --e2e-steps 3 \
      > gpurun_out/abscan_$name.json 2> gpurun_out/abscan_$name.err
  python - "$name" gpurun_out/abscan_$name.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    s = l["scan_256x256"]
    sr = l.get("single_run", {})
    print("%-16s scan %.2f us/step  %.1f G  fallback %d | kernel(26x2048^2) %.2f us | per-step-launch single runs: %s" % (
        sys.argv[1], s["us_per_step"], s["cell_updates_per_s"] / 1e9, s["fallback_steps"], l["roofline"]["us_per_launch"],
        "  ".join("%s: %.2f" % (k, v["per_step_launch"]["us_per_step"]) for k, v in sr.items() if isinstance(v, dict) and "per_step_launch" in v)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
