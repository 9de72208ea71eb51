#!/bin/bash
# A/B timing of builds of libivs.so kept under inovesa_b200/lib/variants/ (development aid):
#   tools/ab.sh [steps] [extra bench.py arguments]  -> one line per variant: step kernel alone (us), step with tracking (ms)
steps=${1:-300}; shift
mkdir -p gpurun_out
for so in inovesa_b200/lib/libivs.so inovesa_b200/lib/variants/libivs_*.so; do
  name=$(basename $so .so)
  IVS_LIB=$PWD/$so python bench.py --steps $steps --warmup 30 --repeats 3 --no-cpu-baseline --no-single-run --no-scan --e2e-steps 3 "$@" \
      > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name $*" gpurun_out/ab_$name.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = l["roofline"]
    print("%-24s kernel %.2f us  frac %.4f  step %.2f us  value %.1f G  clocks %s" % (
        sys.argv[1], r["us_per_launch"], r["frac"], l["ms_per_step"] * 1e3, l["value"] / 1e9, l["clocks"]["sm_mhz"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
