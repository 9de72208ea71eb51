#!/bin/bash
# development round of the column-block kernel: parity against the streaming kernel, then its timing
mkdir -p gpurun_out
timeout 900 python tools/devcheck8.py "$@" > gpurun_out/devcheck8.log 2>&1; echo "devcheck8 rc=$?"; grep -c "equal True" gpurun_out/devcheck8.log; grep "equal False" gpurun_out/devcheck8.log | head -10; tail -1 gpurun_out/devcheck8.log
for q in ${IVS_DEV_QPT:-8}; do
timeout 300 python bench.py --steps 200 --warmup 30 --repeats 3 --no-cpu-baseline --no-single-run --no-scan --e2e-steps 3 --particles 0 --qpt $q > gpurun_out/bench_q$q.json 2> gpurun_out/bench_q$q.err
python - gpurun_out/bench_q$q.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = l["roofline"]
    print("%s kernel %.2f us frac %.4f" % (sys.argv[1], r["us_per_launch"], r["frac"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-2000:])
PY
done
