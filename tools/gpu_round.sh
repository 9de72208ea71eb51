#!/bin/bash
# one GPU call of a development round: the GPU tests, then the benchmark with one and two quads per
# thread (development aid; outputs under gpurun_out/)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for q in 2 1; do
  python bench.py --steps 200 --warmup 20 --qpt $q --no-cpu-baseline --no-single-run --no-scan --e2e-steps 3 \
      > gpurun_out/bench_qpt$q.json 2> gpurun_out/bench_qpt$q.err
  python - gpurun_out/bench_qpt$q.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = l["roofline"]
    print("%s kernel %.2f us frac %.4f step %.2f us value %.1f G blocks %s" % (sys.argv[1], r["us_per_launch"], r["frac"], l["ms_per_step"]*1e3, l["value"]/1e9, l["repeats"]["ms_blocks"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-2000:])
PY
done
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver.json 2> gpurun_out/bench_driver.err
tail -c 3000 gpurun_out/bench_driver.json
