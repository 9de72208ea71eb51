#!/bin/bash
# one GPU call at a milestone: all GPU tests, the driver-style benchmark line, the default benchmark line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver.json 2> gpurun_out/bench_driver.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_driver.json", "gpurun_out/bench_default.json"):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        r = l["roofline"]
        print(f, "value %.1f G  step %.2f us  kernel %.2f us  frac %.4f  e2e %.2f G  scan %s" % (
            l["value"] / 1e9, l["ms_per_step"] * 1e3, r["us_per_launch"], r["frac"], l["e2e"]["value"] / 1e9,
            json.dumps(l.get("scan_256x256", {}).get("frac_of_hbm_peak"))))
        for k, v in l.get("single_run", {}).items():
            if isinstance(v, dict): print("   single", k, "%.2f us/step  (one launch per step %.2f)" % (v["us_per_step"], v["per_step_launch"]["us_per_step"]))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-2000:])
PY
