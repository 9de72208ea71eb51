#!/bin/bash
# development round of the persistent kernel: its tests, then the single-run timings (and the per-phase
# clocks of a build with -DIVS_RES_CLOCKS=1 if there is one)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persistent.py -x -q > gpurun_out/pytest_persist.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_persist.log
python tools/single_runs.py 2000 > gpurun_out/single_runs.log 2>&1; grep -E "fused|per-step|tiled" gpurun_out/single_runs.log
if [ -f inovesa_b200/lib/variants/libivs_clocks.so ]; then
  IVS_LIB=$PWD/inovesa_b200/lib/variants/libivs_clocks.so python tools/single_runs.py 2000 2>&1 | grep "ivs resident" | sort -u
fi
