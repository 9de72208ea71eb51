#!/bin/bash
# single-run step times of the variants of libivs.so under inovesa_b200/lib/variants (development aid)
for so in inovesa_b200/lib/libivs.so inovesa_b200/lib/variants/libivs_*.so; do
  echo "== $so"
  IVS_LIB=$PWD/$so python tools/single_runs.py 3000 fused 2>&1 | grep -E "fused|resident" | grep -v "cta 0" | sort -u
done
