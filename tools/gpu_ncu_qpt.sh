#!/bin/bash
# ncu --set full of one stream_step_kernel launch for one and two quads per thread (development aid)
mkdir -p gpurun_out
for q in 2 1; do
  ncu --set full --clock-control none --import-source on -k regex:stream_step -s 3 -c 1 -f -o gpurun_out/r2_qpt$q \
      python bench.py --steps 6 --warmup 3 --particles 0 --qpt $q --no-cpu-baseline --no-single-run --no-scan --e2e-steps 3 --repeats 1 \
      > gpurun_out/ncu_qpt$q.log 2>&1
  ncu -i gpurun_out/r2_qpt$q.ncu-rep --page details > gpurun_out/r2_qpt${q}_details.txt 2>&1
  ncu -i gpurun_out/r2_qpt$q.ncu-rep --page raw --csv > gpurun_out/r2_qpt${q}_raw.csv 2>&1
  ncu -i gpurun_out/r2_qpt$q.ncu-rep --page source --csv > gpurun_out/r2_qpt${q}_source.csv 2>&1
  ls -la gpurun_out/r2_qpt$q.ncu-rep
  sz=$(stat -c %s gpurun_out/r2_qpt$q.ncu-rep)
  if [ "$sz" -gt 20000000 ]; then rm gpurun_out/r2_qpt$q.ncu-rep; fi
done
