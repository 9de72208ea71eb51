#!/bin/bash
# ncu --set full of one launch of a step kernel (development aid):
#   tools/gpu_ncu.sh NAME QPT KERNEL_REGEX   -> gpurun_out/NAME_{details.txt,raw.csv,source.csv}
name=$1; q=$2; rx=$3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/$name \
    python bench.py --steps 6 --warmup 3 --particles 0 --qpt $q --no-cpu-baseline --no-single-run --no-scan --e2e-steps 3 --repeats 1 \
    > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>&1
ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>&1
ls -la gpurun_out/$name.ncu-rep
sz=$(stat -c %s gpurun_out/$name.ncu-rep)
if [ "$sz" -gt 20000000 ]; then rm gpurun_out/$name.ncu-rep; fi
