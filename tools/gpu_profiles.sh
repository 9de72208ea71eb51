#!/bin/bash
# round-end evidence in one GPU call: launch list of a short benchmark run, ncu --set full of the
# column-block step kernel (benchmark batch) and of the persistent kernel (one 512^2 run), kernel
# times beside the step.  Outputs under gpurun_out/r2_final_*; the summaries go to profiles/.
mkdir -p gpurun_out
B="--no-cpu-baseline --no-single-run --no-scan --e2e-steps 3 --repeats 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/r2_final_launches.csv \
    python bench.py --steps 20 --warmup 5 $B > gpurun_out/r2_final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream8_step -s 3 -c 1 -f -o gpurun_out/r2_final_stream8 \
    python bench.py --steps 6 --warmup 3 --particles 0 $B > gpurun_out/r2_final_stream8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:track_step -s 3 -c 1 -f -o gpurun_out/r2_final_track \
    python bench.py --steps 6 --warmup 3 $B > gpurun_out/r2_final_track.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resident_run -s 1 -c 1 -f -o gpurun_out/r2_final_resident \
    python tools/single_runs.py 2000 fused > gpurun_out/r2_final_resident.log 2>&1
for k in stream8 track resident; do
  ncu -i gpurun_out/r2_final_$k.ncu-rep --page details > gpurun_out/r2_final_${k}_details.txt 2>&1
  ncu -i gpurun_out/r2_final_$k.ncu-rep --page raw --csv > gpurun_out/r2_final_${k}_raw.csv 2>&1
  ncu -i gpurun_out/r2_final_$k.ncu-rep --page source --csv > gpurun_out/r2_final_${k}_source.csv 2>&1
done
python tools/kernel_times.py 200 > gpurun_out/r2_final_kernel_times.txt 2>&1
python tools/single_runs.py 4000 > gpurun_out/r2_final_single_runs.txt 2>&1
tail -30 gpurun_out/r2_final_kernel_times.txt
ls -la gpurun_out/r2_final_*
