ncu --set full --clock-control none --cache-control none --import-source on -k regex:field_kernel -s 55 -c 1 -f -o gpurun_out/r2_field python tools/kernel_times.py 20 > gpurun_out/kt3.log 2>&1
ncu -i gpurun_out/r2_field.ncu-rep --page details > gpurun_out/r2_field_details.txt 2>&1
ncu -i gpurun_out/r2_field.ncu-rep --page source --csv > gpurun_out/r2_field_source.csv 2>&1
grep -E "Duration|Block Size|Issue Slots Busy|Executed Ipc Active|No Eligible" gpurun_out/r2_field_details.txt
python tools/ncu_src.py 30 < gpurun_out/r2_field_source.csv
