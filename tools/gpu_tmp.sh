ncu --set full --clock-control none --import-source on -k regex:rotate_tile -s 3 -c 1 -f -o gpurun_out/r2_rotate python tools/rotate_case.py 2048 > gpurun_out/rot.log 2>&1
ncu -i gpurun_out/r2_rotate.ncu-rep --page details > gpurun_out/r2_rotate_details.txt 2>&1
ncu -i gpurun_out/r2_rotate.ncu-rep --page source --csv > gpurun_out/r2_rotate_source.csv 2>&1
grep -E "rotate_tile|Duration|Grid Size|Block Size|Registers Per|Issue Slots Busy|Executed Ipc Active|Theoretical Occ|Achieved Occ|Dynamic Shared|Block Limit" gpurun_out/r2_rotate_details.txt
python tools/ncu_src.py 25 < gpurun_out/r2_rotate_source.csv | head -60
