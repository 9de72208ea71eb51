# device times of the train kernels (cold, serialised): the Toeplitz kernel of the kernel sequence and the TRAIN
# instantiation of the column-block kernel, last launches of tools/train_times.py 40
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum --clock-control none -k regex:"train_wake_kernel|stream8_step_kernel" --csv --log-file gpurun_out/r2_train_launches.csv python tools/train_times.py 40 > gpurun_out/train_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_train_launches.csv")) if len(r) > 10]
hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
last = {}
for r in rows[1:]:
    key = (r[col["Kernel Name"]][:70], r[col["Grid Size"]], r[col["Block Size"]], r[col["Metric Name"]])
    last[key] = r[col["Metric Value"]] + " " + r[col["Metric Unit"]]
for k in sorted(last):
    print(" | ".join(k), "=", last[k])
PY
