# ncu --set full of the kernels beside the fused step: field kernel (nmax 2048), unfused maps at 2048^2, train wake kernel
ncu --set full --clock-control none -k regex:"field_kernel|kick_y|kick_x|fokker_planck|project_x" -s 0 -c 400 -f -o gpurun_out/r2_side python tools/kernel_times.py 2 > gpurun_out/side.log 2>&1
ncu -i gpurun_out/r2_side.ncu-rep --page raw --csv > gpurun_out/r2_side_raw.csv 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_side_raw.csv")))
hdr = next(r for r in rows if r and r[0] == "ID")
i0 = rows.index(hdr)
col = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
print("units:", [rows[i0 + 1][col[w]] for w in want if w in col])
seen = {}
for r in rows[i0 + 2:]:
    if len(r) < len(hdr):
        continue
    key = (r[col["Kernel Name"]][:60], r[col["launch__grid_size"]])
    seen[key] = [r[col[w]] for w in want if w in col]   # the last launch of each (kernel, grid) pair
for k, v in seen.items():
    print(" | ".join(v))
PY
