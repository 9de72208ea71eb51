mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_e2e.json").read().strip().splitlines()[-1])
print(json.dumps(l["e2e"], indent=1)[:1800]); print("value", l["value"]/1e9)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_kernel_times_ncu.csv python tools/kernel_times.py 20 > gpurun_out/kt.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_kernel_times_ncu.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size"); bi=hdr.index("Block Size")
c=collections.OrderedDict()
for r in rows[1:]:
    key=(r[ki].split("(")[0][-40:], r[gi], r[bi])
    c.setdefault(key, []).append(float(r[vi].replace(",","")))
for k,v in c.items(): print("%-42s grid %-18s block %-14s %3d launches  mean %8.2f us  min %8.2f" % (k[0],k[1],k[2],len(v),sum(v)/len(v)/1e3,min(v)/1e3))
PY
python tools/kernel_times.py 200 | grep -E "rotate|field"
