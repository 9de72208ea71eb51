ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:field_kernel --csv --log-file gpurun_out/field_warm.csv python tools/kernel_times.py 20 > gpurun_out/kt2.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/field_warm.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); bi=hdr.index("Block Size")
c=collections.OrderedDict()
for r in rows[1:]:
    c.setdefault(r[bi], []).append(float(r[vi].replace(",","")))
for k,v in c.items(): print("field_kernel block %-14s %3d launches  mean %8.2f us  min %8.2f (warm: --cache-control none)" % (k,len(v),sum(v)/len(v)/1e3,min(v)/1e3))
PY
