#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output (SASS view): samples and executed instructions per
region, top stall sites.  usage: ncu -i rep --page source --csv | python tools/ncu_src.py [top]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
# find the header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
tot_s = sum(int(r[col["# Samples"]] or 0) for r in data)
tot_i = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print("total samples %d, warp instructions %d, SASS lines %d" % (tot_s, tot_i, len(data)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(data)), key=lambda k: -int(data[k][col["# Samples"]] or 0))
print("--- top stall sites")
for k in order[:top]:
    r = data[k]
    s = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print("%5d %5.1f%% #%-5d ex %8s  %-60s %s" % (s, 100.0 * s / max(tot_s, 1), k, r[col["Instructions Executed"]],
                                               r[col["Source"]][:60], " ".join("%s:%d" % (n, v) for v, n in why if v)))
if len(sys.argv) > 2:  # dump a range of SASS lines
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for k in range(a, b):
        r = data[k]
        print("#%-5d s %4s ex %8s  %s" % (k, r[col["# Samples"]], r[col["Instructions Executed"]], r[col["Source"]][:100]))
