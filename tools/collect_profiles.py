#!/usr/bin/env python
"""Copy the summaries of the last `tools/gpu_profiles.sh` + `tools/gpu_full.sh` run from gpurun_out/
(scratch) into profiles/ (tracked).  Development aid."""
import csv
import os
import shutil
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
for src, dst in (("r2_final_launches.csv", "r2_ncu_launches.csv"), ("r2_final_stream8_details.txt", "r2_ncu_stream8_step_details.txt"),
                 ("r2_final_track_details.txt", "r2_ncu_track_step_details.txt"),
                 ("r2_final_resident_details.txt", "r2_ncu_resident_run_details.txt"),
                 ("r2_final_kernel_times.txt", "r2_kernel_times.txt"), ("r2_final_single_runs.txt", "r2_single_runs.txt"),
                 ("bench_default.json", "r2_bench_default.json"), ("bench_driver.json", "r2_bench_driver_args.json")):
    shutil.copy(os.path.join(G, src), os.path.join(P, dst))
for k, dst in (("stream8", "r2_ncu_stream8_step_top_stalls.txt"), ("resident", "r2_ncu_resident_run_top_stalls.txt")):
    with open(os.path.join(G, "r2_final_%s_source.csv" % k)) as fi, open(os.path.join(P, dst), "w") as fo:
        subprocess.check_call([sys.executable, os.path.join(R, "tools", "ncu_src.py"), "25"], stdin=fi, stdout=fo)
out = []
titles = {"stream8": "stream8_step_kernel, one launch on the benchmark batch (26 x 2048^2, 442 CTAs x 256 threads)",
          "track": "track_step_staged_kernel, one launch (26 x 1e5 particles, 442 blocks x 512 threads)",
          "resident": "resident_run_kernel, one launch = 2000 steps of one 256^2 run (128 CTAs x 512 threads, cooperative)"}
names = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
         "lts__t_sectors_srcunit_tex_op_write.sum", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
         "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "smsp__issue_active.avg.pct_of_peak_sustained_active")
for k, title in titles.items():
    rows = list(csv.reader(open(os.path.join(G, "r2_final_%s_raw.csv" % k))))
    hdr = next(r for r in rows if r and r[0] == "ID")
    i = rows.index(hdr)
    out.append("== " + title + "  (ncu --set full --clock-control none, raw page)")
    for n in names:
        if n in hdr:
            j = hdr.index(n)
            out.append("  %-62s %s %s" % (n, rows[i + 2][j], rows[i + 1][j]))
    out.append("")
open(os.path.join(P, "r2_ncu_raw_metrics.txt"), "w").write("\n".join(out))
with open(os.path.join(G, "pytest_gpu.log")) as f:
    open(os.path.join(P, "r2_pytest_gpu.log"), "w").write("".join(f.readlines()[-6:]))
print("\n".join(out))
