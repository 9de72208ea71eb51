#!/usr/bin/env python
"""The reference's CPU path (oracle/_ref: Inovesa's own sources compiled in place) timed on the host
cores for every BASELINE.json config -- the "reference CPU" columns of BASELINE.md's table:
configs 1, 2, 3, 5 as single runs on one core (the reference is single-threaded), config 4 (the
256-member 256^2 scan) as one replica per core.  Bounded to a few seconds per line.

    python tools/cpu_configs.py [seconds per line]
"""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    cores = os.cpu_count() or 1
    model = ""
    try:
        model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = {"cpu": model, "cores": cores, "budget_s_per_line": budget, "lines": []}
    for name, wl in (("1", "config1"), ("2", "config2"), ("3", "config3"), ("5", "config5")):
        p = bench.workload_params(wl)
        kind, done, dt = bench._ref_worker((bench._pdict(p), 10 ** 6, 2, budget, True))
        n = p.grid_size
        out["lines"].append({"config": name, "grid": n, "kind": kind, "cores": 1, "steps": done, "seconds": dt,
                             "steps_per_s": done / dt, "cell_updates_per_s": float(n) * n * done / dt})
        print(json.dumps(out["lines"][-1]), flush=True)
    p = bench.workload_params("config4")
    with mp.get_context("spawn").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(bench._ref_worker, [(bench._pdict(p), 10 ** 6, 2, budget, True)] * cores)
        wall = time.perf_counter() - t0
    total = sum(r[1] for r in res)
    tmax = max(r[2] for r in res)
    n = p.grid_size
    out["lines"].append({"config": "4", "grid": n, "kind": res[0][0], "cores": cores, "steps": total, "seconds": tmax,
                         "ensemble_steps_per_s": total / tmax, "cell_updates_per_s": float(n) * n * total / tmax,
                         "wall_s": wall})
    print(json.dumps(out["lines"][-1]), flush=True)
    print(json.dumps(out))
