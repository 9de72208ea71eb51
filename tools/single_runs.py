#!/usr/bin/env python
"""Step time of ONE simulation (no ensemble) at the grid sizes of BASELINE.json's configs: the
latency-bound view (state L2-resident, one strip per SM or fewer).  Development aid.

    python tools/single_runs.py [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
from inovesa_b200.params import kara_params  # noqa: E402

CASES = [
    ("config 1: 256^2, free-space CSR, 1000 steps/Ts", dict(grid_size=256, steps_per_Ts=1000, gap=-1.0, bunch_current=50e-6)),
    ("config 2: 512^2, parallel plates, above threshold", dict(grid_size=512, gap=0.032, bunch_current=300e-6, steps_per_Ts=2000)),
    ("config 3: 1024^2, CSR + wall + collimator", dict(grid_size=1024, steps_per_Ts=4000, gap=0.032, wall_conductivity=5.8e7, collimator=0.008)),
    ("config 5: 2048^2, sin RF, alpha1/alpha2", dict(grid_size=2048, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3, gap=-1.0, bunch_current=50e-6)),
]

if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    only = sys.argv[2:]  # labels to run (default: all)
    capi.load()
    with capi.Context(0, capi.ARITH_FMA) as ctx:
        for name, kw in CASES:
            p = kara_params(**kw)
            for flags, label in ((0, "fused"), (capi.SIM_NO_PERSISTENT, "per-step"), (capi.SIM_TILED, "tiled"), (capi.SIM_QPT2, "ring-2q"),
                                 (capi.SIM_UNFUSED, "unfused")):
                if only and label not in only:
                    continue
                sim = capi.Simulation(ctx, p, 1, flags=flags)
                sim.step(64)
                ctx.sync()
                t0 = time.perf_counter()
                sim.step(steps)
                ctx.sync()
                dt = time.perf_counter() - t0
                sim.close()
                n = p.grid_size
                print("%-52s %-8s %7.2f us/step  %8.0f steps/s  %7.1f G cell-updates/s" % (
                    name, label, dt / steps * 1e6, steps / dt, n * n * steps / dt / 1e9), flush=True)
