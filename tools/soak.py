#!/usr/bin/env python
"""Soak test of the step kernels (column-block, persistent): many steps on several shapes, watching for traps
(a stalled barrier or mailbox traps instead of hanging), NaNs and charge drift."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
from inovesa_b200.params import kara_params  # noqa: E402

CASES = [
    (dict(grid_size=2048, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3), 8, 100000),
    (dict(grid_size=2048, steps_per_Ts=4000), 1, 100000),
    (dict(grid_size=1024, steps_per_Ts=2000, bunch_current=300e-6), 5, 100000),
    (dict(grid_size=512, bunch_current=400e-6), 37, 100000),
    (dict(grid_size=256), 256, 100000),
    (dict(grid_size=256, bunch_current=600e-6), 3, 200000),
    # the persistent kernel (one cooperative launch per call)
    (dict(grid_size=256), 1, 400000),
    (dict(grid_size=512, bunch_current=300e-6, gap=0.032), 4, 200000),
    (dict(grid_size=1024, steps_per_Ts=4000, gap=0.032, wall_conductivity=5.8e7, collimator=0.008), 1, 100000),
    (dict(grid_size=768, steps_per_Ts=2000), 2, 100000),
]

if __name__ == "__main__":
    capi.load()
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    with capi.Context(0, capi.ARITH_FMA) as ctx:
        for kw, members, steps in CASES:
            steps = int(steps * scale)
            p = kara_params(**kw)
            sim = capi.Simulation(ctx, p, members)
            t0 = time.time()
            sim.step(steps)
            g, pr, w, _ = sim.download()
            dt = time.time() - t0
            d = capi.derive(p)
            ws = capi.simpson_weights(p.grid_size, d.dq).astype(np.float64)
            q = (pr.astype(np.float64) * ws).sum(1)
            # far above threshold the bunch blows up beyond the grid and charge leaves it, in the
            # reference too (tools/charge_check.py): finite and bounded is the criterion here
            ok = np.all(np.isfinite(g)) and np.all(q > 0.05) and np.all(q < 1.01)
            print("n %4d members %3d steps %6d: %.1f s, %.1f us/step, charge %.6f..%.6f %s"
                  % (p.grid_size, members, steps, dt, dt / steps * 1e6, q.min(), q.max(), "ok" if ok else "BAD"), flush=True)
            sim.close()
            if not ok:
                sys.exit(1)
    print("SOAK OK")
