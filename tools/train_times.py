#!/usr/bin/env python
"""Step time of a train of bunches (multi-bunch filling pattern) on the kernel sequence, and of the train's
wake kernel alone (development aid).  python tools/train_times.py [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
from inovesa_b200.params import kara_params  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
f32 = np.float32
with capi.Context(0, capi.ARITH_FMA) as ctx:
    for n, filling in ((256, [1e-4] * 4), (256, [1e-4] * 16), (1024, [1e-4] * 4), (2048, [1e-4] * 4)):
        p = kara_params(grid_size=n, steps_per_Ts=4000 if n >= 1024 else 1000)
        lay = capi.train_layout(p, filling, n + n // 4)
        d, nb = lay["d"], len(lay["buckets"])
        z, _ = capi.wake_impedance(lay["params"], d, nfreqs=lay["spaced_bins"])
        g0 = capi.initial_grid(d, n, p.zoom)
        grids = np.stack([g0 * f32(fr) for fr in lay["fractions"]]).astype(f32)
        res = {}
        for name, train, fl in (("independent bunches", False, capi.SIM_UNFUSED), ("train", True, capi.SIM_UNFUSED),
                                ("fused independent", False, capi.SIM_NO_PERSISTENT), ("fused train", True, 0)):
            sim = capi.Simulation(ctx, lay["params"], nb, grid=grids, flags=fl)
            if train:
                sim.set_bunch_train(lay["buckets"], lay["spacing_bins"], z, lay["wakescaling"])
            sim.step(32); ctx.sync()
            t0 = time.perf_counter()
            sim.step(steps); ctx.sync()
            res[name] = (time.perf_counter() - t0) / steps * 1e6
            sim.close()
        print("n=%-5d %2d bunches: kernel sequence %.1f us per step with independent wakes (one FFT per bunch), "
              "%.1f us as a train (Toeplitz products, %d x %d windows); column-block kernel: %.1f us independent, %.1f us as a train"
              % (n, nb, res["independent bunches"], res["train"], nb, nb, res["fused independent"], res["fused train"]))
