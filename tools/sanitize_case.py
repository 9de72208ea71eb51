#!/usr/bin/env python
"""a few steps of small simulations through every step path, for compute-sanitizer (development aid):
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
from inovesa_b200.params import kara_params  # noqa: E402

capi.load()
with capi.Context(0, capi.ARITH_FMA) as ctx:
    rng = np.random.default_rng(0)
    for n, members in ((256, 1), (256, 3), (512, 1)):
        p = kara_params(grid_size=n, steps_per_Ts=2000, bunch_current=100e-6)
        ref = None
        for name, fl in (("persistent", 0), ("column-block", capi.SIM_NO_PERSISTENT), ("tiled", capi.SIM_TILED), ("unfused", capi.SIM_UNFUSED)):
            sim = capi.Simulation(ctx, p, members, flags=fl)
            sim.step(6)
            g = sim.download()[0]
            sim.close()
            ref = g if ref is None else ref
            print(n, members, name, "max |diff| vs persistent %.2e" % float(np.abs(g - ref).max()), flush=True)
    parts = (rng.random((500, 2)) * 200 + 28).astype(np.float32)
    sim = capi.Simulation(ctx, kara_params(grid_size=256), 2, particles=parts)
    sim.step(5)
    sim.observe()
    sim.close()
    print("tracking + observe ok")
    # a train of bunches (multi-bunch filling pattern) on the kernel sequence
    p = kara_params(grid_size=128, wall_conductivity=5.8e7, collimator=0.008, gap=0.032)
    lay = capi.train_layout(p, [1e-4, 0, 2e-4], 160)
    z, _ = capi.wake_impedance(lay["params"], lay["d"], nfreqs=lay["spaced_bins"])
    g0 = capi.initial_grid(lay["d"], 128)
    grids = np.stack([g0 * np.float32(fr) for fr in lay["fractions"]])
    sim = capi.Simulation(ctx, lay["params"], 2, grid=grids, flags=capi.SIM_UNFUSED)
    sim.set_bunch_train(lay["buckets"], lay["spacing_bins"], z, lay["wakescaling"])
    sim.step(20)
    sim.fields(spectrum=False)
    sim.close()
    print("train ok")
