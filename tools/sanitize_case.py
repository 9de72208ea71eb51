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
