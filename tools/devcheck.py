#!/usr/bin/env python
"""Quick device check while developing a step kernel: streaming vs tile vs unfused kernel
sequence on a few grids (seconds on a B200).  Not a test -- tests/ holds the parity suite.

    python tools/devcheck.py [grid sizes ...]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
from inovesa_b200.params import kara_params  # noqa: E402


def rel_l1(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-300)


def states(ctx, p, flags, steps, members=1):
    sim = capi.Simulation(ctx, p, members, flags=flags)
    out = []
    for k in steps:
        sim.step(k)
        out.append(sim.download())
    sim.close()
    return out


CASES = {
    64: dict(grid_size=64, interpolation=2, gap=0.0),
    128: dict(grid_size=128),
    256: dict(grid_size=256),
    512: dict(grid_size=512, gap=0.032, bunch_current=300e-6, steps_per_Ts=2000),
    768: dict(grid_size=768, steps_per_Ts=2000),
    1536: dict(grid_size=1536, steps_per_Ts=4000, linear_rf=0),
    1024: dict(grid_size=1024, steps_per_Ts=4000, gap=0.032, wall_conductivity=5.8e7, collimator=0.008),
    2048: dict(grid_size=2048, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3),
    2049: dict(grid_size=2048, steps_per_Ts=1000),
}

if __name__ == "__main__":
    capi.load()
    sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 2048]
    ok = True
    for mode in (capi.ARITH_STRICT, capi.ARITH_FMA):
        with capi.Context(0, mode) as ctx:
            for n in sizes:
                p = kara_params(**CASES[n])
                t0 = time.time()
                a = states(ctx, p, 0, [1, 7, 24])
                t = states(ctx, p, capi.SIM_TILED, [1, 7, 24])
                b = states(ctx, p, capi.SIM_UNFUSED, [1, 7, 24])
                for k, ((ga, pa, wa, _), (gt, pt, wt, _), (gb, pb, wb, _)) in enumerate(zip(a, t, b)):
                    eq = np.array_equal(ga, gt)
                    ok &= (eq if k == 0 else rel_l1(ga, gt) < 1e-6)
                    print("mode %d n %4d chk %d: stream==tiled %s (L1 %.2e, wake eq %s, prof %.1e) | vs unfused L1 %.2e"
                          % (mode, n, k, eq, rel_l1(ga, gt), np.array_equal(wa, wt),
                             np.abs(pa - pt).max() / np.abs(pt).max(), rel_l1(ga, gb)), flush=True)
                print("   %.1f s" % (time.time() - t0), flush=True)
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
