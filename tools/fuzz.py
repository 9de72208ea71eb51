#!/usr/bin/env python
"""Randomised sweep: the default step path (persistent kernel where the run fits the machine, else
the column-block kernel) vs one launch per step vs tile kernel vs unfused kernel sequence on random
parameter combinations (interpolation order, Fokker-Planck type / derivation, RF model, drift
polynomial, steps per period, current, grid size).  A development aid; tests/ has the fixed cases."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
from inovesa_b200.params import kara_params  # noqa: E402


def rel_l1(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-300)


if __name__ == "__main__":
    capi.load()
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    for mode in (capi.ARITH_STRICT, capi.ARITH_FMA):
        with capi.Context(0, mode) as ctx:
            for k in range(count):
                kw = dict(grid_size=int(rng.choice([256, 256, 512, 768, 1024])),
                          steps_per_Ts=int(rng.choice([500, 1000, 2000, 4000, 8000])),
                          interpolation=int(rng.choice([2, 3, 4, 4])), derivation=int(rng.choice([3, 4])),
                          fp_type=int(rng.choice([0, 1, 2, 3, 3])), linear_rf=int(rng.choice([0, 1])),
                          bunch_current=float(rng.choice([0.0, 20e-6, 100e-6, 300e-6])),
                          gap=float(rng.choice([-1.0, 0.0])))
                if rng.random() < 0.5:
                    kw.update(alpha1=float(rng.uniform(-5e-4, 5e-4)), alpha2=float(rng.uniform(-5e-3, 5e-3)))
                members = int(rng.choice([1, 2, 5]))
                steps = int(rng.choice([3, 17, 40]))
                p = kara_params(**kw)
                cur = [max(kw["bunch_current"], 1e-9) * (1 + 0.3 * i) for i in range(members)]
                out = {}
                err = None
                for name, flags in (("stream", 0), ("perstep", capi.SIM_NO_PERSISTENT), ("tiled", capi.SIM_TILED),
                                    ("unfused", capi.SIM_UNFUSED)):
                    try:
                        sim = capi.Simulation(ctx, p, members, flags=flags, bunch_currents=cur)
                        sim.step(steps)
                        out[name] = sim.download()
                        sim.close()
                    except capi.IvsError as e:
                        err = (name, e.code)
                        break
                if err:
                    print("%2d mode %d %s members %d steps %d: %s -> skipped" % (k, mode, kw, members, steps, err))
                    continue
                e_t = rel_l1(out["stream"][0], out["tiled"][0])
                e_u = rel_l1(out["stream"][0], out["unfused"][0])
                same = all(np.array_equal(a, b) for a, b in zip(out["stream"][:3], out["perstep"][:3]))  # bit for bit
                ok = same and e_t < 1e-6 and e_u < 1e-5 and np.all(np.isfinite(out["stream"][0]))
                bad += not ok
                print("%2d mode %d n %4d it %d der %d fp %d lin %d I %.0e gap %+.0f spT %4d m %d s %2d: vs tiled %.1e vs unfused %.1e persistent == per-step %s %s"
                      % (k, mode, kw["grid_size"], kw["interpolation"], kw["derivation"], kw["fp_type"], kw["linear_rf"],
                         kw["bunch_current"], kw["gap"], kw["steps_per_Ts"], members, steps, e_t, e_u, same, "ok" if ok else "MISMATCH"),
                      flush=True)
    print("FUZZ OK" if not bad else "FUZZ: %d mismatches" % bad)
    sys.exit(1 if bad else 0)
