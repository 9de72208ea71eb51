#!/usr/bin/env python
"""Quick device check of the column-block kernel (ivs_stream8.cu) against the streaming kernel
(bit for bit: same arithmetic, same summation trees) while developing.  Not a test.

    python tools/devcheck8.py [case ...]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
from inovesa_b200.params import kara_params  # noqa: E402
from devcheck import CASES, states  # noqa: E402

if __name__ == "__main__":
    capi.load()
    NEW = 0  # the kernel under test
    for f in os.environ.get("IVS_DEV_FLAG", "STREAM8").split("+"):
        NEW |= getattr(capi, "SIM_" + f)
    sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 768, 1024, 2048, 2049]
    ok = True
    for mode in (capi.ARITH_FMA, capi.ARITH_STRICT):
        with capi.Context(0, mode) as ctx:
            for n in sizes:
                p = kara_params(**CASES[n])
                for members, strips in ((1, 0), (3, 0), (2, 5)):
                    t0 = time.time()
                    a = states(ctx, p, capi.SIM_QPT2, [1, 7, 24], members)
                    b = states(ctx, p, NEW, [1, 7, 24], members) if strips == 0 else None
                    if b is None:
                        sim = capi.Simulation(ctx, p, members, flags=NEW, strips=strips)
                        b = []
                        for k in [1, 7, 24]:
                            sim.step(k)
                            b.append(sim.download())
                        sim.close()
                    for k, ((ga, pa, wa, _), (gb, pb, wb, _)) in enumerate(zip(a, b)):
                        eq = np.array_equal(ga, gb) and np.array_equal(pa, pb) and np.array_equal(wa, wb)
                        ok &= bool(eq)
                        bad = np.argwhere(ga != gb)
                        print("mode %d case %4d members %d strips %d chk %d: equal %s  (grid diff cells %d%s, prof eq %s, wake eq %s)  %.1fs" % (
                            mode, n, members, strips, k, eq, len(bad), (" first " + str(bad[0])) if len(bad) else "",
                            np.array_equal(pa, pb), np.array_equal(wa, wb), time.time() - t0))
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
