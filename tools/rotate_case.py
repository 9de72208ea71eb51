#!/usr/bin/env python
"""a few launches of the RotationMap tile kernel at 2048^2 (for ncu; development aid)"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
with capi.Context(0, capi.ARITH_FMA) as ctx:
    g = np.random.default_rng(1).random((n, n), dtype=np.float32)
    d_in, d_out = ctx.to_device(g), ctx.malloc(g.nbytes)
    for _ in range(6):
        ctx.check(ctx.lib.ivs_rotate(ctx.h, d_in, d_out, n, 0.0015, -6.0, 6.0, -6.0, 6.0, 4, 0))
    ctx.sync()
