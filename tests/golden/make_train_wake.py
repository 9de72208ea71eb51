#!/usr/bin/env python
"""Golden vectors of ElectricField::wakePotential for a train of bunches (PhaseSpace::nb = 3), made by the
reference's own sources compiled in place (oracle/_ref, needs /root/reference): the same cases as
tests/test_oracle_vs_ref.py::test_train_wake_potential_bit_exact, for machines without the reference.

    python tests/golden/make_train_wake.py   ->  tests/golden/train_wake.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import Ref  # noqa: E402

f32 = np.float32
ref = Ref(False)
n, nmax = 32, 256
g0 = ref.initial_grid(n, -12, 12, -12, 12)
grids = np.stack([g0 * f32(0.5), np.roll(g0, 3, axis=0) * f32(0.3), np.roll(g0, -2, axis=0) * f32(0.2)])
profiles = np.stack([ref.reduce(g, -12, 12, -12, 12)[0] for g in grids])
z, _ = ref.make_impedance(nmax, 1e12, 47.7, 1e6, -1)
out = dict(profiles=profiles, z=z)
for k, (buckets, spacing) in enumerate((([4, 2, 0], 40), ([2, 1, 0], 32), ([0, 1, 5], 44))):
    wake, padded, used = ref.train_wake_potential(grids, -12, 12, -12, 12, buckets, spacing, z, 1e-3)
    out["case%d_buckets" % k] = np.asarray(buckets, np.uint32)
    out["case%d_spacing" % k] = np.asarray(spacing, np.uint32)
    out["case%d_scaling" % k] = np.asarray(used, f32)
    out["case%d_wake" % k] = wake
    out["case%d_padded" % k] = padded
np.savez_compressed(os.path.join(HERE, "train_wake.npz"), **out)
print("wrote train_wake.npz:", {k: v.shape for k, v in out.items()})
