#!/usr/bin/env python
"""Generates tests/golden/impedance_tables.npz: the ORACLE's wake and radiation impedance tables
(oracle/vfps_oracle.c: orc_make_impedance, restating src/Z/ParallelPlatesCSR.cpp:25-70,
ResistiveWall.cpp:24-57, CollimatorImpedance.cpp:12-25, ImpedanceFactory.cpp:21-84) for the
BASELINE configs that use the parallel-plates CSR impedance.  The oracle evaluates the Airy
functions by quadrature (independent of any library), which takes minutes per table -- too slow
for a test, so the tables are committed and handed to the oracle simulation through
Oracle.sim(..., zwake=, zrad=).  Re-run after any change to the oracle's impedance code:

    python tests/golden/make_impedance_tables.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from inovesa_b200.params import kara_params  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

# name -> ring parameters (tests/test_gpu_sim.py uses the same dictionaries)
CASES = {
    "config2_512": dict(grid_size=512, steps_per_Ts=1000, gap=0.032, bunch_current=300e-6),
    "config3_1024": dict(grid_size=1024, steps_per_Ts=4000, gap=0.032, wall_conductivity=5.8e7,
                         collimator=0.008, bunch_current=200e-6),
}

def _table(job):
    name, kw, which = job
    o = Oracle(False)
    p = kara_params(**kw)
    d = o.derive(p)
    if which == "zwake":
        z, nz = o.make_impedance(d.padded_bins, d.fmax, d.R_bend, d.f_rev, p.gap, p.use_csr,
                                 p.wall_conductivity, p.wall_susceptibility, p.collimator)
        assert nz
    else:  # radiation impedance: CSR only (src/main.cpp:724-727)
        z, _ = o.make_impedance(d.padded_bins, d.fmax, d.R_bend, d.f_rev, p.gap if p.gap > 0 else -1.0)
    print(name, which, d.padded_bins, float(np.abs(z).max()), flush=True)
    return name + "_" + which, z


if __name__ == "__main__":
    import multiprocessing as mp
    jobs = [(name, kw, which) for name, kw in CASES.items() for which in ("zwake", "zrad")]
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        out = dict(pool.map(_table, jobs))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "impedance_tables.npz"), **out)
