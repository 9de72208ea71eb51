"""ctypes access to the checkers under oracle/  -- TEST INFRASTRUCTURE ONLY.

Loads oracle/liboracle[_fma].so (our C restatement, vfps_oracle.c) and
oracle/_ref/libinovesa_ref[_fma].so (the reference's own sources compiled in place, see
oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module; nothing under inovesa_b200/ does.
"""
import ctypes as C
import os

import numpy as np

from inovesa_b200.params import Derived, RingParams

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u32, cf, cd, ci = C.c_uint32, C.c_float, C.c_double, C.c_int
P = C.c_void_p


def _opt(a):
    """optional float array -> pointer or None"""
    return None if a is None else a.ctypes.data_as(P)


def lib_path(kind, fma):
    if kind == "oracle":
        return os.path.join(HERE, "liboracle_fma.so" if fma else "liboracle.so")
    return os.path.join(HERE, "_ref", "libinovesa_ref_fma.so" if fma else "libinovesa_ref.so")


def available(kind, fma=False):
    return os.path.exists(lib_path(kind, fma))


class Oracle:
    """our C restatement (oracle/vfps_oracle.c)"""

    def __init__(self, fma=False):
        self.fma = fma
        self.lib = L = C.CDLL(lib_path("oracle", fma))
        assert bool(L.orc_is_fma()) == bool(fma)
        L.orc_ruler_delta.restype = cf
        L.orc_ruler_delta.argtypes = [u32, cf, cf]
        L.orc_ruler_zerobin.restype = cf
        L.orc_ruler_zerobin.argtypes = [u32, cf, cf]
        L.orc_ruler_axis.argtypes = [u32, cf, cf, _fp]
        L.orc_calc_coefficients.argtypes = [cf, u32, _fp]
        L.orc_kick_update_sm.argtypes = [_fp, u32, u32, u32, _up, _fp]
        L.orc_kick_apply.argtypes = [u32, u32, ci, _fp, _fp, _fp]
        L.orc_kick_track.argtypes = [u32, ci, _fp, _fp, u32]
        L.orc_rf_offsets_linear.argtypes = [u32, cf, cf, cf, cf, cf, _fp]
        L.orc_rf_offsets_nonlinear.argtypes = [u32, cf, cf, cf, cf, cf, cf, cf, cf, cf, cf, _fp]
        L.orc_drift_offsets.argtypes = [u32, cf, cf, cf, cf, cf, _fp, u32, cf, _fp]
        L.orc_fp_table.argtypes = [u32, cf, cf, cf, u32, u32, _up, _fp]
        L.orc_fp_weights5.argtypes = [u32, cf, cf, cf, u32, u32, _fp]
        L.orc_fp_apply.argtypes = [u32, cf, cf, cf, u32, u32, _fp, _fp]
        L.orc_fp_track.argtypes = [u32, cf, cf, cf, u32, u32, u32, _fp, _fp, u32]
        L.orc_rotation_apply.restype = ci
        L.orc_rotation_apply.argtypes = [u32, cf, cf, cf, cf, cf, u32, ci, u32, _fp, _fp]
        L.orc_rotation_track.argtypes = [u32, cf, cf, cf, cf, cf, _fp, u32]
        L.orc_simpson_weights.argtypes = [u32, cf, _fp]
        L.orc_project_x.argtypes = [u32, _fp, _fp, _fp]
        L.orc_project_y.argtypes = [u32, _fp, _fp, _fp]
        L.orc_integrate.restype = cf
        L.orc_integrate.argtypes = [u32, _fp, _fp]
        L.orc_normalize.argtypes = [u32, _fp, cf, cf]
        L.orc_moments.argtypes = [u32, _fp, _fp, cf, cf, _fp, _fp, _fp]
        L.orc_initial_grid.argtypes = [u32, cf, cf, cf, cf, cd, _fp]
        L.orc_wake_potential.argtypes = [u32, _fp, _fp, u32, cf, _fp, P, P]
        L.orc_csr_power.argtypes = [u32, _fp, _fp, u32, cf, cf, cf, P, _fp]
        L.orc_train_wake_potential.restype = ci
        L.orc_train_wake_potential.argtypes = [u32, u32, _fp, _up, u32, _fp, u32, cf, _fp, P]
        L.orc_z_free_space.argtypes = [C.c_size_t, cf, cf, _fp]
        L.orc_z_parallel_plates.argtypes = [C.c_size_t, cf, cf, cd, _fp]
        L.orc_z_resistive_wall.argtypes = [C.c_size_t, cf, cf, cd, cd, cd, cd, _fp]
        L.orc_z_const.argtypes = [C.c_size_t, cf, cf, _fp]
        L.orc_z_collimator.argtypes = [C.c_size_t, cd, cd, _fp]
        L.orc_make_impedance.restype = ci
        L.orc_make_impedance.argtypes = [C.c_size_t, cf, cd, cd, cd, ci, cd, cd, cd, _fp]
        L.orc_airy.restype = ci
        L.orc_airy.argtypes = [cd] + [C.POINTER(cd)] * 4
        L.orc_derive.argtypes = [C.POINTER(RingParams), C.POINTER(Derived)]
        L.orc_sim_create.restype = P
        L.orc_sim_create.argtypes = [C.POINTER(RingParams), P, P, u32]
        L.orc_sim_create_z.restype = P
        L.orc_sim_create_z.argtypes = [C.POINTER(RingParams), P, P, u32, P, P]
        L.orc_sim_destroy.argtypes = [P]
        L.orc_sim_step.argtypes = [P, u32]
        L.orc_sim_derived.argtypes = [P, C.POINTER(Derived)]
        L.orc_sim_get.argtypes = [P, P, P, P, P]
        L.orc_sim_tables.argtypes = [P, P, P, P, P, P]
        L.orc_sim_set_tables.argtypes = [P, P, P, P, P]
        L.orc_sim_observe.argtypes = [P, _fp, P, P]
        L.orc_sim_set_rf_modulation.argtypes = [P, _fp, u32]
        L.orc_rf_calc_kick.argtypes = [ci, u32] + [cf] * 13 + [_fp]
        L.orc_rf_modulation.argtypes = [u32, cf, cf, cf, cd, cd, cf, _fp, _fp]

    # --- single functions -------------------------------------------------------------
    def coefficients(self, f, it):
        ic = np.zeros(4, f32)
        self.lib.orc_calc_coefficients(f, it, ic)
        return ic[:it]

    def kick_apply(self, grid, offset, it, axis):
        n = grid.shape[0]
        out = np.empty_like(grid)
        self.lib.orc_kick_apply(n, it, axis, np.ascontiguousarray(offset, f32),
                                np.ascontiguousarray(grid, f32), out)
        return out

    def kick_track(self, n, axis, offset, particles):
        p = np.ascontiguousarray(particles, f32).copy()
        self.lib.orc_kick_track(n, axis, np.ascontiguousarray(offset, f32), p, len(p))
        return p

    def fp_weights5(self, n, pmin, pmax, e1, fptype=3, derivation=4):
        w = np.zeros((n, 5), f32)
        self.lib.orc_fp_weights5(n, pmin, pmax, e1, fptype, derivation, w)
        return w

    def fp_apply(self, grid, pmin, pmax, e1, fptype=3, derivation=4):
        n = grid.shape[0]
        out = np.empty_like(grid)
        self.lib.orc_fp_apply(n, pmin, pmax, e1, fptype, derivation,
                              np.ascontiguousarray(grid, f32), out)
        return out

    def fp_track(self, grid, pmin, pmax, e1, fptype, derivation, fptrack, particles):
        n = grid.shape[0]
        p = np.ascontiguousarray(particles, f32).copy()
        self.lib.orc_fp_track(n, pmin, pmax, e1, fptype, derivation, fptrack,
                              np.ascontiguousarray(grid, f32), p, len(p))
        return p

    def rotation_apply(self, grid, qmin, qmax, pmin, pmax, angle, it, clamp=0, rotmapsize=0):
        n = grid.shape[0]
        out = np.zeros_like(grid)
        rc = self.lib.orc_rotation_apply(n, qmin, qmax, pmin, pmax, angle, it, clamp, rotmapsize,
                                         np.ascontiguousarray(grid, f32), out)
        if rc == 1:
            raise ValueError("invalid_argument")
        return out

    def rotation_track(self, n, qmin, qmax, pmin, pmax, angle, particles):
        p = np.ascontiguousarray(particles, f32).copy()
        self.lib.orc_rotation_track(n, qmin, qmax, pmin, pmax, angle, p, len(p))
        return p

    def simpson_weights(self, n, delta):
        ws = np.zeros(n, f32)
        self.lib.orc_simpson_weights(n, delta, ws)
        return ws

    def ruler(self, n, lo, hi):
        ax = np.zeros(n, f32)
        self.lib.orc_ruler_axis(n, lo, hi, ax)
        return ax, float(self.lib.orc_ruler_delta(n, lo, hi)), float(self.lib.orc_ruler_zerobin(n, lo, hi))

    def reduce(self, grid, qmin, qmax, pmin, pmax):
        """projections, integral and moments like Ref.reduce"""
        n = grid.shape[0]
        g = np.ascontiguousarray(grid, f32)
        a0, dq, _ = self.ruler(n, qmin, qmax)
        a1, dp, _ = self.ruler(n, pmin, pmax)
        ws = self.simpson_weights(n, dq)
        px, py = np.zeros(n, f32), np.zeros(n, f32)
        self.lib.orc_project_x(n, g, ws, px)
        self.lib.orc_project_y(n, g, ws, py)
        integral = float(self.lib.orc_integrate(n, px, ws))
        m = np.zeros(6, f32)
        for k, (pr, ax, d) in enumerate(((px, a0, dq), (py, a1, dp))):
            a, b, c = (np.zeros(1, f32) for _ in range(3))
            self.lib.orc_moments(n, pr, ax, d, integral, a, b, c)
            m[3 * k:3 * k + 3] = a[0], b[0], c[0]
        return px, py, integral, m

    def initial_grid(self, n, qmin, qmax, pmin, pmax, zoom=1.0):
        g = np.zeros((n, n), f32)
        self.lib.orc_initial_grid(n, qmin, qmax, pmin, pmax, zoom, g)
        return g

    def wake_potential(self, profile, z, wakescaling, full=False):
        n, nmax = len(profile), len(z) // 2 if z.ndim == 1 else z.shape[0]
        wake = np.zeros(n, f32)
        ff = np.zeros((nmax // 2 + 1, 2), f32)
        pw = np.zeros(nmax, f32)
        self.lib.orc_wake_potential(n, np.ascontiguousarray(profile, f32),
                                    np.ascontiguousarray(z, f32).reshape(-1), nmax, wakescaling,
                                    wake, _opt(ff), _opt(pw))
        return (wake, ff, pw) if full else wake

    def train_wake_potential(self, profiles, buckets, spacing_bins, z, wakescaling, full=False):
        """wake potentials [nb][n] of a train of bunches (profiles [nb][n], bucket numbers, spacing in bins)"""
        profiles = np.ascontiguousarray(profiles, f32)
        nb, n = profiles.shape
        nmax = z.reshape(-1, 2).shape[0]
        wake = np.zeros((nb, n), f32)
        pw = np.zeros(nmax, f32)
        rc = self.lib.orc_train_wake_potential(n, nb, profiles.reshape(-1), np.ascontiguousarray(buckets, np.uint32),
                                               spacing_bins, np.ascontiguousarray(z, f32).reshape(-1), nmax,
                                               wakescaling, wake.reshape(-1), _opt(pw))
        if rc != 0:
            raise ValueError("a bunch of the train does not fit the padded profile")
        return (wake, pw) if full else wake

    def csr_power(self, profile, z, dq, f_scale, cutoff):
        n, nmax = len(profile), z.reshape(-1, 2).shape[0]
        spec = np.zeros(nmax, f32)
        power = np.zeros(1, f32)
        self.lib.orc_csr_power(n, np.ascontiguousarray(profile, f32),
                               np.ascontiguousarray(z, f32).reshape(-1), nmax, dq, f_scale,
                               cutoff, _opt(spec), power)
        return spec, float(power[0])

    def make_impedance(self, nfreqs, fmax, R_bend, f_rev, gap, use_csr=1, s=0.0, xi=0.0, coll=0.0):
        z = np.zeros((nfreqs, 2), f32)
        nz = self.lib.orc_make_impedance(nfreqs, fmax, R_bend, f_rev, gap, use_csr, s, xi, coll,
                                         z.reshape(-1))
        return z, bool(nz)

    def airy(self, x):
        v = [cd() for _ in range(4)]
        over = self.lib.orc_airy(x, *[C.byref(t) for t in v])
        return [t.value for t in v], bool(over)

    def derive(self, params):
        d = Derived()
        self.lib.orc_derive(C.byref(params), C.byref(d))
        return d

    def rf_modulation(self, steps, phasespread, amplspread, modampl, modtimeincrement, revolutionpart,
                      syncphase, gauss):
        """DynamicRFKickMap::__calcModulation with given normal deviates gauss[steps][2]"""
        out = np.zeros((steps, 2), f32)
        g = np.ascontiguousarray(gauss, f32).reshape(-1)
        self.lib.orc_rf_modulation(steps, phasespread, amplspread, modampl, modtimeincrement,
                                   revolutionpart, syncphase, g, out.reshape(-1))
        return out

    def sim(self, params, grid=None, particles=None, zwake=None, zrad=None):
        """zwake / zrad: impedance tables made earlier by make_impedance with the same parameters
        (tests/golden/impedance_tables.npz for the parallel-plates configs)"""
        return _Sim(self.lib, "orc", params, grid, particles, zwake, zrad)


class Ref:
    """the reference's own sources compiled in place (oracle/_ref)"""

    def __init__(self, fma=False):
        self.fma = fma
        self.lib = L = C.CDLL(lib_path("ref", fma))
        L.ref_last_error.restype = C.c_char_p
        L.ref_kickmap_apply.argtypes = [u32, u32, ci, _fp, _fp, _fp, P, u32]
        L.ref_fokker_planck_apply.argtypes = [u32, cf, cf, cf, u32, u32, P, P, P]
        L.ref_fokker_planck_track.argtypes = [u32, cf, cf, cf, u32, u32, u32, _fp, _fp, u32]
        L.ref_rotation_apply.argtypes = [u32, cf, cf, cf, cf, cf, u32, ci, u32, _fp, _fp, P, u32]
        L.ref_calc_coefficients.argtypes = [cf, u32, _fp]
        L.ref_phasespace_reduce.argtypes = [u32, cf, cf, cf, cf, _fp, P, P, P, P]
        L.ref_initial_grid.argtypes = [u32, cf, cf, cf, cf, cd, _fp]
        L.ref_normalize.argtypes = [u32, cf, cf, cf, cf, _fp, _fp]
        L.ref_wake_potential.argtypes = [u32, cf, cf, cf, cf, _fp, _fp, u32, cd, _fp, P, P, P]
        L.ref_csr_power.argtypes = [u32, cf, cf, cf, cf, cd, _fp, _fp, u32, cf, P, _fp]
        L.ref_train_wake_potential.argtypes = [u32, u32, cf, cf, cf, cf, _fp, _up, u32, _fp, u32, cd, _fp, P, P]
        L.ref_make_impedance.argtypes = [C.c_size_t, cf, cd, cd, cd, ci, cd, cd, cd, _fp,
                                         C.POINTER(ci)]
        L.ref_airy.argtypes = [cd] + [C.POINTER(cd)] * 4
        L.ref_sim_create.restype = P
        L.ref_sim_create.argtypes = [C.POINTER(RingParams), P, P, u32]
        L.ref_sim_destroy.argtypes = [P]
        L.ref_sim_step.argtypes = [P, u32]
        L.ref_sim_derived.argtypes = [P, C.POINTER(Derived)]
        L.ref_sim_get.argtypes = [P, P, P, P, P]
        L.ref_sim_tables.argtypes = [P, P, P, P, P, P]
        L.ref_sim_observe.argtypes = [P, _fp, P, P]
        L.ref_sim_units.argtypes = [P, C.POINTER(cd * 3)]
        L.ref_sim_set_rf_modulation.argtypes = [P, _fp, u32, ci]

    def _check(self, rc):
        if rc == 1:
            raise ValueError(self.lib.ref_last_error().decode())
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def coefficients(self, f, it):
        ic = np.zeros(4, f32)
        self._check(self.lib.ref_calc_coefficients(f, it, ic))
        return ic[:it]

    def kick_apply(self, grid, offset, it, axis, particles=None):
        n = grid.shape[0]
        out = np.empty_like(grid)
        p = None if particles is None else np.ascontiguousarray(particles, f32).copy()
        self._check(self.lib.ref_kickmap_apply(
            n, it, axis, np.ascontiguousarray(offset, f32), np.ascontiguousarray(grid, f32), out,
            _opt(p), 0 if p is None else len(p)))
        return out if particles is None else (out, p)

    def fp_weights5(self, n, pmin, pmax, e1, fptype=3, derivation=4):
        w = np.zeros((n, 5), f32)
        self._check(self.lib.ref_fokker_planck_apply(n, pmin, pmax, e1, fptype, derivation,
                                                     None, None, _opt(w)))
        return w

    def fp_apply(self, grid, pmin, pmax, e1, fptype=3, derivation=4):
        n = grid.shape[0]
        g = np.ascontiguousarray(grid, f32)
        out = np.empty_like(g)
        self._check(self.lib.ref_fokker_planck_apply(n, pmin, pmax, e1, fptype, derivation,
                                                     _opt(g), _opt(out), None))
        return out

    def fp_track(self, grid, pmin, pmax, e1, fptype, derivation, fptrack, particles):
        n = grid.shape[0]
        p = np.ascontiguousarray(particles, f32).copy()
        self._check(self.lib.ref_fokker_planck_track(
            n, pmin, pmax, e1, fptype, derivation, fptrack, np.ascontiguousarray(grid, f32), p,
            len(p)))
        return p

    def rotation_apply(self, grid, qmin, qmax, pmin, pmax, angle, it, clamp=0, rotmapsize=0,
                       particles=None):
        n = grid.shape[0]
        out = np.zeros_like(grid)
        p = None if particles is None else np.ascontiguousarray(particles, f32).copy()
        self._check(self.lib.ref_rotation_apply(
            n, qmin, qmax, pmin, pmax, angle, it, clamp, rotmapsize,
            np.ascontiguousarray(grid, f32), out, _opt(p), 0 if p is None else len(p)))
        return out if particles is None else (out, p)

    def reduce(self, grid, qmin, qmax, pmin, pmax):
        n = grid.shape[0]
        px, py = np.zeros(n, f32), np.zeros(n, f32)
        integ = np.zeros(1, f32)
        m = np.zeros(6, f32)
        self._check(self.lib.ref_phasespace_reduce(
            n, qmin, qmax, pmin, pmax, np.ascontiguousarray(grid, f32), _opt(px), _opt(py),
            _opt(integ), _opt(m)))
        return px, py, float(integ[0]), m

    def initial_grid(self, n, qmin, qmax, pmin, pmax, zoom=1.0):
        g = np.zeros((n, n), f32)
        self._check(self.lib.ref_initial_grid(n, qmin, qmax, pmin, pmax, zoom, g))
        return g

    def normalize(self, grid, qmin, qmax, pmin, pmax):
        out = np.empty_like(grid)
        self._check(self.lib.ref_normalize(grid.shape[0], qmin, qmax, pmin, pmax,
                                           np.ascontiguousarray(grid, f32), out))
        return out

    def wake_potential(self, grid, qmin, qmax, pmin, pmax, z, raw_scaling):
        """returns wake[n], formfactor[nmax/2+1,2], padded wake[nmax], scaling actually used"""
        n, nmax = grid.shape[0], z.reshape(-1, 2).shape[0]
        wake = np.zeros(n, f32)
        ff = np.zeros((nmax // 2 + 1, 2), f32)
        pw = np.zeros(nmax, f32)
        used = np.zeros(1, f32)
        self._check(self.lib.ref_wake_potential(
            n, qmin, qmax, pmin, pmax, np.ascontiguousarray(grid, f32),
            np.ascontiguousarray(z, f32).reshape(-1), nmax, raw_scaling, wake, _opt(ff), _opt(pw),
            _opt(used)))
        return wake, ff, pw, float(used[0])

    def train_wake_potential(self, grids, qmin, qmax, pmin, pmax, buckets, spacing_bins, z, raw_scaling):
        """ElectricField::wakePotential with PhaseSpace::nb = len(grids): wake [nb][n], padded wake [nmax],
        scaling actually used"""
        grids = np.ascontiguousarray(grids, f32)
        nb, n = grids.shape[0], grids.shape[1]
        nmax = z.reshape(-1, 2).shape[0]
        wake = np.zeros((nb, n), f32)
        pw = np.zeros(nmax, f32)
        used = np.zeros(1, f32)
        self._check(self.lib.ref_train_wake_potential(
            n, nb, qmin, qmax, pmin, pmax, grids.reshape(-1), np.ascontiguousarray(buckets, np.uint32), spacing_bins,
            np.ascontiguousarray(z, f32).reshape(-1), nmax, raw_scaling, wake.reshape(-1), _opt(pw), _opt(used)))
        return wake, pw, float(used[0])

    def csr_power(self, grid, qmin, qmax, pmin, pmax, bl, z, cutoff):
        n, nmax = grid.shape[0], z.reshape(-1, 2).shape[0]
        spec = np.zeros(nmax, f32)
        power = np.zeros(1, f32)
        self._check(self.lib.ref_csr_power(
            n, qmin, qmax, pmin, pmax, bl, np.ascontiguousarray(grid, f32),
            np.ascontiguousarray(z, f32).reshape(-1), nmax, cutoff, _opt(spec), power))
        return spec, float(power[0])

    def make_impedance(self, nfreqs, fmax, R_bend, f_rev, gap, use_csr=1, s=0.0, xi=0.0, coll=0.0):
        z = np.zeros((nfreqs, 2), f32)
        nz = ci(0)
        self._check(self.lib.ref_make_impedance(nfreqs, fmax, R_bend, f_rev, gap, use_csr, s, xi,
                                                coll, z.reshape(-1), C.byref(nz)))
        return z, bool(nz.value)

    def airy(self, x):
        v = [cd() for _ in range(4)]
        over = self.lib.ref_airy(x, *[C.byref(t) for t in v])
        return [t.value for t in v], bool(over)

    def sim(self, params, grid=None, particles=None):
        return _Sim(self.lib, "ref", params, grid, particles)


class _Sim:
    """whole-simulation handle with the same surface for Oracle and Ref"""

    def __init__(self, lib, prefix, params, grid, particles, zwake=None, zrad=None):
        self.lib, self.px = lib, prefix
        self.n = params.grid_size
        g = None if grid is None else np.ascontiguousarray(grid, f32)
        p = None if particles is None else np.ascontiguousarray(particles, f32)
        self.np = 0 if p is None else len(p)
        if zwake is not None or zrad is not None:
            assert prefix == "orc"
            zw = None if zwake is None else np.ascontiguousarray(zwake, f32)
            zr = None if zrad is None else np.ascontiguousarray(zrad, f32)
            self.h = lib.orc_sim_create_z(C.byref(params), _opt(g), _opt(p), self.np, _opt(zw), _opt(zr))
        else:
            self.h = getattr(lib, prefix + "_sim_create")(C.byref(params), _opt(g), _opt(p), self.np)
        if not self.h:
            raise RuntimeError("sim_create failed")
        self.derived = Derived()
        getattr(lib, prefix + "_sim_derived")(self.h, C.byref(self.derived))
        self.nmax = self.derived.padded_bins

    def step(self, nsteps=1):
        getattr(self.lib, self.px + "_sim_step")(self.h, nsteps)

    def get(self):
        n = self.n
        grid, prof, wake = np.zeros((n, n), f32), np.zeros(n, f32), np.zeros(n, f32)
        parts = np.zeros((self.np, 2), f32) if self.np else None
        getattr(self.lib, self.px + "_sim_get")(self.h, _opt(grid), _opt(prof), _opt(wake),
                                                 _opt(parts))
        return grid, prof, wake, parts

    def tables(self):
        n, nm = self.n, self.nmax
        rf, dr = np.zeros(n, f32), np.zeros(n, f32)
        fp5 = np.zeros((n, 5), f32)
        zw, zr = np.zeros((nm, 2), f32), np.zeros((nm, 2), f32)
        getattr(self.lib, self.px + "_sim_tables")(self.h, _opt(rf), _opt(dr), _opt(fp5),
                                                    _opt(zw), _opt(zr))
        return dict(rf=rf, drift=dr, fp5=fp5, zwake=zw, zrad=zr)

    def set_tables(self, rf=None, drift=None, fp5=None, zwake=None):
        assert self.px == "orc"
        a = [None if t is None else np.ascontiguousarray(t, f32) for t in (rf, drift, fp5, zwake)]
        self.lib.orc_sim_set_tables(self.h, *[_opt(t) for t in a])

    def observe(self):
        obs = np.zeros(8, f32)
        spec = np.zeros(self.nmax, f32)
        projy = np.zeros(self.n, f32)
        getattr(self.lib, self.px + "_sim_observe")(self.h, obs, _opt(spec), _opt(projy))
        return obs, spec, projy

    def set_rf_modulation(self, phase_ampl, interpolation=4):
        """DynamicRFKickMap with the given queue of (phase, amplitude) pairs, one per step"""
        seq = np.ascontiguousarray(phase_ampl, f32).reshape(-1, 2)
        if self.px == "orc":
            self.lib.orc_sim_set_rf_modulation(self.h, seq.reshape(-1), len(seq))
        else:
            rc = self.lib.ref_sim_set_rf_modulation(self.h, seq.reshape(-1), len(seq), interpolation)
            if rc:
                raise RuntimeError(self.lib.ref_last_error().decode())

    def units(self):
        """(Volt, WattPerHertz, Watt) factors of the reference's ElectricField objects"""
        assert self.px == "ref"
        out = (cd * 3)()
        self.lib.ref_sim_units(self.h, C.byref(out))
        return tuple(out)

    def close(self):
        if self.h:
            getattr(self.lib, self.px + "_sim_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
