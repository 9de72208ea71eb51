"""ctypes binding of the C ABI (include/ivs.h, include/ivs_host.h) exported by lib/libivs.so.

This is the reference-facing call surface used by the tests and the benchmark: HOST numpy
buffers go in, the CUDA handler does the work, HOST buffers come out -- the same shape as the
reference's `if (_oclh) enqueue... else cpu-loop` dispatch, minus the CPU branch: there is no
CPU fallback here.  If libivs.so is missing or no B200 is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

from .params import Derived, RfModel, RingParams, SimDesc

HERE = os.path.dirname(os.path.abspath(__file__))
# IVS_LIB: another build of the same library (A/B timing of kernel variants, tools/ab.sh)
LIB_PATH = os.environ.get("IVS_LIB") or os.path.join(HERE, "lib", "libivs.so")
f32 = np.float32

IVS_OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_UNSUPPORTED, ERR_HALO, ERR_NO_DEVICE = 1, 2, 3, 4, 5
ARITH_FMA, ARITH_STRICT = 0, 1
SIM_UNFUSED, SIM_NO_GRAPH, SIM_TILED, SIM_QPT1, SIM_QPT2, SIM_STREAM8, SIM_TRACK_INLINE, SIM_NO_PERSISTENT = 2, 4, 8, 16, 32, 64, 128, 256

P, u32, u64, cf, cd, ci, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_double, C.c_int, C.c_size_t
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

# name -> (restype, argtypes); every symbol include/*.h declares
SIGNATURES = {
    "ivs_ctx_create": (ci, [ci, ci, C.POINTER(P)]),
    "ivs_ctx_destroy": (ci, [P]),
    "ivs_last_error": (C.c_char_p, [P]),
    "ivs_device_count": (ci, []),
    "ivs_version": (C.c_char_p, []),
    "ivs_ctx_stream": (P, [P]),
    "ivs_ctx_launch_count": (u64, [P]),
    "ivs_malloc": (ci, [P, sz, C.POINTER(P)]),
    "ivs_free": (ci, [P, P]),
    "ivs_upload": (ci, [P, P, P, sz]),
    "ivs_download": (ci, [P, P, P, sz]),
    "ivs_copy": (ci, [P, P, P, sz]),
    "ivs_memset_zero": (ci, [P, P, sz]),
    "ivs_host_alloc": (ci, [P, sz, C.POINTER(P)]),
    "ivs_host_free": (ci, [P, P]),
    "ivs_flush": (ci, [P]),
    "ivs_sync": (ci, [P]),
    "ivs_kick_y": (ci, [P, P, P, u32, P, u32]),
    "ivs_kick_x": (ci, [P, P, P, u32, P, u32]),
    "ivs_fokker_planck": (ci, [P, P, P, u32, P]),
    "ivs_fokker_planck_weights": (ci, [u32, cf, cf, cf, u32, u32, _fp]),
    "ivs_rotate": (ci, [P, P, P, u32, cf, cf, cf, cf, cf, u32, ci]),
    "ivs_apply_source_map": (ci, [P, P, P, u32, P, u32]),
    "ivs_identity": (ci, [P, P, P, u32]),
    "ivs_simpson_weights": (ci, [u32, cf, _fp]),
    "ivs_project_x": (ci, [P, P, u32, P, P]),
    "ivs_project_y": (ci, [P, P, u32, P, P]),
    "ivs_integrate": (ci, [P, P, u32, P, P]),
    "ivs_normalize": (ci, [P, P, u32, cf, P]),
    "ivs_moments": (ci, [P, P, u32, cf, cf, P, P]),
    "ivs_wake_potential": (ci, [P, P, u32, P, u32, cf, P, P]),
    "ivs_wake_function": (ci, [P, P, u32, cf, u32, P]),
    "ivs_train_wake_potential": (ci, [P, P, u32, u32, P, cf, P]),
    "ivs_csr_power": (ci, [P, P, u32, P, u32, cf, cf, cf, P, P]),
    "ivs_track_kick": (ci, [P, P, u32, P, u32, ci]),
    "ivs_track_fokker_planck": (ci, [P, P, u32, P, u32, P, u32, cf, cf, u64, u64]),
    "ivs_track_rotate": (ci, [P, P, u32, cf, cf, cf]),
    "ivs_sim_create": (ci, [P, C.POINTER(SimDesc), C.POINTER(P)]),
    "ivs_sim_destroy": (ci, [P]),
    "ivs_sim_step": (ci, [P, u32]),
    "ivs_sim_steps_done": (u64, [P]),
    "ivs_sim_fallback_steps": (u64, [P]),
    "ivs_sim_state": (ci, [P, C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(P)]),
    "ivs_sim_download": (ci, [P, P, P, P, P]),
    "ivs_sim_upload_grid": (ci, [P, _fp]),
    "ivs_sim_observe": (ci, [P, _fp]),
    "ivs_sim_set_radiation": (ci, [P, _fp, cf, cf]),
    "ivs_sim_fields": (ci, [P, P, P]),
    "ivs_sim_renormalize": (ci, [P]),
    "ivs_sim_set_rf_offset": (ci, [P, _fp]),
    "ivs_sim_set_rf_modulation": (ci, [P, C.POINTER(RfModel), P, u32]),
    "ivs_sim_set_bunch_train": (ci, [P, _up, u32, _fp, u32, cf]),
    "ivs_rf_model_init": (ci, [C.POINTER(RingParams), C.POINTER(Derived), C.POINTER(RfModel)]),
    "ivs_rf_modulation": (ci, [u64, u32, cf, cf, cf, cd, cd, cf, P, _fp]),
    # include/ivs_host.h
    "ivs_default_params": (None, [C.POINTER(RingParams)]),
    "ivs_derive": (ci, [C.POINTER(RingParams), C.POINTER(Derived)]),
    "ivs_rf_kick_offsets": (ci, [C.POINTER(RingParams), C.POINTER(Derived), _fp]),
    "ivs_drift_offsets": (ci, [C.POINTER(RingParams), C.POINTER(Derived), _fp]),
    "ivs_rf_kick_linear": (ci, [u32, cf, cf, cf, cf, cf, _fp]),
    "ivs_rf_kick_sine": (ci, [u32, cf, cf, cf, cf, cf, cf, cf, cf, cf, cf, _fp]),
    "ivs_drift_slip": (ci, [u32, cf, cf, cf, cf, cf, _fp, u32, cf, _fp]),
    "ivs_make_impedance": (ci, [sz, cf, cd, cd, cd, ci, cd, cd, cd, _fp, C.POINTER(ci)]),
    "ivs_wake_impedance": (ci, [C.POINTER(RingParams), C.POINTER(Derived), _fp, C.POINTER(ci)]),
    "ivs_radiation_impedance": (ci, [C.POINTER(RingParams), C.POINTER(Derived), _fp]),
    "ivs_read_impedance_file": (ci, [C.c_char_p, P, sz, C.POINTER(sz)]),
    "ivs_add_impedance_file": (ci, [C.c_char_p, _fp, sz]),
    "ivs_initial_grid": (ci, [C.POINTER(Derived), u32, cd, _fp]),
    "ivs_gaussian_projection": (ci, [u32, cf, cf, cd, _fp]),
    "ivs_haissinski_grid": (ci, [C.POINTER(Derived), u32, _fp, P, u32, cf, u32, cd, _fp, C.POINTER(cd)]),
    "ivs_train_wake_windows": (ci, [_fp, u32, u32, u32, _up, u32, _dp]),
    "ivs_airy": (ci, [cd] + [C.POINTER(cd)] * 4),
}

_lib = None


class IvsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ivs error %d: %s" % (code, msg))
        self.code = code


def load():
    """dlopen lib/libivs.so and bind every declared symbol; raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not found: run `python -m inovesa_b200.build` (there is no "
                              "CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(P)


# ---------------------------------------------------------------------------- host helpers

def derive(params):
    d = Derived()
    rc = load().ivs_derive(C.byref(params), C.byref(d))
    if rc:
        raise IvsError(rc, "ivs_derive")
    return d


def rf_kick_offsets(params, d):
    out = np.zeros(params.grid_size, f32)
    load().ivs_rf_kick_offsets(C.byref(params), C.byref(d), out)
    return out


def drift_offsets(params, d):
    out = np.zeros(params.grid_size, f32)
    load().ivs_drift_offsets(C.byref(params), C.byref(d), out)
    return out


def fokker_planck_weights(n, pmin, pmax, e1, fptype=3, derivation=4):
    w = np.zeros((n, 5), f32)
    rc = load().ivs_fokker_planck_weights(n, pmin, pmax, e1, fptype, derivation, w.reshape(-1))
    if rc:
        raise IvsError(rc, "ivs_fokker_planck_weights")
    return w


def simpson_weights(n, delta):
    ws = np.zeros(n, f32)
    load().ivs_simpson_weights(n, delta, ws)
    return ws


def make_impedance(nfreqs, fmax, R_bend, f_rev, gap, use_csr=1, s=0.0, xi=0.0, coll=0.0):
    z = np.zeros((nfreqs, 2), f32)
    nz = ci(0)
    load().ivs_make_impedance(nfreqs, fmax, R_bend, f_rev, gap, use_csr, s, xi, coll,
                              z.reshape(-1), C.byref(nz))
    return z, bool(nz.value)


def wake_impedance(params, d, impedance_file=None, nfreqs=None):
    """makeImpedance for the wake (src/main.cpp:717-721, src/Z/ImpedanceFactory.cpp:21-84); impedance_file:
    a table of `line-number real imag` records added on top (:70-75); nfreqs: the table's length if not the
    padded profile's (the spaced length of a train of bunches, :718)"""
    nf = int(d.padded_bins if nfreqs is None else nfreqs)
    z = np.zeros((nf, 2), f32)
    nz = ci(0)
    if nfreqs is None:
        load().ivs_wake_impedance(C.byref(params), C.byref(d), z.reshape(-1), C.byref(nz))
    else:
        load().ivs_make_impedance(nf, d.fmax, d.R_bend, d.f_rev, params.gap, 1 if params.use_csr else 0,
                                  params.wall_conductivity, params.wall_susceptibility, params.collimator,
                                  z.reshape(-1), C.byref(nz))
    if impedance_file:
        rc = load().ivs_add_impedance_file(str(impedance_file).encode(), z.reshape(-1), nf)
        if rc:
            raise IvsError(rc, "impedance file %s: cannot be read or has fewer than %d records" % (impedance_file, nf))
        return z, True
    return z, bool(nz.value)


def read_impedance_file(path, cap=1 << 20):
    """Impedance::readData (src/Z/Impedance.cpp:94-111): complex64 array"""
    z = np.zeros((cap, 2), f32)
    n = sz(0)
    rc = load().ivs_read_impedance_file(str(path).encode(), z.ctypes.data_as(P), cap, C.byref(n))
    if rc:
        raise IvsError(rc, "impedance file %s cannot be opened" % path)
    return z[: min(n.value, cap)].copy().view(np.complex64).reshape(-1)


def radiation_impedance(params, d):
    z = np.zeros((d.padded_bins, 2), f32)
    load().ivs_radiation_impedance(C.byref(params), C.byref(d), z.reshape(-1))
    return z


def initial_grid(d, n, zoom=1.0):
    g = np.zeros((n, n), f32)
    load().ivs_initial_grid(C.byref(d), n, zoom, g.reshape(-1))
    return g


def haissinski_grid(params, d, zwake=None, wakescaling=None, iterations=2000, relax=0.1, rf=None):
    """stationary start distribution of the step map (see include/ivs_host.h); returns (grid, residual)"""
    n = params.grid_size
    rf = rf_kick_offsets(params, d) if rf is None else np.ascontiguousarray(rf, f32)
    if zwake is None:
        zwake, nz = wake_impedance(params, d)
        zwake = zwake if nz else None
    zw = None if zwake is None else np.ascontiguousarray(zwake, f32).reshape(-1)
    g = np.zeros((n, n), f32)
    res = cd(0)
    rc = load().ivs_haissinski_grid(C.byref(d), n, rf, _ptr(zw), d.padded_bins,
                                    d.wakescaling if wakescaling is None else wakescaling,
                                    iterations, relax, g.reshape(-1), C.byref(res))
    if rc not in (0, 1):
        raise IvsError(rc, "ivs_haissinski_grid")
    return g, res.value


def train_layout(params, filling, spacing_bins=None):
    """The filling pattern of src/main.cpp:248-321: `filling` = one current per bucket (0 = empty).
    Returns dict(buckets, fractions, Ib, spacing_bins, spaced_bins, wakescaling, params, d): bucket numbers
    counted from the END of the pattern (:258-268), the bunches' charge fractions, the derived scalars for
    the whole beam (bunch_current = Ib), the bunch spacing in bins (:304-307), the padded length of the
    train (:318-321) and ElectricField::_wakescaling for an impedance of that length.  spacing_bins overrides
    the bunch spacing (1 / f_RF in the reference, :215): the padded length follows as buckets * spacing."""
    filling = [float(c) for c in filling]
    buckets = [len(filling) - 1 - i for i, c in enumerate(filling) if c > 0]
    bunches = [c for c in filling if c > 0]
    if not bunches:
        raise ValueError("empty filling pattern")
    Ib = float(sum(bunches))
    p2 = RingParams.from_buffer_copy(params)
    p2.bunch_current = Ib
    d = derive(p2)
    n = params.grid_size
    spacing_ps = (1.0 / d.f_RF) * 2.99792458e8 / d.bl / params.pq_size
    spaced = int(np.ceil(n * len(filling) * spacing_ps))
    spacing = int(d.spacing_bins)
    if spacing_bins is not None:
        spacing = int(spacing_bins)
        spaced = max(len(filling) * spacing, max(buckets) * spacing + n)
    if params.round_padding or spacing_bins is not None:
        spaced = 1 << max(spaced - 1, 1).bit_length()
    raw = f32(d.wakescaling) * f32(d.padded_bins)  # exact for the rounded (power-of-two) padding
    return dict(buckets=buckets, fractions=[c / Ib for c in bunches], Ib=Ib, spacing_bins=spacing,
                spaced_bins=spaced, wakescaling=float(f32(raw) / f32(spaced)), params=p2, d=d)


def train_wake_windows(z, n, buckets, spacing_bins):
    """wake-function windows [nb][nb][2 n] (double) of a train of bunches (ivs_train_wake_windows)"""
    z = np.ascontiguousarray(z, f32).reshape(-1, 2)
    b = np.ascontiguousarray(buckets, np.uint32)
    w = np.zeros((len(b), len(b), 2 * n), np.float64)
    rc = load().ivs_train_wake_windows(z.reshape(-1), z.shape[0], n, len(b), b, spacing_bins, w.reshape(-1))
    if rc != 0:
        raise IvsError(rc, "ivs_train_wake_windows: " + ("a bunch does not fit the padded profile" if rc == -2 else "bad arguments"))
    return w


def rf_model(params, d):
    m = RfModel()
    rc = load().ivs_rf_model_init(C.byref(params), C.byref(d), C.byref(m))
    if rc:
        raise IvsError(rc, "ivs_rf_model_init")
    return m


def rf_modulation(steps, phasespread, amplspread, modampl, modtimeincrement, revolutionpart, syncphase,
                  seed=0, gauss=None):
    """DynamicRFKickMap::__calcModulation: (phase, amplitude) per step, [steps][2]"""
    out = np.zeros((steps, 2), f32)
    g = None if gauss is None else np.ascontiguousarray(gauss, f32).reshape(-1)
    rc = load().ivs_rf_modulation(seed, steps, phasespread, amplspread, modampl, modtimeincrement,
                                  revolutionpart, syncphase, _ptr(g), out.reshape(-1))
    if rc:
        raise IvsError(rc, "ivs_rf_modulation")
    return out


def airy(x):
    v = [cd() for _ in range(4)]
    over = load().ivs_airy(x, *[C.byref(t) for t in v])
    return [t.value for t in v], bool(over)


# ---------------------------------------------------------------------------- device context

class Context:
    """ivs_ctx: one CUDA device + one in-order stream (the reference's OCLH)."""

    def __init__(self, device=0, arith=ARITH_FMA):
        self.lib = load()
        h = P()
        rc = self.lib.ivs_ctx_create(device, arith, C.byref(h))
        if rc:
            raise IvsError(rc, self.lib.ivs_last_error(None).decode())
        self.h = h
        self.arith = arith
        self._bufs = []

    def check(self, rc):
        if rc:
            raise IvsError(rc, self.lib.ivs_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            for b in self._bufs:
                self.lib.ivs_free(self.h, b)
            self._bufs = []
            self.lib.ivs_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # memory ---------------------------------------------------------------------------
    def malloc(self, nbytes):
        p = P()
        self.check(self.lib.ivs_malloc(self.h, nbytes, C.byref(p)))
        self._bufs.append(p)
        return p

    def free(self, p):
        self._bufs = [b for b in self._bufs if b.value != p.value]
        self.check(self.lib.ivs_free(self.h, p))

    def to_device(self, a):
        a = np.ascontiguousarray(a)
        p = self.malloc(a.nbytes)
        self.check(self.lib.ivs_upload(self.h, p, _ptr(a), a.nbytes))
        self.sync()
        return p

    def to_host(self, p, shape, dtype=np.float32):
        out = np.empty(shape, dtype)
        self.check(self.lib.ivs_download(self.h, _ptr(out), p, out.nbytes))
        return out

    def sync(self):
        self.check(self.lib.ivs_sync(self.h))

    @property
    def launches(self):
        return int(self.lib.ivs_ctx_launch_count(self.h))

    # host-buffer convenience wrappers of the unfused ops (upload -> kernel -> download) ------
    def _map(self, fn, grid, *extra):
        grid = np.ascontiguousarray(grid, f32)
        n = grid.shape[-1]
        d_in = self.to_device(grid)
        d_out = self.malloc(grid.nbytes)
        try:
            self.check(fn(self.h, d_in, d_out, n, *extra))
            return self.to_host(d_out, grid.shape)
        finally:
            self.free(d_in)
            self.free(d_out)

    def kick_y(self, grid, offset, it=4):
        d_off = self.to_device(np.ascontiguousarray(offset, f32))
        try:
            return self._map(self.lib.ivs_kick_y, grid, d_off, it)
        finally:
            self.free(d_off)

    def kick_x(self, grid, offset, it=4):
        d_off = self.to_device(np.ascontiguousarray(offset, f32))
        try:
            return self._map(self.lib.ivs_kick_x, grid, d_off, it)
        finally:
            self.free(d_off)

    def fokker_planck(self, grid, w5):
        d_w = self.to_device(np.ascontiguousarray(w5, f32))
        try:
            return self._map(self.lib.ivs_fokker_planck, grid, d_w)
        finally:
            self.free(d_w)

    def rotate(self, grid, angle, qmin, qmax, pmin, pmax, it=4, clamp=0):
        return self._map(self.lib.ivs_rotate, grid, angle, qmin, qmax, pmin, pmax, it, clamp)

    def apply_source_map(self, grid, index, weight):
        grid = np.ascontiguousarray(grid, f32)
        ncells, ip = index.shape
        table = np.zeros((ncells, ip), dtype=[("index", np.uint32), ("weight", np.float32)])
        table["index"], table["weight"] = index, weight
        d_in, d_t = self.to_device(grid), self.to_device(table)
        d_out = self.malloc(grid.nbytes)
        try:
            self.check(self.lib.ivs_apply_source_map(self.h, d_in, d_out, ncells, d_t, ip))
            return self.to_host(d_out, grid.shape)
        finally:
            for p in (d_in, d_t, d_out):
                self.free(p)

    def reduce(self, grid, qmin, qmax, pmin, pmax):
        """X/Y projections, integral and moments of one grid like PhaseSpace does
        (updateXProjection, integrate, variance(0), updateYProjection, variance(1))"""
        grid = np.ascontiguousarray(grid, f32)
        n = grid.shape[0]
        dq = f32((f32(qmax) - f32(qmin)) / f32(n - 1))
        dp = f32((f32(pmax) - f32(pmin)) / f32(n - 1))
        d_g = self.to_device(grid)
        d_ws = self.to_device(simpson_weights(n, dq))
        d_px, d_py = self.malloc(4 * n), self.malloc(4 * n)
        d_int, d_m = self.malloc(4), self.malloc(4 * 6)
        try:
            L, h = self.lib, self.h
            self.check(L.ivs_project_x(h, d_g, n, d_ws, d_px))
            self.check(L.ivs_integrate(h, d_px, n, d_ws, d_int))
            self.check(L.ivs_moments(h, d_px, n, qmin, dq, d_int, d_m))
            self.check(L.ivs_project_y(h, d_g, n, d_ws, d_py))
            self.check(L.ivs_moments(h, d_py, n, pmin, dp, d_int, P(d_m.value + 12)))
            return (self.to_host(d_px, n), self.to_host(d_py, n), float(self.to_host(d_int, 1)[0]),
                    self.to_host(d_m, 6))
        finally:
            for p in (d_g, d_ws, d_px, d_py, d_int, d_m):
                self.free(p)

    def normalize(self, grid, qmin, qmax):
        grid = np.ascontiguousarray(grid, f32)
        n = grid.shape[0]
        dq = f32((f32(qmax) - f32(qmin)) / f32(n - 1))
        d_g, d_ws = self.to_device(grid), self.to_device(simpson_weights(n, dq))
        d_px, d_int = self.malloc(4 * n), self.malloc(4)
        try:
            L, h = self.lib, self.h
            self.check(L.ivs_project_x(h, d_g, n, d_ws, d_px))
            self.check(L.ivs_integrate(h, d_px, n, d_ws, d_int))
            self.check(L.ivs_normalize(h, d_g, n, 1.0, d_int))
            return self.to_host(d_g, grid.shape)
        finally:
            for p in (d_g, d_ws, d_px, d_int):
                self.free(p)

    def wake_potential(self, profile, z, wakescaling, want_formfactor=False):
        profile = np.ascontiguousarray(profile, f32)
        z = np.ascontiguousarray(z, f32).reshape(-1, 2)
        n, nmax = len(profile), z.shape[0]
        d_p, d_z = self.to_device(profile), self.to_device(z)
        d_w = self.malloc(4 * n)
        d_f = self.malloc(8 * (nmax // 2 + 1)) if want_formfactor else None
        try:
            self.check(self.lib.ivs_wake_potential(self.h, d_p, n, d_z, nmax, wakescaling, d_w, d_f))
            w = self.to_host(d_w, n)
            if want_formfactor:
                return w, self.to_host(d_f, (nmax // 2 + 1, 2))
            return w
        finally:
            for p in (d_p, d_z, d_w, d_f):
                if p is not None:
                    self.free(p)

    def train_wake_potential(self, profiles, windows, wakescaling):
        """wake potentials [nb][n] of a train of bunches from their profiles [nb][n] and the windows of
        train_wake_windows"""
        profiles = np.ascontiguousarray(profiles, f32)
        windows = np.ascontiguousarray(windows, np.float64)
        nb, n = profiles.shape
        assert windows.shape == (nb, nb, 2 * n)
        d_p, d_w = self.to_device(profiles), self.to_device(windows)
        d_o = self.malloc(4 * nb * n)
        try:
            self.check(self.lib.ivs_train_wake_potential(self.h, d_p, n, nb, d_w, wakescaling, d_o))
            return self.to_host(d_o, (nb, n))
        finally:
            for p in (d_p, d_w, d_o):
                self.free(p)

    def wake_function(self, z, wakescaling, n):
        z = np.ascontiguousarray(z, f32).reshape(-1, 2)
        nmax = z.shape[0]
        d_z = self.to_device(z)
        d_w = self.malloc(4 * (2 * n - 1))
        try:
            self.check(self.lib.ivs_wake_function(self.h, d_z, nmax, wakescaling, n, d_w))
            return self.to_host(d_w, 2 * n - 1)
        finally:
            self.free(d_z)
            self.free(d_w)

    def csr_power(self, profile, z, dq, f_scale, cutoff):
        profile = np.ascontiguousarray(profile, f32)
        z = np.ascontiguousarray(z, f32).reshape(-1, 2)
        n, nmax = len(profile), z.shape[0]
        d_p, d_z = self.to_device(profile), self.to_device(z)
        d_s, d_pw = self.malloc(4 * nmax), self.malloc(4)
        try:
            self.check(self.lib.ivs_csr_power(self.h, d_p, n, d_z, nmax, dq, f_scale, cutoff, d_s, d_pw))
            return self.to_host(d_s, nmax), float(self.to_host(d_pw, 1)[0])
        finally:
            for p in (d_p, d_z, d_s, d_pw):
                self.free(p)

    def _track(self, particles, call):
        parts = np.ascontiguousarray(particles, f32)
        d_p = self.to_device(parts)
        try:
            call(d_p, len(parts))
            return self.to_host(d_p, parts.shape)
        finally:
            self.free(d_p)

    def track_kick(self, particles, offset, n, axis):
        d_off = self.to_device(np.ascontiguousarray(offset, f32))
        try:
            return self._track(particles, lambda d, k: self.check(
                self.lib.ivs_track_kick(self.h, d, k, d_off, n, axis)))
        finally:
            self.free(d_off)

    def track_fokker_planck(self, particles, grid, w5, fptrack, e1=0.0, dp=1.0, seed=0, step=0):
        grid = np.ascontiguousarray(grid, f32)
        n = grid.shape[0]
        d_g, d_w = self.to_device(grid), self.to_device(np.ascontiguousarray(w5, f32))
        try:
            return self._track(particles, lambda d, k: self.check(
                self.lib.ivs_track_fokker_planck(self.h, d, k, d_g, n, d_w, fptrack, e1, dp, seed, step)))
        finally:
            self.free(d_g)
            self.free(d_w)

    def track_rotate(self, particles, angle, zb0, zb1):
        return self._track(particles, lambda d, k: self.check(
            self.lib.ivs_track_rotate(self.h, d, k, angle, zb0, zb1)))


# ---------------------------------------------------------------------------- simulation

class Simulation:
    """ivs_sim: `members` independent single-bunch simulations advanced in lock-step.

    Built either from ring parameters (tables derived with the host helpers, which restate
    the reference's set-up code) or from explicit tables (`tables=dict(rf=, drift=, fp5=,
    zwake=, zrad=)`, e.g. taken from the oracle to isolate the device arithmetic).
    """

    def __init__(self, ctx, params, members=1, grid=None, particles=None, flags=0, tables=None,
                 bunch_currents=None, seed=0, strips=0, impedance_file=None):
        self.ctx, self.lib = ctx, ctx.lib
        self.params = params
        self.d = d = derive(params)
        n = self.n = params.grid_size
        self.members = members
        t = dict(tables or {})
        rf = np.ascontiguousarray(t["rf"] if "rf" in t else rf_kick_offsets(params, d), f32)
        drift = np.ascontiguousarray(t["drift"] if "drift" in t else drift_offsets(params, d), f32)
        if "fp5" in t:
            fp5 = None if t["fp5"] is None else np.ascontiguousarray(t["fp5"], f32)
        elif d.e1 > 0:
            fp5 = fokker_planck_weights(n, d.pmin, d.pmax, d.e1, params.fp_type, params.derivation)
        else:
            fp5 = None
        if "zwake" in t:
            zw = None if t["zwake"] is None else np.ascontiguousarray(t["zwake"], f32)
        else:
            zw, nz = wake_impedance(params, d, impedance_file)
            zw = zw if nz else None
        self.nmax = d.padded_bins
        if grid is None:
            grid = initial_grid(d, n, params.zoom)
        grid = np.ascontiguousarray(grid, f32)
        if grid.ndim == 2:
            grid = np.ascontiguousarray(np.broadcast_to(grid, (members, n, n)))
        assert grid.shape == (members, n, n)
        # ElectricField::_wakescaling per member: only the bunch current differs in a scan
        if bunch_currents is None:
            scal = np.full(members, d.wakescaling, f32)
        else:
            scal = np.zeros(members, f32)
            for i, ib in enumerate(bunch_currents):
                p2 = RingParams.from_buffer_copy(params)
                p2.bunch_current = ib
                scal[i] = derive(p2).wakescaling
        self.wakescaling = scal
        self.np = 0
        parts = None
        if particles is not None:
            parts = np.ascontiguousarray(particles, f32)
            if parts.ndim == 2:
                parts = np.ascontiguousarray(np.broadcast_to(parts, (members,) + parts.shape))
            self.np = parts.shape[1]
        desc = SimDesc(n=n, members=members, it=params.interpolation,
                       nmax=self.nmax, n_particles=self.np,
                       fptrack=params.fp_track, flags=flags, strips=strips, dq=d.dq, dp=d.dp, e1=d.e1,
                       qmin=d.qmin, pmin=d.pmin, reserved_f=0.0,
                       rf_offset=_ptr(rf), drift_offset=_ptr(drift), fp_weights=_ptr(fp5),
                       impedance=_ptr(zw), wakescaling=_ptr(scal), grid=_ptr(grid),
                       particles=_ptr(parts), seed=seed)
        h = P()
        ctx.check(self.lib.ivs_sim_create(ctx.h, C.byref(desc), C.byref(h)))
        self.h = h
        self.tables = dict(rf=rf, drift=drift, fp5=fp5, zwake=zw)

    def set_radiation(self, zrad=None, cutoff=None):
        zrad = radiation_impedance(self.params, self.d) if zrad is None else zrad
        f_scale = f32(2.99792458e8 / float(self.d.bl_f))
        cutoff = self.params.f_cutoff if cutoff is None else cutoff
        self.ctx.check(self.lib.ivs_sim_set_radiation(
            self.h, np.ascontiguousarray(zrad, f32).reshape(-1), f_scale, cutoff))

    def step(self, nsteps=1):
        self.ctx.check(self.lib.ivs_sim_step(self.h, nsteps))

    def download(self, grid=True, profile=True, wake=True, particles=True):
        n, m = self.n, self.members
        g = np.empty((m, n, n), f32) if grid else None
        pr = np.empty((m, n), f32) if profile else None
        w = np.empty((m, n), f32) if wake else None
        pt = np.empty((m, self.np, 2), f32) if (particles and self.np) else None
        self.ctx.check(self.lib.ivs_sim_download(self.h, _ptr(g), _ptr(pr), _ptr(w), _ptr(pt)))
        return g, pr, w, pt

    def upload_grid(self, grid):
        grid = np.ascontiguousarray(grid, f32).reshape(-1)
        self.ctx.check(self.lib.ivs_sim_upload_grid(self.h, grid))

    def observe(self):
        obs = np.zeros((self.members, 8), f32)
        self.ctx.check(self.lib.ivs_sim_observe(self.h, obs.reshape(-1)))
        return obs

    def fields(self, spectrum=True, wake=True):
        """CSR spectrum [members][nmax] and wake potential [members][n] of the current profile"""
        sp = np.zeros((self.members, self.nmax), f32) if spectrum else None
        w = np.zeros((self.members, self.n), f32) if wake else None
        self.ctx.check(self.lib.ivs_sim_fields(self.h, _ptr(sp), _ptr(w)))
        return sp, w

    def renormalize(self):
        self.ctx.check(self.lib.ivs_sim_renormalize(self.h))

    def set_rf_offset(self, rf):
        """new RF kick offsets for the following steps (the DynamicRFKickMap hook)"""
        rf = np.ascontiguousarray(rf, f32)
        assert rf.shape == (self.n,)
        self.ctx.check(self.lib.ivs_sim_set_rf_offset(self.h, rf))

    def set_bunch_train(self, buckets, spacing_bins, z, wakescaling):
        """the members become the bunches of one train (PhaseSpace::nb = members): z = the wake impedance at
        the spaced length, wakescaling = the whole beam's (train_layout); needs flags=SIM_UNFUSED"""
        b = np.ascontiguousarray(buckets, np.uint32)
        assert b.shape == (self.members,)
        z = np.ascontiguousarray(z, f32).reshape(-1, 2)
        self.ctx.check(self.lib.ivs_sim_set_bunch_train(self.h, b, spacing_bins, z.reshape(-1), z.shape[0], wakescaling))

    def set_rf_modulation(self, phase_ampl, model=None):
        """make the RF map a DynamicRFKickMap with this queue of (phase, amplitude) pairs; None: static again"""
        if phase_ampl is None:
            self.ctx.check(self.lib.ivs_sim_set_rf_modulation(self.h, None, None, 0))
            return
        seq = np.ascontiguousarray(phase_ampl, f32).reshape(-1, 2)
        m = rf_model(self.params, self.d) if model is None else model
        self.ctx.check(self.lib.ivs_sim_set_rf_modulation(self.h, C.byref(m), _ptr(seq), len(seq)))

    @property
    def steps_done(self):
        return int(self.lib.ivs_sim_steps_done(self.h))

    @property
    def fallback_steps(self):
        """steps redone through the kernel sequence because a wake kick left the fused kernel's pad"""
        return int(self.lib.ivs_sim_fallback_steps(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.ivs_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
