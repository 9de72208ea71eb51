"""ctypes mirrors of the C structs in include/ivs_host.h and include/ivs.h.

Field order and types must match the headers exactly (checked by tests/test_capi_layout.py
against sizeof() values the library reports).
"""
import ctypes as C


class RingParams(C.Structure):
    """ivs_ring_params (include/ivs_host.h): the reference's program options that reach the
    hot path, /root/reference/src/IO/ProgramOptions.cpp:22-305."""

    _fields_ = [
        ("grid_size", C.c_uint32),
        ("steps_per_Ts", C.c_uint32),
        ("interpolation", C.c_uint32),
        ("derivation", C.c_uint32),
        ("fp_type", C.c_uint32),
        ("fp_track", C.c_uint32),
        ("linear_rf", C.c_uint32),
        ("round_padding", C.c_uint32),
        ("use_csr", C.c_uint32),
        ("reserved", C.c_uint32),
        ("pq_size", C.c_float),
        ("shift_x", C.c_float),
        ("shift_y", C.c_float),
        ("alpha0", C.c_float),
        ("alpha1", C.c_float),
        ("alpha2", C.c_float),
        ("f_s", C.c_float),
        ("f_rev", C.c_float),
        ("harmonic", C.c_float),
        ("f_cutoff", C.c_float),
        ("padding", C.c_double),
        ("steps_per_Trev", C.c_double),
        ("E0", C.c_double),
        ("sE", C.c_double),
        ("V_RF", C.c_double),
        ("R_bend", C.c_double),
        ("t_damp", C.c_double),
        ("gap", C.c_double),
        ("collimator", C.c_double),
        ("wall_conductivity", C.c_double),
        ("wall_susceptibility", C.c_double),
        ("zoom", C.c_double),
        ("bunch_current", C.c_float),
        ("reserved_f", C.c_float),
    ]


class Derived(C.Structure):
    """ivs_derived (include/ivs_host.h): scalars of /root/reference/src/main.cpp:163-355."""

    _fields_ = [
        ("qmin", C.c_float), ("qmax", C.c_float), ("pmin", C.c_float), ("pmax", C.c_float),
        ("dq", C.c_float), ("dp", C.c_float),
        ("zerobin_x", C.c_float), ("zerobin_y", C.c_float),
        ("angle", C.c_float), ("fmax", C.c_float), ("e1", C.c_float),
        ("slip", C.c_float * 3),
        ("wakescaling", C.c_float),
        ("bl_f", C.c_float), ("dE_f", C.c_float),
        ("dE", C.c_double), ("R_bend", C.c_double), ("f_rev", C.c_double), ("f_RF", C.c_double),
        ("V0", C.c_double), ("V_eff", C.c_double), ("fs", C.c_double), ("alpha0", C.c_double),
        ("bl", C.c_double), ("Ib", C.c_double), ("Qb", C.c_double), ("steps", C.c_double),
        ("t_damp", C.c_double), ("dt", C.c_double), ("revolutionpart", C.c_double),
        ("padded_bins", C.c_uint32), ("spacing_bins", C.c_uint32),
        ("laststep_per_Ts", C.c_uint32), ("reserved", C.c_uint32),
    ]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


def default_params(**overrides):
    """The reference's defaults (src/IO/ProgramOptions.cpp:29-267)."""
    p = RingParams(
        grid_size=256, steps_per_Ts=1000, interpolation=4, derivation=4, fp_type=3, fp_track=3,
        linear_rf=1, round_padding=1, use_csr=1, reserved=0,
        pq_size=12.0, shift_x=0.0, shift_y=0.0, alpha0=4e-3, alpha1=0.0, alpha2=0.0, f_s=0.0,
        f_rev=9e6, harmonic=50.0, f_cutoff=23e9, padding=8.0, steps_per_Trev=0.0, E0=1.3e9,
        sE=4.7e-4, V_RF=1e6, R_bend=-1.0, t_damp=-1.0, gap=0.03, collimator=0.0,
        wall_conductivity=0.0, wall_susceptibility=0.0, zoom=1.0, bunch_current=1e-3,
        reserved_f=0.0)
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def kara_params(**overrides):
    """The 'KARA-like' common physical inputs of SURVEY.md section 8(d) used by every config."""
    base = dict(f_rev=2.7157e6, harmonic=184.0, E0=1.3e9, sE=4.7e-4, V_RF=1.4e6, f_s=8.5e3,
                R_bend=5.559, t_damp=-1.0, pq_size=12.0, padding=8.0, round_padding=1,
                derivation=4, interpolation=4, fp_type=3, fp_track=3, f_cutoff=23e9,
                linear_rf=1, gap=-1.0, bunch_current=50e-6)
    base.update(overrides)
    return default_params(**base)


class RfModel(C.Structure):
    """ivs_rf_model (include/ivs.h): constants of RFKickMap::_calcKick(phase, ampl)"""

    _fields_ = [("linear", C.c_uint32)] + [(k, C.c_float) for k in (
        "tan_angle", "xcenter", "syncphase", "bl2phase", "qmin", "delta0", "delta1", "dE_scale",
        "revolutionpart", "V_RF", "V0")]


class SimDesc(C.Structure):
    """ivs_sim_desc (include/ivs.h)."""

    _fields_ = [
        ("n", C.c_uint32), ("members", C.c_uint32), ("it", C.c_uint32), ("nmax", C.c_uint32),
        ("n_particles", C.c_uint32), ("fptrack", C.c_uint32), ("flags", C.c_uint32),
        ("strips", C.c_uint32),
        ("dq", C.c_float), ("dp", C.c_float), ("e1", C.c_float), ("qmin", C.c_float),
        ("pmin", C.c_float), ("reserved_f", C.c_float),
        ("rf_offset", C.c_void_p), ("drift_offset", C.c_void_p), ("fp_weights", C.c_void_p),
        ("impedance", C.c_void_p), ("wakescaling", C.c_void_p), ("grid", C.c_void_p),
        ("particles", C.c_void_p),
        ("seed", C.c_uint64),
    ]
