// ref_driver.cpp -- C entry points over the UNMODIFIED reference classes.  TEST INFRASTRUCTURE ONLY.
//
// The reference sources are compiled where they lie under /root/reference (never copied) by
// oracle/Makefile against the shim headers in oracle/shim/, together with this driver, into
// oracle/_ref/libinovesa_ref*.so.  The driver (our code) does what main.cpp would do -- which
// cannot be compiled here because of boost::program_options -- and exposes it to the tests:
//   * single maps (KickMap, RFKickMap, DriftMap, FokkerPlanckMap, RotationMap) on caller data
//   * PhaseSpace projections / moments, ElectricField wake potential and CSR power
//   * impedance tables of makeImpedance()
//   * the whole simulation: scalar derivation as src/main.cpp:163-355,663-664,682, object
//     wiring as :478-763, loop body as :948-1053
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the result.
#include <dlfcn.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#define INOVESA_ALLOW_PS_RESET 1
#include "IO/Display.hpp"
#include "HelperFunctions.hpp"
#include "PS/ElectricField.hpp"
#include "SM/DynamicRFKickMap.hpp"
#include <queue>
#include <array>
#include "PS/PhaseSpace.hpp"
#include "SM/DriftMap.hpp"
#include "SM/FokkerPlanckMap.hpp"
#include "SM/Identity.hpp"
#include "SM/KickMap.hpp"
#include "SM/RFKickMap.hpp"
#include "SM/RotationMap.hpp"
#include "SM/WakePotentialMap.hpp"
#include "Z/ConstImpedance.hpp"
#include "Z/Impedance.hpp"
#include "Z/ImpedanceFactory.hpp"

#include <boost/math/constants/constants.hpp>
#include <boost/math/special_functions/airy.hpp>
#include <boost/math/special_functions/sign.hpp>

#include "ivs_host.h"

using namespace vfps;
using boost::math::constants::two_pi;

namespace {

std::string g_error;

void quiet() {
    Display::silent_mode = true;
    // FFTWWrapper.cpp creates $XDG_DATA_HOME/inovesa/fftwisdom/ ; keep that inside the repo
    if (std::getenv("XDG_DATA_HOME") == nullptr) {
        Dl_info info;
        std::string dir = "/tmp";
        if (dladdr(reinterpret_cast<void*>(&quiet), &info) && info.dli_fname) {
            std::string so(info.dli_fname);
            const size_t pos = so.rfind('/');
            if (pos != std::string::npos) dir = so.substr(0, pos);
        }
        dir += "/xdg";
        setenv("XDG_DATA_HOME", dir.c_str(), 0);
    }
}

// exposes protected members of the maps to the driver
struct KickMapProbe : public KickMap {
    using KickMap::KickMap;
    const std::vector<meshaxis_t>& offsets() const { return _offset; }
};
struct RFKickMapProbe : public RFKickMap {
    using RFKickMap::RFKickMap;
    const std::vector<meshaxis_t>& offsets() const { return _offset; }
};
struct DriftMapProbe : public DriftMap {
    using DriftMap::DriftMap;
    const std::vector<meshaxis_t>& offsets() const { return _offset; }
};
struct FokkerPlanckProbe : public FokkerPlanckMap {
    using FokkerPlanckMap::FokkerPlanckMap;
    // stencil as 5 weights on y-2..y+2 (index/weight table -> dense), plus the raw table
    void dense(uint32_t n, float* w5) const {
        for (uint32_t y = 0; y < n; y++) {
            for (int k = 0; k < 5; k++) w5[y * 5 + k] = 0;
            for (uint32_t j = 0; j < _ip; j++) {
                const hi h = _hinfo[y * _ip + j];
                if (h.weight == 0) continue;
                const int rel = int(h.index) - int(y) + 2;
                w5[y * 5 + rel] = h.weight;
            }
        }
    }
    void raw(uint32_t n, uint32_t* index, float* weight) const {
        for (uint32_t i = 0; i < n * _ip; i++) {
            index[i] = _hinfo[i].index;
            weight[i] = _hinfo[i].weight;
        }
    }
};

std::shared_ptr<PhaseSpace> make_ps(uint32_t n, float qmin, float qmax, float pmin, float pmax,
                                    const float* data, double bl = 1, double dE = 1) {
    PhaseSpace::resetSize(n, 1);
    std::vector<integral_t> filling{1};
    return std::make_shared<PhaseSpace>(qmin, qmax, bl, pmin, pmax, dE, nullptr, 1, 1, filling,
                                        1, data);
}

template <typename F>
int guarded(F&& f) {
    try {
        quiet();
        f();
        return 0;
    } catch (const std::invalid_argument& e) {
        g_error = e.what();
        return 1;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 2;
    }
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

// ---------------------------------------------------------------- single maps

// KickMap with explicit offsets (swapOffset), apply to `in`, optionally push particles.
// axis: 0 = x, 1 = y.  particles: np pairs (x,y) updated in place.
int ref_kickmap_apply(uint32_t n, uint32_t it, int axis, const float* offset, const float* in,
                      float* out, float* particles, uint32_t np) {
    return guarded([&] {
        auto g1 = make_ps(n, -1, 1, -1, 1, in);
        auto g2 = make_ps(n, -1, 1, -1, 1, in);
        KickMap km(g1, g2, static_cast<SourceMap::InterpolationType>(it), false,
                   axis == 0 ? KickMap::Axis::x : KickMap::Axis::y, nullptr);
        std::vector<meshaxis_t> offs(offset, offset + n);
        km.swapOffset(offs);
        km.apply();
        std::copy_n(g2->getData(), size_t(n) * n, out);
        for (uint32_t i = 0; i < np; i++) {
            PhaseSpace::Position pos{particles[2 * i], particles[2 * i + 1]};
            km.applyTo(pos);
            particles[2 * i] = pos.x;
            particles[2 * i + 1] = pos.y;
        }
    });
}

// FokkerPlanckMap on an energy axis [pmin,pmax]; returns the dense 5-point weights too
int ref_fokker_planck_apply(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                            uint32_t derivation, const float* in, float* out, float* w5) {
    return guarded([&] {
        auto g1 = make_ps(n, pmin, pmax, pmin, pmax, in);
        auto g2 = make_ps(n, pmin, pmax, pmin, pmax, in);
        FokkerPlanckProbe fpm(g1, g2, n, n, static_cast<FokkerPlanckMap::FPType>(fptype),
                              FokkerPlanckMap::FPTracking::none, e1,
                              static_cast<FokkerPlanckMap::DerivationType>(derivation), nullptr);
        if (w5) fpm.dense(n, w5);
        if (in && out) {
            fpm.apply();
            std::copy_n(g2->getData(), size_t(n) * n, out);
        }
    });
}

// FokkerPlanckMap::applyTo for the deterministic tracking modes (0,1,2)
int ref_fokker_planck_track(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                            uint32_t derivation, uint32_t fptrack, const float* grid,
                            float* particles, uint32_t np) {
    return guarded([&] {
        auto g1 = make_ps(n, pmin, pmax, pmin, pmax, grid);
        auto g2 = make_ps(n, pmin, pmax, pmin, pmax, grid);
        FokkerPlanckMap fpm(g1, g2, n, n, static_cast<FokkerPlanckMap::FPType>(fptype),
                            static_cast<FokkerPlanckMap::FPTracking>(fptrack), e1,
                            static_cast<FokkerPlanckMap::DerivationType>(derivation), nullptr);
        for (uint32_t i = 0; i < np; i++) {
            PhaseSpace::Position pos{particles[2 * i], particles[2 * i + 1]};
            fpm.applyTo(pos);
            particles[2 * i] = pos.x;
            particles[2 * i + 1] = pos.y;
        }
    });
}

// RotationMap; rotmapsize = 0 (on the fly) or n*n (table).  Returns 1 for the
// std::invalid_argument cases of src/SM/RotationMap.cpp:83-87.
int ref_rotation_apply(uint32_t n, float qmin, float qmax, float pmin, float pmax, float angle,
                       uint32_t it, int clamp, uint32_t rotmapsize, const float* in, float* out,
                       float* particles, uint32_t np) {
    return guarded([&] {
        auto g1 = make_ps(n, qmin, qmax, pmin, pmax, in);
        auto g2 = make_ps(n, qmin, qmax, pmin, pmax, in);
        RotationMap rm(g1, g2, n, n, angle, static_cast<SourceMap::InterpolationType>(it),
                       clamp != 0, rotmapsize, nullptr);
        rm.apply();
        std::copy_n(g2->getData(), size_t(n) * n, out);
        for (uint32_t i = 0; i < np; i++) {
            PhaseSpace::Position pos{particles[2 * i], particles[2 * i + 1]};
            rm.applyTo(pos);
            particles[2 * i] = pos.x;
            particles[2 * i + 1] = pos.y;
        }
    });
}

// SourceMap::calcCoefficiants
int ref_calc_coefficients(float f, uint32_t it, float* ic) {
    return guarded([&] { SourceMap::calcCoefficiants(ic, f, it); });
}

// ---------------------------------------------------------------- PhaseSpace

// projections [2][n], integral, moments {mean_q,var_q,rms_q,mean_p,var_p,rms_p}, Simpson weights
int ref_phasespace_reduce(uint32_t n, float qmin, float qmax, float pmin, float pmax,
                          const float* data, float* projx, float* projy, float* integral,
                          float* moments6) {
    return guarded([&] {
        auto g = make_ps(n, qmin, qmax, pmin, pmax, data);
        g->updateXProjection();
        g->integrate();
        g->variance(0);
        g->updateYProjection();
        g->variance(1);
        if (projx) std::copy_n(g->getProjection(0).origin(), n, projx);
        if (projy) std::copy_n(g->getProjection(1).origin(), n, projy);
        if (integral) *integral = g->getIntegral();
        if (moments6) {
            moments6[0] = g->getMoment(0, 0)[0];
            moments6[1] = g->getMoment(0, 1)[0];
            moments6[2] = g->getBunchLength()[0];
            moments6[3] = g->getMoment(1, 0)[0];
            moments6[4] = g->getMoment(1, 1)[0];
            moments6[5] = g->getEnergySpread()[0];
        }
    });
}

// default-constructed Gaussian phase space (data == nullptr) after main()'s renormalisation
int ref_initial_grid(uint32_t n, float qmin, float qmax, float pmin, float pmax, double zoom,
                     float* grid) {
    return guarded([&] {
        PhaseSpace::resetSize(n, 1);
        std::vector<integral_t> filling{1};
        PhaseSpace ps(qmin, qmax, 1.0, pmin, pmax, 1.0, nullptr, 1, 1, filling, zoom);
        ps.updateXProjection();  // src/main.cpp:528-532
        ps.normalize();
        std::copy_n(ps.getData(), size_t(n) * n, grid);
    });
}

// normalize(): scale data so that the integral becomes 1 (single bunch, filling_set = 1)
int ref_normalize(uint32_t n, float qmin, float qmax, float pmin, float pmax, const float* data,
                  float* out) {
    return guarded([&] {
        auto g = make_ps(n, qmin, qmax, pmin, pmax, data);
        g->updateXProjection();
        g->integrate();
        g->normalize();
        std::copy_n(g->getData(), size_t(n) * n, out);
    });
}

// ---------------------------------------------------------------- ElectricField

// wakePotential() for a given grid, impedance table (nmax complex) and scaling; like
// test/TestElectricField.cpp the field is built with the `wakescalining` constructor.
// The wake-capable constructor is the (Ib,E0,sigma_delta,dt) one: choose them so that
// Ib*dt*c/bl/(dp*sigma_delta*E0) == wakescaling_raw (bl = 1).
int ref_wake_potential(uint32_t n, float qmin, float qmax, float pmin, float pmax,
                       const float* grid, const float* z, uint32_t nmax, double wakescaling_raw,
                       float* wake, float* formfactor, float* padded_wake,
                       float* wakescaling_used) {
    return guarded([&] {
        auto g = make_ps(n, qmin, qmax, pmin, pmax, grid, 1.0, 1.0);
        g->updateXProjection();
        std::vector<impedance_t> zv(nmax);
        for (uint32_t i = 0; i < nmax; i++) zv[i] = impedance_t(z[2 * i], z[2 * i + 1]);
        auto imp = std::make_shared<Impedance>(zv, 1.0f, nullptr);
        std::vector<uint32_t> buckets{0};
        // scaling = Ib*dt*c/scale(Meter)/(delta1*sigma_delta*E0): scale = 1, E0 = c, Ib = raw*dp
        const double dp = g->getDelta(1);
        ElectricField f(g, imp, buckets, 0, nullptr, 1.0, 1.0f, wakescaling_raw * dp,
                        physcons::c, 1.0, 1.0);
        const meshaxis_t* w = f.wakePotential();
        std::copy_n(w, n, wake);
        // (the driver is compiled with -fno-access-control to read these private members)
        if (formfactor) {
            for (uint32_t i = 0; i <= nmax / 2; i++) {
                formfactor[2 * i] = f._formfactor[i].real();
                formfactor[2 * i + 1] = f._formfactor[i].imag();
            }
        }
        if (padded_wake) std::copy_n(f.getPaddedWakePotential(), nmax, padded_wake);
        if (wakescaling_used) *wakescaling_used = f._wakescaling;
    });
}

// wakePotential() of a train of nb bunches (PhaseSpace::nb > 1): grids [nb][n][n] holding the bunches'
// charge fractions, bucket numbers and the spacing in bins as src/main.cpp:248-321 hands them to the field
int ref_train_wake_potential(uint32_t n, uint32_t nb, float qmin, float qmax, float pmin, float pmax,
                             const float* grids, const uint32_t* bucket, uint32_t spacing_bins,
                             const float* z, uint32_t nmax, double wakescaling_raw, float* wake,
                             float* padded_wake, float* wakescaling_used) {
    return guarded([&] {
        PhaseSpace::resetSize(n, nb);
        std::vector<integral_t> filling(nb, integral_t(1) / nb);
        auto g = std::make_shared<PhaseSpace>(qmin, qmax, 1.0, pmin, pmax, 1.0, nullptr, 1, 1, filling, 1, grids);
        g->updateXProjection();
        std::vector<impedance_t> zv(nmax);
        for (uint32_t i = 0; i < nmax; i++) zv[i] = impedance_t(z[2 * i], z[2 * i + 1]);
        auto imp = std::make_shared<Impedance>(zv, 1.0f, nullptr);
        std::vector<uint32_t> buckets(bucket, bucket + nb);
        const double dp = g->getDelta(1);
        ElectricField f(g, imp, buckets, spacing_bins, nullptr, 1.0, 1.0f, wakescaling_raw * dp,
                        physcons::c, 1.0, 1.0);
        const meshaxis_t* w = f.wakePotential();
        std::copy_n(w, (size_t)nb * n, wake);
        if (padded_wake) std::copy_n(f.getPaddedWakePotential(), nmax, padded_wake);
        if (wakescaling_used) *wakescaling_used = f._wakescaling;
        PhaseSpace::resetSize(n, 1);
    });
}

// updateCSR(): spectrum (nmax) and intensity; bl sets the "Meter" scale of axis 0
int ref_csr_power(uint32_t n, float qmin, float qmax, float pmin, float pmax, double bl,
                  const float* grid, const float* z, uint32_t nmax, float cutoff, float* spectrum,
                  float* power) {
    return guarded([&] {
        auto g = make_ps(n, qmin, qmax, pmin, pmax, grid, bl, 1.0);
        g->updateXProjection();
        std::vector<impedance_t> zv(nmax);
        for (uint32_t i = 0; i < nmax; i++) zv[i] = impedance_t(z[2 * i], z[2 * i + 1]);
        auto imp = std::make_shared<Impedance>(zv, 1.0f, nullptr);
        std::vector<uint32_t> buckets{0};
        ElectricField f(g, imp, buckets, 0, nullptr, 1.0, 1.0f);
        const csrpower_t* s = f.updateCSR(cutoff);
        if (spectrum) std::copy_n(s, nmax, spectrum);
        *power = f.getCSRPower()[0];
    });
}

// ---------------------------------------------------------------- impedances

int ref_make_impedance(size_t nfreqs, float fmax, double R_bend, double f_rev, double gap,
                       int use_csr, double s, double xi, double collimator, float* z,
                       int* nonzero) {
    return guarded([&] {
        auto imp = makeImpedance(nfreqs, nullptr, fmax, R_bend, f_rev, gap, use_csr != 0, s, xi,
                                 collimator);
        *nonzero = (imp != nullptr);
        for (size_t i = 0; i < nfreqs; i++) {
            const impedance_t v = imp ? (*imp)[i] : impedance_t(0, 0);
            z[2 * i] = v.real();
            z[2 * i + 1] = v.imag();
        }
    });
}

int ref_airy(double x, double* ai, double* aip, double* bi, double* bip) {
    int over = 0;
    *ai = boost::math::airy_ai(x);
    *aip = boost::math::airy_ai_prime(x);
    try {
        *bi = boost::math::airy_bi(x);
        *bip = boost::math::airy_bi_prime(x);
    } catch (...) {
        over = 1;
        *bi = *bip = INFINITY;
    }
    return over;
}

// ---------------------------------------------------------------- whole simulation

struct ref_sim {
    ivs_derived d;
    uint32_t n;
    std::shared_ptr<PhaseSpace> t1, t2, t3;
    std::shared_ptr<SourceMap> rfm;
    std::unique_ptr<DriftMap> drm;
    std::unique_ptr<SourceMap> fpm;
    std::shared_ptr<Impedance> wake_impedance, rdtn_impedance;
    std::unique_ptr<ElectricField> rdtn_field, wake_field;
    std::unique_ptr<SourceMap> wm_identity;
    std::unique_ptr<WakePotentialMap> wkm;
    std::vector<PhaseSpace::Position> trackme;
    uint64_t step = 0;
    float fc;
};

// scalar derivation, following src/main.cpp:163-355 statement by statement
static void derive(const ivs_ring_params& o, ivs_derived& d) {
    std::memset(&d, 0, sizeof(d));
    const auto ps_bins = o.grid_size;
    const auto pqsize = o.pq_size;
    const auto qcenter = -o.shift_x * pqsize / (ps_bins - 1);
    const auto pcenter = -o.shift_y * pqsize / (ps_bins - 1);
    const auto pqhalf = pqsize / 2;
    d.qmax = qcenter + pqhalf;
    d.qmin = qcenter - pqhalf;
    d.pmax = pcenter + pqhalf;
    d.pmin = pcenter - pqhalf;
    const auto sE = o.sE;
    const auto E0 = o.E0;
    const auto dE = sE * E0;
    const auto use_set_bend = (o.R_bend > 0);
    const auto f_rev = static_cast<double>(o.f_rev);
    const auto R_bend = use_set_bend ? o.R_bend : physcons::c / (two_pi<double>() * f_rev);
    const auto harmonic_number = static_cast<double>(o.harmonic);
    const auto f_RF = f_rev * harmonic_number;
    const auto bunchspacing = 1.0 / (f_RF);
    const auto V_RF = o.V_RF;
    const auto lorentzgamma = E0 / physcons::me;
    const auto V0 = physcons::e * std::pow(lorentzgamma, 4) / (3 * physcons::epsilon0 * R_bend);
    const auto W0 = V0 * physcons::e;
    const double V_eff = std::sqrt(V_RF * V_RF - V0 * V0);
    auto fs = static_cast<double>(o.f_s);
    auto alpha0_tmp = static_cast<double>(o.alpha0);
    if (std::fpclassify(fs) == FP_ZERO) {
        fs = f_rev * std::sqrt(alpha0_tmp * harmonic_number * V_eff / (two_pi<double>() * E0));
    } else {
        alpha0_tmp = boost::math::sign(fs) * two_pi<double>() * E0 / (harmonic_number * V_eff) *
                     std::pow(fs / f_rev, 2);
    }
    std::vector<meshaxis_t> alpha{{static_cast<meshaxis_t>(alpha0_tmp), o.alpha1, o.alpha2}};
    const double bl = physcons::c * dE / harmonic_number / std::pow(f_rev, 2.0) / V_eff * fs;
    std::vector<integral_t> bunches{o.bunch_current};
    const double Ib = std::accumulate(bunches.begin(), bunches.end(), 0.0);
    const double Qb = Ib / f_rev;
    const double steps = (o.steps_per_Trev > 0) ? o.steps_per_Trev * f_rev / fs
                                                : std::max(o.steps_per_Ts, 1u);
    const auto calc_damp = E0 * physcons::e / W0 / f_rev;
    const auto set_damp = o.t_damp;
    const auto t_damp = (set_damp < 0) ? calc_damp : set_damp;
    const double dt = 1.0 / (fs * steps);
    const double revolutionpart = f_rev * dt;
    const double spacing_ps = (bunchspacing * physcons::c / bl / pqsize);
    const meshindex_t spacing_bins = std::round(ps_bins * spacing_ps);
    const frequency_t fmax = ps_bins * vfps::physcons::c / (pqsize * bl);
    double padding = std::max(o.padding, 1.0);
    size_t padded_bins = std::ceil(ps_bins * padding);
    if (o.round_padding) padded_bins = upper_power_of_two(padded_bins);
    const meshaxis_t angle = two_pi<double>() / steps;
    const timeaxis_t e1 = (t_damp > 0) ? 2.0 / (fs * t_damp * steps) : 0;

    d.dE = dE; d.R_bend = R_bend; d.f_rev = f_rev; d.f_RF = f_RF; d.V0 = V0; d.V_eff = V_eff;
    d.fs = fs; d.alpha0 = alpha0_tmp; d.bl = bl; d.Ib = Ib; d.Qb = Qb; d.steps = steps;
    d.t_damp = t_damp; d.dt = dt; d.revolutionpart = revolutionpart;
    d.padded_bins = padded_bins; d.spacing_bins = spacing_bins;
    d.laststep_per_Ts = std::ceil(steps);
    d.angle = angle; d.fmax = fmax; d.e1 = e1;
    d.slip[0] = angle;                          // src/main.cpp:663-664
    d.slip[1] = alpha[1] / alpha[0] * angle;
    d.slip[2] = alpha[2] / alpha[0] * angle;
}

void* ref_sim_create(const ivs_ring_params* op, const float* grid, const float* particles,
                     uint32_t np) {
    ref_sim* s = nullptr;
    const int rc = guarded([&] {
        const ivs_ring_params& o = *op;
        s = new ref_sim;
        derive(o, s->d);
        ivs_derived& d = s->d;
        const uint32_t n = o.grid_size;
        s->n = n;
        s->fc = o.f_cutoff;
        const auto interpolationtype = static_cast<SourceMap::InterpolationType>(o.interpolation);
        const bool interpol_clamp = false;
        std::vector<integral_t> bunches{1};
        PhaseSpace::resetSize(n, 1);  // src/main.cpp:478
        if (grid == nullptr) {
            s->t1.reset(new PhaseSpace(d.qmin, d.qmax, d.bl, d.pmin, d.pmax, d.dE, nullptr, d.Qb,
                                       d.Ib, bunches, o.zoom));  // :484-485
            s->t1->updateXProjection();  // :528-532 (RenormalizeCharge default 0)
            s->t1->normalize();
        } else {
            // caller-provided start distribution (PhaseSpace ctor `data`, PhaseSpace.hpp:60)
            s->t1.reset(new PhaseSpace(d.qmin, d.qmax, d.bl, d.pmin, d.pmax, d.dE, nullptr, d.Qb,
                                       d.Ib, bunches, o.zoom, grid));
        }
        s->t2 = std::make_shared<PhaseSpace>(*s->t1);  // :534-535
        s->t3 = std::make_shared<PhaseSpace>(*s->t1);
        d.dq = s->t1->getDelta(0);
        d.dp = s->t1->getDelta(1);
        d.zerobin_x = s->t1->getAxis(0)->zerobin();
        d.zerobin_y = s->t1->getAxis(1)->zerobin();
        d.bl_f = s->t1->getScale(0, "Meter");
        d.dE_f = s->t1->getScale(1, "ElectronVolt");

        if (o.linear_rf) {  // :640-649
            s->rfm.reset(new RFKickMapProbe(s->t2, s->t1, d.angle, d.f_RF, interpolationtype,
                                            interpol_clamp, nullptr));
        } else {
            s->rfm.reset(new RFKickMapProbe(s->t2, s->t1, d.revolutionpart, d.V_eff, d.f_RF, d.V0,
                                            interpolationtype, interpol_clamp, nullptr));
        }
        const std::vector<meshaxis_t> slip{{d.slip[0], d.slip[1], d.slip[2]}};
        s->drm.reset(new DriftMapProbe(s->t1, s->t3, slip, o.E0, interpolationtype,
                                       interpol_clamp, nullptr));  // :677-679
        if (d.e1 > 0) {  // :685-707
            s->fpm.reset(new FokkerPlanckProbe(
                s->t3, s->t1, n, n, static_cast<FokkerPlanckMap::FPType>(o.fp_type),
                static_cast<FokkerPlanckMap::FPTracking>(o.fp_track), d.e1,
                static_cast<FokkerPlanckMap::DerivationType>(o.derivation), nullptr));
        } else {
            s->fpm.reset(new Identity(s->t3, s->t1, nullptr));
        }
        s->wake_impedance = makeImpedance(d.padded_bins, nullptr, d.fmax, d.R_bend, d.f_rev, o.gap,
                                          o.use_csr != 0, o.wall_conductivity,
                                          o.wall_susceptibility, o.collimator);  // :717-721
        s->rdtn_impedance = makeImpedance(d.padded_bins, nullptr, d.fmax, d.R_bend, d.f_rev,
                                          (o.gap > 0) ? o.gap : -1);  // :724-727
        std::vector<uint32_t> bucketnumbers{0};
        s->rdtn_field.reset(new ElectricField(s->t1, s->rdtn_impedance, bucketnumbers, 0, nullptr,
                                              d.f_rev, d.revolutionpart));  // :731-734
        if (s->wake_impedance != nullptr) {  // :748-763
            s->wake_field.reset(new ElectricField(s->t1, s->wake_impedance, bucketnumbers,
                                                  d.spacing_bins, nullptr, d.f_rev,
                                                  d.revolutionpart, d.Ib, o.E0, o.sE, d.dt));
            s->wkm.reset(new WakePotentialMap(s->t1, s->t2, s->wake_field.get(), interpolationtype,
                                              interpol_clamp, nullptr));
        } else {
            s->wm_identity.reset(new Identity(s->t1, s->t2, nullptr));
        }
        {
            // ElectricField::_wakescaling = Ib*dt*c/bl/(dp*sE*E0)/nmax (hpp:99-100, cpp:58);
            // reported even when no wake map is built
            const meshaxis_t raw = d.Ib * d.dt * physcons::c / s->t1->getScale(0, "Meter") /
                                   (s->t1->getDelta(1) * o.sE * o.E0);
            d.wakescaling = raw / d.padded_bins;
            if (s->wake_field) d.wakescaling = s->wake_field->_wakescaling;
        }
        for (uint32_t i = 0; i < np; i++) {
            s->trackme.push_back({particles[2 * i], particles[2 * i + 1]});
        }
        s->t1->updateXProjection();  // :893-894
        s->t1->integrate();
    });
    if (rc != 0) {
        delete s;
        return nullptr;
    }
    return s;
}

void ref_sim_destroy(void* h) { delete static_cast<ref_sim*>(h); }

// loop body of src/main.cpp:948-1053 (no output, no renormalisation: RenormalizeCharge = 0)
int ref_sim_step(void* h, uint32_t nsteps) {
    ref_sim* s = static_cast<ref_sim*>(h);
    return guarded([&] {
        for (uint32_t i = 0; i < nsteps; i++) {
            if (s->wkm) s->wkm->update();
            s->t1->integrate();
            SourceMap* wm = s->wkm ? static_cast<SourceMap*>(s->wkm.get()) : s->wm_identity.get();
            wm->apply();
            if (!s->trackme.empty()) wm->applyToAll(s->trackme);
            s->rfm->apply();
            if (!s->trackme.empty()) s->rfm->applyToAll(s->trackme);
            s->drm->apply();
            if (!s->trackme.empty()) s->drm->applyToAll(s->trackme);
            s->fpm->apply();
            if (!s->trackme.empty()) s->fpm->applyToAll(s->trackme);
            s->t1->updateXProjection();
            s->step++;
        }
    });
}

int ref_sim_derived(void* h, ivs_derived* d) {
    *d = static_cast<ref_sim*>(h)->d;
    return 0;
}

// state after the last step: grid_t1, its X projection, the wake offsets of the last update()
int ref_sim_get(void* h, float* grid, float* profile, float* wake, float* particles) {
    ref_sim* s = static_cast<ref_sim*>(h);
    const uint32_t n = s->n;
    if (grid) std::copy_n(s->t1->getData(), size_t(n) * n, grid);
    if (profile) std::copy_n(s->t1->getProjection(0).origin(), n, profile);
    if (wake) {
        if (s->wkm) std::copy_n(s->wkm->getForce(), n, wake);
        else std::fill_n(wake, n, 0.0f);
    }
    if (particles) {
        for (size_t i = 0; i < s->trackme.size(); i++) {
            particles[2 * i] = s->trackme[i].x;
            particles[2 * i + 1] = s->trackme[i].y;
        }
    }
    return 0;
}

// static tables of the maps: RF offsets, drift offsets (n each), FP dense weights (n*5),
// wake / radiation impedance (padded_bins complex each; wake may be absent -> returns zeros)
int ref_sim_tables(void* h, float* rf, float* drift, float* fp5, float* zwake, float* zrad) {
    ref_sim* s = static_cast<ref_sim*>(h);
    const uint32_t n = s->n;
    if (rf) std::copy_n(static_cast<RFKickMapProbe*>(s->rfm.get())->offsets().data(), n, rf);
    if (drift) std::copy_n(static_cast<DriftMapProbe*>(s->drm.get())->offsets().data(), n, drift);
    if (fp5) {
        auto* fp = dynamic_cast<FokkerPlanckProbe*>(s->fpm.get());
        if (fp) fp->dense(n, fp5);
        else for (uint32_t y = 0; y < n; y++) for (int k = 0; k < 5; k++) fp5[y * 5 + k] = (k == 2);
    }
    const uint32_t nm = s->d.padded_bins;
    if (zwake) {
        for (uint32_t i = 0; i < nm; i++) {
            const impedance_t v = s->wake_impedance ? (*s->wake_impedance)[i] : impedance_t(0, 0);
            zwake[2 * i] = v.real();
            zwake[2 * i + 1] = v.imag();
        }
    }
    if (zrad) {
        for (uint32_t i = 0; i < nm; i++) {
            const impedance_t v = (*s->rdtn_impedance)[i];
            zrad[2 * i] = v.real();
            zrad[2 * i + 1] = v.imag();
        }
    }
    return 0;
}

// out-step observables of src/main.cpp:961-1033:
// obs = {integral, mean_q, var_q, sigma_q, mean_p, var_p, sigma_p, csr_power}
int ref_sim_observe(void* h, float* obs, float* csr_spectrum, float* projy) {
    ref_sim* s = static_cast<ref_sim*>(h);
    return guarded([&] {
        s->t1->integrate();
        s->t1->variance(0);
        s->t1->updateYProjection();
        s->t1->variance(1);
        s->rdtn_field->updateCSR(s->fc);
        obs[0] = s->t1->getIntegral();
        obs[1] = s->t1->getMoment(0, 0)[0];
        obs[2] = s->t1->getMoment(0, 1)[0];
        obs[3] = s->t1->getBunchLength()[0];
        obs[4] = s->t1->getMoment(1, 0)[0];
        obs[5] = s->t1->getMoment(1, 1)[0];
        obs[6] = s->t1->getEnergySpread()[0];
        obs[7] = s->rdtn_field->getCSRPower()[0];
        if (csr_spectrum) std::copy_n(s->rdtn_field->getCSRSpectrum(), s->d.padded_bins, csr_spectrum);
        if (projy) std::copy_n(s->t1->getProjection(1).origin(), s->n, projy);
    });
}

// DynamicRFKickMap (src/SM/DynamicRFKickMap.cpp) with a GIVEN modulation queue: the object is built
// by the reference's sine-RF constructor with zero noise / modulation and its private queue of
// (phase, amplitude) pairs is replaced (the reference fills it from an unseeded mt19937, :36,80-94).
// Only for the non-linear RF: the linear constructor of this snapshot binds to the wrong
// RFKickMap constructor.
int ref_sim_set_rf_modulation(void* h, const float* phase_ampl, uint32_t count, int interpolation) {
    ref_sim* s = static_cast<ref_sim*>(h);
    return guarded([&] {
        const ivs_derived& d = s->d;
        auto* m = new DynamicRFKickMap(s->t2, s->t1, s->n, s->n, d.revolutionpart, d.V_eff, d.f_RF, d.V0,
                                       0, 0, 0, 0, 0, static_cast<SourceMap::InterpolationType>(interpolation),
                                       false, nullptr);
        std::queue<std::array<meshaxis_t, 2>> q;
        for (uint32_t i = 0; i < count; i++) q.push({{phase_ampl[2 * i], phase_ampl[2 * i + 1]}});
        m->_next_modulation = std::move(q);
        s->rfm.reset(m);
    });
}

// unit factors the reference writes beside /WakePotential/data, /CSR/Spectrum/data and
// /CSR/Intensity/data (src/IO/HDF5File.cpp:259-276): ElectricField::volts (src/PS/ElectricField.cpp:23),
// factor4WattPerHertz and factor4Watts (:39-40)
int ref_sim_units(void* h, double* out3) {
    ref_sim* s = static_cast<ref_sim*>(h);
    return guarded([&] {
        const ElectricField* f = s->wake_field ? s->wake_field.get() : s->rdtn_field.get();
        out3[0] = f->volts;
        out3[1] = s->rdtn_field->factor4WattPerHertz;
        out3[2] = s->rdtn_field->factor4Watts;
    });
}

}  // extern "C"
