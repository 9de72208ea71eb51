// The reference's unit tests for the hot path (test/TestKickMap.cpp, TestRotationMap.cpp,
// TestElectricField.cpp, TestPhaseSpace.cpp, TestRuler.cpp), re-stated against the C++ host
// mirror include/vfps_b200.hpp: same objects, same calls, same known answers (golden_vectors.hpp
// is generated from tests/golden/reference_tests.json by tests/test_cpp_host_mirror.py).
//
//   test_host_mirror --host-only   : what needs no GPU (Ruler, argument errors, loud failure
//                                    without a device)
//   test_host_mirror               : everything, on cuda:0
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "golden_vectors.hpp"
#include "vfps_b200.hpp"

using namespace vfps;

static int failures = 0, checks = 0;
#define CHECK(cond)                                                                   \
    do {                                                                              \
        checks++;                                                                     \
        if (!(cond)) { failures++; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
#define CHECK_CLOSE(a, b, tol)                                                        \
    do {                                                                              \
        checks++;                                                                     \
        const double a__ = (a), b__ = (b);                                            \
        if (!(std::fabs(a__ - b__) <= (tol))) {                                       \
            failures++;                                                               \
            std::printf("FAIL %s:%d  %s = %.9g, expected %.9g (tol %g)\n", __FILE__, __LINE__, #a, a__, b__, (double)(tol)); \
        }                                                                             \
    } while (0)
template <class E, class F>
static bool throws(F f) {
    try { f(); } catch (const E&) { return true; } catch (...) { return false; }
    return false;
}

static void host_only() {
    // test/TestRuler.cpp
    Ruler<float> r(golden::ruler_steps, golden::ruler_min, golden::ruler_max);
    CHECK_CLOSE(r.zerobin(), golden::ruler_zerobin, 0);
    CHECK(throws<std::invalid_argument>([] { Ruler<float> bad(4, 1.0f, 1.0f); }));
    CHECK(throws<std::invalid_argument>([] { Ruler<float> bad(4, 2.0f, 1.0f); }));
    // no CPU path: a null handler is an argument error, never a silent fall-back
    CHECK(throws<std::invalid_argument>([] {
        PhaseSpace ps(4, -1, 1, 1, -1, 1, 1, nullptr, 1, 1);
    }));
}

static std::shared_ptr<PhaseSpace> grid4(cudahptr_t h, const float* axes, const float* data) {
    return std::make_shared<PhaseSpace>(4, axes[0], axes[1], 1, axes[2], axes[3], 1, h, 1, 1,
                                        std::vector<integral_t>{1}, 1, data);
}

static void kickmaps(cudahptr_t h) {  // test/TestKickMap.cpp:18-172
    const float axes[4] = {-1, 1, -1, 1};
    for (const auto& k : golden::kickmap) {
        auto g1 = grid4(h, axes, k.in), g2 = grid4(h, axes, k.in);
        KickMap km(g1, g2, SourceMap::linear, false, k.axis == 0 ? KickMap::Axis::x : KickMap::Axis::y, h);
        std::vector<meshaxis_t> offs(k.offset, k.offset + 4);
        km.swapOffset(offs);
        km.apply();
        g2->syncDevice(PhaseSpace::clCopyDirection::dev2cpu);
        for (int i = 0; i < 16; i++) CHECK_CLOSE(g2->getData()[i], k.out[i], golden::kickmap_tolerance);
        std::vector<PhaseSpace::Position> pos;
        for (int i = 0; i < 9; i++) pos.push_back({0.5f * i, 0.5f * i});
        km.applyToAll(pos);
        for (int i = 0; i < 9; i++) {
            CHECK_CLOSE(pos[i].x, k.particles_out[i][0], golden::kickmap_tolerance);
            CHECK_CLOSE(pos[i].y, k.particles_out[i][1], golden::kickmap_tolerance);
        }
    }
}

static void rotationmaps(cudahptr_t h) {  // test/TestRotationMap.cpp
    for (const auto& r : golden::rotationmap) {
        auto g1 = grid4(h, r.axes, r.in), g2 = grid4(h, r.axes, r.in);
        RotationMap rm(g1, g2, 4, 4, r.angle, (SourceMap::InterpolationType)r.it, r.clamp != 0, r.rotmapsize, h);
        rm.apply();
        g2->syncDevice(PhaseSpace::clCopyDirection::dev2cpu);
        for (int i = 0; i < 16; i++) CHECK_CLOSE(g2->getData()[i], r.out[i], golden::rotationmap_tolerance);
        if (r.has_particle) {
            PhaseSpace::Position p{r.particle_in[0], r.particle_in[1]};
            rm.applyTo(p);
            CHECK_CLOSE(p.x, r.particle_out[0], r.particle_tolerance);
            CHECK_CLOSE(p.y, r.particle_out[1], r.particle_tolerance);
        }
    }
    for (const auto& bad : golden::rotationmap_invalid) {  // :82-105, :139-160
        const float axes[4] = {-1, 1, -1, 1};
        const float zeros[16] = {0};
        auto g1 = grid4(h, axes, zeros), g2 = grid4(h, axes, zeros);
        CHECK(throws<std::invalid_argument>([&] {
            RotationMap rm(g1, g2, 4, 4, 1.0f, (SourceMap::InterpolationType)bad.it, bad.clamp != 0, bad.rotmapsize, h);
        }));
    }
}

static void phasespace(cudahptr_t h) {  // test/TestPhaseSpace.cpp:38-107
    PhaseSpace ps(128, -6, 6, 1, -6, 6, 1, h, 1, 1);
    CHECK_CLOSE(ps.getIntegral(), 1.0, golden::phasespace_tolerance);
    ps.variance(0);
    ps.variance(1);
    CHECK_CLOSE(ps.getMoment(0, 0), 0.0, golden::phasespace_tolerance);
    CHECK_CLOSE(ps.getMoment(1, 0), 0.0, golden::phasespace_tolerance);
    CHECK_CLOSE(ps.getBunchLength(), 1.0, golden::phasespace_tolerance);
    CHECK_CLOSE(ps.getEnergySpread(), 1.0, golden::phasespace_tolerance);
    CHECK(throws<std::invalid_argument>([&] {  // non-normalised filling
        PhaseSpace bad(32, -6, 6, 1, -6, 6, 1, h, 1, 1, std::vector<integral_t>{0.5f});
    }));
    PhaseSpace copy(ps);  // data copy
    copy.syncDevice(PhaseSpace::clCopyDirection::dev2cpu);
    ps.syncDevice(PhaseSpace::clCopyDirection::dev2cpu);
    CHECK(std::memcmp(copy.getData(), ps.getData(), 128 * 128 * sizeof(float)) == 0);
}

static void electricfield(cudahptr_t h) {  // test/TestElectricField.cpp:147-239
    const meshindex_t n = golden::ef_n;
    const size_t nmax = golden::ef_nmax;
    auto ps = std::make_shared<PhaseSpace>(n, golden::ef_axes[0], golden::ef_axes[1], 1, golden::ef_axes[2],
                                           golden::ef_axes[3], 1, h, 1, 1);
    // Z = 1 below Nyquist: the wake potential is the bunch profile times the scaling (:147-168)
    std::vector<impedance_t> one(nmax, 0);
    for (size_t i = 0; i < nmax / 2; i++) one[i] = 1;  // ConstImpedance: n < nfreqs/2 (test/TestConstImpedance.cpp)
    auto z1 = std::make_shared<Impedance>(one, 1, h);
    CHECK(z1->nFreqs() == nmax && (*z1)[nmax / 2 - 1] == impedance_t(1) && (*z1)[nmax / 2] == impedance_t(0));
    const float scaling = 3;
    ElectricField f1(ps, z1, {0}, 0, h, 1, 1, scaling);
    const float* w = f1.wakePotential();
    ps->syncDevice(PhaseSpace::clCopyDirection::dev2cpu);
    for (meshindex_t x = 0; x < n; x++) CHECK_CLOSE(w[x], ps->getProjection(0)[x] * scaling, golden::ef_roundtrip_tolerance);
    // free-space CSR is causal: a (zero-sum) change of the profile behind the centre leaves the wake
    // potential ahead of it untouched (:175-239).  Fixture scaling: Ib = dt = 1e-3, sigma_z = 3 m,
    // sigma_delta * E0 = 1e3 * 1e9; f0 of the impedance = f_rev.
    auto ps3 = std::make_shared<PhaseSpace>(n, golden::ef_axes[0], golden::ef_axes[1], 3, golden::ef_axes[2],
                                            golden::ef_axes[3], 1, h, 1, 1);
    const double R = 2.99792458e8 / (2 * 3.14159265358979323846 * golden::ef_frev);
    auto zfs = std::shared_ptr<Impedance>(makeImpedance(nmax, h, golden::ef_fmax, R, golden::ef_frev, -1));
    ElectricField f2(ps3, zfs, {0}, 0, h, golden::ef_frev, 1, 1e-3, 1e9, 1e3, 1e-3);
    std::vector<float> w0(f2.wakePotential(), f2.wakePotential() + n);
    ps3->syncDevice(PhaseSpace::clCopyDirection::dev2cpu);
    const meshindex_t cx = n / 2;
    const float* px = ps3->getProjection(0);
    const float delta = golden::ef_change * px[cx];
    const float sign[4] = {-1, 1, 1, -1};
    for (int k = 0; k < 4; k++) {  // profile[cx + k] += sign * delta, by scaling that row of the grid
        const float factor = 1 + sign[k] * delta / px[cx + k];
        for (meshindex_t y = 0; y < n; y++) ps3->getData()[size_t(cx + k) * n + y] *= factor;
    }
    ps3->syncDevice(PhaseSpace::clCopyDirection::cpu2dev);
    ps3->updateXProjection();
    const float* w1 = f2.wakePotential();
    float moved = 0;
    for (meshindex_t x = 0; x < cx; x++) CHECK_CLOSE(w1[x], w0[x], golden::ef_causality_limit);
    for (meshindex_t x = cx; x < n; x++) moved = std::max(moved, std::fabs(w1[x] - w0[x]));
    CHECK(moved > golden::ef_causality_limit);
    CHECK(*f2.updateCSR(0) > 0);
}

static void one_step(cudahptr_t h) {
    // the loop body of main() (src/main.cpp:948-1053) through the mirrored classes, against the
    // same step through the simulation object of the C ABI: the maps are the same kernels
    const meshindex_t n = 64;
    ivs_ring_params p;
    ivs_default_params(&p);
    p.grid_size = n;
    p.f_s = 8.5e3f; p.f_rev = 2.7157e6f; p.harmonic = 184; p.V_RF = 1.4e6; p.R_bend = 5.559; p.bunch_current = 50e-6f;
    ivs_derived d;
    CHECK(ivs_derive(&p, &d) == 0);
    auto t1 = std::make_shared<PhaseSpace>(n, d.qmin, d.qmax, d.bl, d.pmin, d.pmax, d.dE, h, d.Qb, d.Ib);
    t1->updateXProjection();
    t1->normalize();  // src/main.cpp:528-532
    auto t2 = std::make_shared<PhaseSpace>(*t1), t3 = std::make_shared<PhaseSpace>(*t1);
    auto z = std::shared_ptr<Impedance>(makeImpedance(d.padded_bins, h, d.fmax, d.R_bend, d.f_rev, p.gap));
    ElectricField field(t1, z, {0}, d.spacing_bins, h, d.f_rev, (float)d.revolutionpart, d.Ib, p.E0, p.sE, d.dt);
    WakePotentialMap wm(t1, t2, &field, SourceMap::cubic, false, h);
    RFKickMap rfm(t2, t1, d.angle, (float)d.f_RF, SourceMap::cubic, false, h);
    DriftMap drm(t1, t3, {d.slip[0], d.slip[1], d.slip[2]}, (float)p.E0, SourceMap::cubic, false, h);
    FokkerPlanckMap fpm(t3, t1, n, n, FokkerPlanckMap::FPType::full, FokkerPlanckMap::FPTracking::none, d.e1,
                        FokkerPlanckMap::DerivationType::cubic, h);
    std::vector<PhaseSpace::Position> trackme = {{20.5f, 30.25f}, {40.0f, 33.0f}};
    for (int step = 0; step < 10; step++) {
        wm.update();
        t1->integrate();
        wm.apply(); wm.applyToAll(trackme);
        rfm.apply(); rfm.applyToAll(trackme);
        drm.apply(); drm.applyToAll(trackme);
        fpm.apply(); fpm.applyToAll(trackme);
        t1->updateXProjection();
        h->flush();
    }
    t1->syncDevice(PhaseSpace::clCopyDirection::dev2cpu);
    CHECK_CLOSE(t1->integrate(), 1.0, 1e-4);
    for (const auto& q : trackme) CHECK(q.x >= 1 && q.x <= n - 1 && q.y >= 1 && q.y <= n - 1);
    // the same 10 steps by the simulation object (unfused kernel sequence)
    std::vector<float> rf(n), dr(n), fp(size_t(n) * 5), zw(size_t(d.padded_bins) * 2), grid(size_t(n) * n);
    int nz = 0;
    ivs_rf_kick_offsets(&p, &d, rf.data());
    ivs_drift_offsets(&p, &d, dr.data());
    ivs_fokker_planck_weights(n, d.pmin, d.pmax, d.e1, p.fp_type, p.derivation, fp.data());
    ivs_wake_impedance(&p, &d, zw.data(), &nz);
    ivs_initial_grid(&d, n, p.zoom, grid.data());
    ivs_sim_desc sd;
    std::memset(&sd, 0, sizeof(sd));
    sd.n = n; sd.members = 1; sd.it = 4; sd.nmax = d.padded_bins; sd.flags = IVS_SIM_UNFUSED;
    sd.dq = d.dq; sd.dp = d.dp; sd.e1 = d.e1; sd.qmin = d.qmin; sd.pmin = d.pmin;
    sd.rf_offset = rf.data(); sd.drift_offset = dr.data(); sd.fp_weights = fp.data();
    sd.impedance = zw.data(); sd.wakescaling = &d.wakescaling; sd.grid = grid.data();
    ivs_sim* sim = nullptr;
    h->check(ivs_sim_create(h->ctx(), &sd, &sim));
    h->check(ivs_sim_step(sim, 10));
    std::vector<float> out(size_t(n) * n);
    h->check(ivs_sim_download(sim, out.data(), nullptr, nullptr, nullptr));
    ivs_sim_destroy(sim);
    double num = 0, den = 0;
    for (size_t i = 0; i < out.size(); i++) { num += std::fabs(t1->getData()[i] - out[i]); den += std::fabs(out[i]); }
    CHECK_CLOSE(num / den, 0.0, 1e-6);  // same kernels, same tables: only the start normalisation path differs
}

int main(int argc, char** argv) {
    const bool host = argc > 1 && std::strcmp(argv[1], "--host-only") == 0;
    host_only();
    if (host) {
        if (ivs_device_count() == 0) {  // without a GPU the handler must fail loudly
            CHECK(throws<DeviceError>([] { CudaHandler h(0); }));
        }
    } else {
        auto h = std::make_shared<CudaHandler>(0, IVS_ARITH_STRICT);
        kickmaps(h);
        rotationmaps(h);
        phasespace(h);
        electricfield(h);
        one_step(h);
    }
    std::printf("%s: %d checks, %d failures\n", failures ? "FAILED" : "OK", checks, failures);
    return failures ? 1 : 0;
}
