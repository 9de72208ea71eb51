// vfps_b200.hpp -- C++ host side above the C ABI (include/ivs.h, include/ivs_host.h): the classes
// of the reference's hot path with the reference's names, argument meaning and error behaviour,
// every device operation delegated to libivs.so.  Header only; link with -livs.
//
//   reference class (file)                           here
//   OCLH                inc/CL/OpenCLHandler.hpp      vfps::CudaHandler      (cudahptr_t)
//   Ruler<T>            inc/PS/Ruler.hpp              vfps::Ruler<T>
//   PhaseSpace          inc/PS/PhaseSpace.hpp         vfps::PhaseSpace
//   Impedance, makeImpedance  inc/Z/*.hpp             vfps::Impedance, vfps::makeImpedance
//   ElectricField       inc/PS/ElectricField.hpp      vfps::ElectricField
//   SourceMap           inc/SM/SourceMap.hpp          vfps::SourceMap
//   KickMap / DriftMap / RFKickMap / WakeKickMap / WakePotentialMap   inc/SM/*.hpp
//   FokkerPlanckMap / RotationMap / Identity          inc/SM/*.hpp
//
// Deliberate differences, all forced by "many simulations per process, one GPU path":
//   * the grid size is a constructor argument of PhaseSpace (the reference keeps it in process-wide
//     statics, inc/PS/PhaseSpace.hpp:336-356); single bunch only in this C++ mirror (a train of bunches goes through
//     ivs_sim_set_bunch_train of the C ABI, ivs.h);
//   * there is no CPU path: a null handler is std::invalid_argument, a device error is
//     vfps::DeviceError (the reference's cl::Error), never a silent fall-back;
//   * host mirrors are stale until syncDevice(dev2cpu) (the reference's syncCLMem contract,
//     src/PS/PhaseSpace.cpp:381-414); getData()/getProjection() read the mirrors.
#ifndef VFPS_B200_HPP
#define VFPS_B200_HPP

#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "ivs.h"
#include "ivs_host.h"

namespace vfps {

// inc/defines.hpp:31-43: "all of the below types have to be the same"
typedef float meshaxis_t;
typedef float meshdata_t;
typedef float projection_t;
typedef float integral_t;
typedef float timeaxis_t;
typedef float frequency_t;
typedef float csrpower_t;
typedef uint32_t meshindex_t;
typedef std::complex<float> impedance_t;

class DeviceError : public std::runtime_error {
public:
    DeviceError(int code, const std::string& what) : std::runtime_error(what), code(code) {}
    const int code;
};

/// the device handle every class takes (the reference's oclhptr_t): one GPU, one in-order stream
class CudaHandler {
public:
    explicit CudaHandler(int device = 0, int arith = IVS_ARITH_FMA) {
        const int st = ivs_ctx_create(device, arith, &_ctx);
        if (st != IVS_OK) throw DeviceError(st, "ivs_ctx_create failed (no B200 / no CUDA driver?)");
    }
    ~CudaHandler() noexcept { if (_ctx) ivs_ctx_destroy(_ctx); }
    CudaHandler(const CudaHandler&) = delete;
    CudaHandler& operator=(const CudaHandler&) = delete;
    ivs_ctx* ctx() const { return _ctx; }
    /// status code -> the reference's exceptions
    void check(int st) const {
        if (st == IVS_OK) return;
        const std::string msg = ivs_last_error(_ctx);
        if (st == IVS_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
        throw DeviceError(st, msg);
    }
    void flush() const { check(ivs_flush(_ctx)); }    // OCLH::flush
    void finish() const { check(ivs_sync(_ctx)); }    // OCLH::finish
    void* alloc(size_t bytes) const { void* p = nullptr; check(ivs_malloc(_ctx, bytes, &p)); return p; }
    void free(void* p) const noexcept { if (p) ivs_free(_ctx, p); }
    void upload(void* dst, const void* src, size_t bytes) const { check(ivs_upload(_ctx, dst, src, bytes)); }
    void download(void* dst, const void* src, size_t bytes) const { check(ivs_download(_ctx, dst, src, bytes)); }
    void copy(void* dst, const void* src, size_t bytes) const { check(ivs_copy(_ctx, dst, src, bytes)); }
private:
    ivs_ctx* _ctx = nullptr;
};
typedef std::shared_ptr<CudaHandler> cudahptr_t;

inline cudahptr_t require(cudahptr_t h) {
    if (!h) throw std::invalid_argument("vfps_b200: a CudaHandler is required (there is no CPU path)");
    return h;
}

/// device buffer owned by the host object that allocated it
class DeviceBuffer {
public:
    DeviceBuffer() = default;
    DeviceBuffer(cudahptr_t h, size_t bytes) : _h(h), _p(h->alloc(bytes)), _bytes(bytes) {}
    ~DeviceBuffer() noexcept { if (_h) _h->free(_p); }
    DeviceBuffer(DeviceBuffer&& o) noexcept : _h(std::move(o._h)), _p(o._p), _bytes(o._bytes) { o._p = nullptr; }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept { std::swap(_h, o._h); std::swap(_p, o._p); std::swap(_bytes, o._bytes); return *this; }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    void* get() const { return _p; }
    size_t bytes() const { return _bytes; }
private:
    cudahptr_t _h;
    void* _p = nullptr;
    size_t _bytes = 0;
};

/// inc/PS/Ruler.hpp:25-133
template <class T>
class Ruler {
public:
    Ruler(meshindex_t steps, T min, T max, std::map<std::string, T> scale = {})
        : _steps(steps), _max(max), _min(min), _delta((max - min) / static_cast<T>(steps - 1)),
          _scale(std::move(scale)),
          _zerobin(((min + max) / (min - max) + 1) * (steps - 1) / 2) {
        if (min >= max) throw std::invalid_argument("Tried to set up Ruler with min >= max.");
        _scale["Meter"];  // keys exist like in the reference's use
    }
    T at(meshindex_t d) const { return _min + (static_cast<T>(d) * _delta); }
    T operator[](meshindex_t d) const { return at(d); }
    T max() const { return _max; }
    T min() const { return _min; }
    meshindex_t steps() const { return _steps; }
    T delta() const { return _delta; }
    T zerobin() const { return _zerobin; }
    T scale(const std::string& unit) const { return _scale.at(unit); }
    T length() const { return _max - _min; }
private:
    meshindex_t _steps;
    T _max, _min, _delta;
    std::map<std::string, T> _scale;
    T _zerobin;
};

/// inc/PS/PhaseSpace.hpp -- f[x][y], y (energy) contiguous, one bunch
class PhaseSpace {
public:
    struct Position { meshaxis_t x, y; };
    enum class clCopyDirection { cpu2dev, dev2cpu };

    /// the reference's (qmin, qmax, qscale, pmin, pmax, pscale, oclh, charge, current, filling,
    /// zoom, data) constructor with the grid size in front (src/PS/PhaseSpace.cpp:17-66,151-166)
    PhaseSpace(meshindex_t n, meshaxis_t qmin, meshaxis_t qmax, double qscale, meshaxis_t pmin,
               meshaxis_t pmax, double pscale, cudahptr_t cudah, double beam_charge, double beam_current,
               const std::vector<integral_t>& filling = {1}, double zoom = 1,
               const meshdata_t* data = nullptr)
        : charge(beam_charge), current(beam_current), _n(n),
          _axis{{Ruler<meshaxis_t>(n, qmin, qmax, {{"Meter", (meshaxis_t)qscale}}),
                 Ruler<meshaxis_t>(n, pmin, pmax, {{"ElectronVolt", (meshaxis_t)pscale}})}},
          _cudah(require(cudah)), _data(size_t(n) * n), _ws(n),
          _projection{{std::vector<projection_t>(n), std::vector<projection_t>(n)}},
          data_buf(_cudah, size_t(n) * n * sizeof(meshdata_t)), projectionX_buf(_cudah, n * sizeof(projection_t)),
          projectionY_buf(_cudah, n * sizeof(projection_t)), integral_buf(_cudah, sizeof(integral_t)),
          ws_buf(_cudah, n * sizeof(meshdata_t)), moments_buf(_cudah, 3 * sizeof(float)) {
        if (filling.size() != 1) throw std::invalid_argument("vfps_b200: single bunch only");
        double sum = 0;
        for (integral_t f : filling) sum += f;
        if (std::abs(sum - 1) > 1e-5) throw std::invalid_argument("Please provide normalized filling pattern.");
        _filling_set = filling[0];
        ivs_simpson_weights(n, _axis[0].delta(), _ws.data());  // the same weights serve both axes
        _cudah->upload(ws_buf.get(), _ws.data(), n * sizeof(meshdata_t));
        if (data) {
            std::copy(data, data + size_t(n) * n, _data.begin());
            syncDevice(clCopyDirection::cpu2dev);
        } else {  // Gaussian product, normalised (gaus + createFromProjections)
            std::vector<projection_t> p0(n), p1(n);
            ivs_gaussian_projection(n, qmin, qmax, zoom, p0.data());
            ivs_gaussian_projection(n, pmin, pmax, zoom, p1.data());
            for (meshindex_t x = 0; x < n; x++)
                for (meshindex_t y = 0; y < n; y++) _data[size_t(x) * n + y] = p0[x] * p1[y];
            syncDevice(clCopyDirection::cpu2dev);
            updateXProjection();
            integrate();
            normalize();
        }
        updateXProjection();
        updateYProjection();
        integrate();
    }
    PhaseSpace(const PhaseSpace& o)
        : PhaseSpace(o._n, o._axis[0].min(), o._axis[0].max(), o._axis[0].scale("Meter"), o._axis[1].min(),
                     o._axis[1].max(), o._axis[1].scale("ElectronVolt"), o._cudah, o.charge, o.current,
                     {o._filling_set}, 1, o.syncedData()) {}

    const double charge, current;

    /// host mirror (valid after syncDevice(dev2cpu))
    meshdata_t* getData() { return _data.data(); }
    const meshdata_t* getData() const { return _data.data(); }
    meshindex_t nMeshCells(uint_fast8_t) const { return _n; }
    meshindex_t nMeshCells() const { return _n * _n; }
    const Ruler<meshaxis_t>* getAxis(uint_fast8_t x) const { return &_axis[x]; }
    meshaxis_t getDelta(uint_fast8_t x) const { return _axis[x].delta(); }
    meshaxis_t getMax(uint_fast8_t x) const { return _axis[x].max(); }
    meshaxis_t getMin(uint_fast8_t x) const { return _axis[x].min(); }
    meshaxis_t getScale(uint_fast8_t x, const std::string& unit) const { return _axis[x].scale(unit); }
    meshaxis_t x(meshaxis_t q) const { return clampidx((q - _axis[0].min()) / _axis[0].delta()); }
    meshaxis_t y(meshaxis_t p) const { return clampidx((p - _axis[1].min()) / _axis[1].delta()); }

    /// src/PS/PhaseSpace.cpp:291-323 / :325-340 / :200-239 / :342-372 / :241-289, on the device
    void updateXProjection() { chk(ivs_project_x(c(), data_buf.get(), _n, ws_buf.get(), projectionX_buf.get())); }
    void updateYProjection() { chk(ivs_project_y(c(), data_buf.get(), _n, ws_buf.get(), projectionY_buf.get())); }
    integral_t integrate() {
        chk(ivs_integrate(c(), projectionX_buf.get(), _n, ws_buf.get(), integral_buf.get()));
        _cudah->download(&_integral, integral_buf.get(), sizeof(integral_t));
        _cudah->finish();
        return _integral;
    }
    integral_t getIntegral() const { return _integral; }
    integral_t normalize() {
        chk(ivs_normalize(c(), data_buf.get(), _n, _filling_set, integral_buf.get()));
        return _integral;
    }
    void variance(uint_fast8_t axis) {
        const DeviceBuffer& proj = axis == 0 ? projectionX_buf : projectionY_buf;
        chk(ivs_moments(c(), proj.get(), _n, _axis[axis].min(), _axis[axis].delta(), integral_buf.get(), moments_buf.get()));
        float m[3];
        _cudah->download(m, moments_buf.get(), sizeof(m));
        _cudah->finish();
        _moment[axis] = {{m[0], m[1], m[2]}};
    }
    void average(uint_fast8_t axis) { variance(axis); }
    meshaxis_t getMoment(uint_fast8_t axis, uint_fast8_t m) const { return _moment[axis][m]; }  // 0 mean, 1 variance
    meshaxis_t getBunchLength() const { return _moment[0][2]; }
    meshaxis_t getEnergySpread() const { return _moment[1][2]; }
    const projection_t* getProjection(uint_fast8_t x) const { return _projection[x].data(); }

    /// PhaseSpace::syncCLMem (src/PS/PhaseSpace.cpp:381-414)
    void syncDevice(clCopyDirection dir) {
        if (dir == clCopyDirection::cpu2dev) {
            _cudah->upload(data_buf.get(), _data.data(), _data.size() * sizeof(meshdata_t));
        } else {
            _cudah->download(_data.data(), data_buf.get(), _data.size() * sizeof(meshdata_t));
            _cudah->download(_projection[0].data(), projectionX_buf.get(), _n * sizeof(projection_t));
            _cudah->download(_projection[1].data(), projectionY_buf.get(), _n * sizeof(projection_t));
            _cudah->download(&_integral, integral_buf.get(), sizeof(integral_t));
            _cudah->finish();
        }
    }
    cudahptr_t handler() const { return _cudah; }

private:
    const meshindex_t _n;
    std::array<Ruler<meshaxis_t>, 2> _axis;
    cudahptr_t _cudah;
    std::vector<meshdata_t> _data;
    std::vector<meshdata_t> _ws;
    std::array<std::vector<projection_t>, 2> _projection;
    integral_t _integral = 0, _filling_set = 1;
    std::array<std::array<meshaxis_t, 3>, 2> _moment{};
    ivs_ctx* c() const { return _cudah->ctx(); }
    void chk(int st) const { _cudah->check(st); }
    meshaxis_t clampidx(meshaxis_t v) const { return std::max(0.0f, std::min(v, meshaxis_t(_n - 1))); }
    const meshdata_t* syncedData() const {
        const_cast<PhaseSpace*>(this)->syncDevice(clCopyDirection::dev2cpu);
        return _data.data();
    }
public:
    // device buffers the maps and the field reach into (reference: data_buf, projectionX_clbuf, ...)
    DeviceBuffer data_buf, projectionX_buf, projectionY_buf, integral_buf, ws_buf, moments_buf;
};

/// inc/Z/Impedance.hpp -- nfreqs complex<float>, upper half zero, table uploaded once
class Impedance {
public:
    Impedance(const std::vector<impedance_t>& z, frequency_t f_max, cudahptr_t cudah)
        : _nfreqs(z.size()), _axis(z.size(), 0, f_max, {{"Hertz", 1}}), _data(z), _cudah(require(cudah)),
          data_buf(_cudah, z.size() * sizeof(impedance_t)) {
        syncDevice();
    }
    Impedance(size_t nfreqs, frequency_t f_max, cudahptr_t cudah)
        : Impedance(std::vector<impedance_t>(nfreqs, 0), f_max, cudah) {}
    const impedance_t* data() const { return _data.data(); }
    const impedance_t operator[](size_t n) const { return _data[n]; }
    size_t nFreqs() const { return _nfreqs; }
    const Ruler<frequency_t>* getRuler() const { return &_axis; }
    Impedance& operator+=(const Impedance& rhs) {  // src/Z/Impedance.cpp:67-76
        for (size_t i = 0; i < _nfreqs; i++) _data[i] += rhs._data[i];
        syncDevice();
        return *this;
    }
    void syncDevice() { _cudah->upload(data_buf.get(), _data.data(), _data.size() * sizeof(impedance_t)); }
    static constexpr double Z0 = 376.730313461;  // inc/Z/Impedance.hpp:94
private:
    size_t _nfreqs;
    Ruler<frequency_t> _axis;
    std::vector<impedance_t> _data;
    cudahptr_t _cudah;
public:
    DeviceBuffer data_buf;
};

/// makeImpedance (src/Z/ImpedanceFactory.cpp:21-84) without the file contribution
inline std::unique_ptr<Impedance> makeImpedance(size_t nfreqs, cudahptr_t cudah, frequency_t fmax, double R_bend,
                                                double frev, double gap, bool use_csr = true, double s = 0,
                                                double xi = 0, double collimator = 0) {
    std::vector<impedance_t> z(nfreqs);
    int nonzero = 0;
    const int st = ivs_make_impedance(nfreqs, fmax, R_bend, frev, gap, use_csr, s, xi, collimator,
                                      reinterpret_cast<float*>(z.data()), &nonzero);
    if (st != 0) throw std::invalid_argument("makeImpedance: invalid parameters");
    return std::unique_ptr<Impedance>(new Impedance(z, fmax, cudah));
}

/// inc/PS/ElectricField.hpp -- wake potential and CSR power of the bunch profile
class ElectricField {
public:
    /// (ps, impedance, bucketnumbers, spacing_bins, oclh, f_rev, revolutionpart, wakescaling)
    ElectricField(std::shared_ptr<PhaseSpace> ps, std::shared_ptr<Impedance> impedance,
                  const std::vector<uint32_t>& bucketnumbers, meshindex_t /*spacing_bins*/, cudahptr_t cudah,
                  double f_rev, meshaxis_t /*revolutionpart*/ = 1, meshaxis_t wakescaling = 0.0)
        : volts(ps->getAxis(1)->delta() * ps->getScale(1, "ElectronVolt")), _nmax(impedance->nFreqs()),
          _phasespace(ps), _impedance(impedance), _cudah(require(cudah)), _n(ps->nMeshCells(0)),
          _wakescaling(wakescaling / _nmax), _f_rev(f_rev), _wakepotential(_n), _csrspectrum(_nmax),
          wakepotential_buf(_cudah, _n * sizeof(meshaxis_t)), spectrum_buf(_cudah, _nmax * sizeof(csrpower_t)),
          power_buf(_cudah, sizeof(csrpower_t)) {
        if (bucketnumbers.size() != 1) throw std::invalid_argument("vfps_b200: single bunch only");
    }
    /// (…, revolutionpart, Ib, E0, sigma_delta, dt): scaling of inc/PS/ElectricField.hpp:81-103
    ElectricField(std::shared_ptr<PhaseSpace> ps, std::shared_ptr<Impedance> impedance,
                  const std::vector<uint32_t>& bucketnumbers, meshindex_t spacing_bins, cudahptr_t cudah,
                  double f_rev, meshaxis_t revolutionpart, double Ib, double E0, double sigma_delta, double dt)
        : ElectricField(ps, impedance, bucketnumbers, spacing_bins, cudah, f_rev, revolutionpart,
                        meshaxis_t(Ib * dt * 2.99792458e8 / ps->getScale(0, "Meter") /
                                   (ps->getDelta(1) * sigma_delta * E0))) {}

    const meshaxis_t volts;
    size_t getNMax() const { return _nmax; }
    std::shared_ptr<Impedance> getImpedance() const { return _impedance; }
    /// src/PS/ElectricField.cpp:186-253; reads the phase space's X projection on the device
    meshaxis_t* wakePotential() {
        chk(ivs_wake_potential(c(), _phasespace->projectionX_buf.get(), _n, _impedance->data_buf.get(),
                               (uint32_t)_nmax, _wakescaling, wakepotential_buf.get(), nullptr));
        _cudah->download(_wakepotential.data(), wakepotential_buf.get(), _n * sizeof(meshaxis_t));
        _cudah->finish();
        return _wakepotential.data();
    }
    const std::vector<meshaxis_t>& getWakePotentials() const { return _wakepotential; }
    /// src/PS/ElectricField.cpp:133-184
    const csrpower_t* updateCSR(frequency_t cutoff_frequency) {
        const float f_scale = 2.99792458e8f / _phasespace->getScale(0, "Meter");
        chk(ivs_csr_power(c(), _phasespace->projectionX_buf.get(), _n, _impedance->data_buf.get(), (uint32_t)_nmax,
                          _phasespace->getDelta(0), f_scale, cutoff_frequency, spectrum_buf.get(), power_buf.get()));
        _cudah->download(_csrspectrum.data(), spectrum_buf.get(), _nmax * sizeof(csrpower_t));
        _cudah->download(&_csrintensity, power_buf.get(), sizeof(csrpower_t));
        _cudah->finish();
        return &_csrintensity;
    }
    const csrpower_t* getCSRPower() const { return &_csrintensity; }
    const csrpower_t* getCSRSpectrum() const { return _csrspectrum.data(); }
private:
    const size_t _nmax;
    std::shared_ptr<PhaseSpace> _phasespace;
    std::shared_ptr<Impedance> _impedance;
    cudahptr_t _cudah;
    const meshindex_t _n;
    const meshaxis_t _wakescaling;
    const double _f_rev;
    std::vector<meshaxis_t> _wakepotential;
    std::vector<csrpower_t> _csrspectrum;
    csrpower_t _csrintensity = 0;
    ivs_ctx* c() const { return _cudah->ctx(); }
    void chk(int st) const { _cudah->check(st); }
public:
    DeviceBuffer wakepotential_buf, spectrum_buf, power_buf;
};

/// inc/SM/SourceMap.hpp
class SourceMap {
public:
    enum InterpolationType : uint_fast8_t { none = 1, linear = 2, quadratic = 3, cubic = 4 };
    SourceMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, cudahptr_t cudah)
        : _in(in), _out(out), _cudah(require(cudah)), _n(in->nMeshCells(0)) {
        if (out->nMeshCells(0) != _n) throw std::invalid_argument("SourceMap: in and out differ in size");
    }
    virtual ~SourceMap() noexcept = default;
    virtual void apply() = 0;
    virtual void applyTo(PhaseSpace::Position& pos) const = 0;
    /// src/SM/SourceMap.cpp:111-117
    void applyToAll(std::vector<PhaseSpace::Position>& particles) const {
        for (auto& p : particles) applyTo(p);
    }
protected:
    std::shared_ptr<PhaseSpace> _in, _out;
    cudahptr_t _cudah;
    const meshindex_t _n;
    ivs_ctx* c() const { return _cudah->ctx(); }
    void chk(int st) const { _cudah->check(st); }
};

/// inc/SM/Identity.hpp
class Identity : public SourceMap {
public:
    Identity(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, meshindex_t, meshindex_t, cudahptr_t cudah)
        : SourceMap(in, out, cudah) {}
    void apply() override { chk(ivs_identity(c(), _in->data_buf.get(), _out->data_buf.get(), _n)); }
    void applyTo(PhaseSpace::Position&) const override {}
};

/// inc/SM/KickMap.hpp -- 1-D interpolated shift along x (per-y offset) or y (per-x offset)
class KickMap : public SourceMap {
public:
    enum class Axis : bool { x = 0, y = 1 };
    KickMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, InterpolationType it,
            bool /*interpol_clamp: the reference's CPU path never clamps kick maps (src/SM/KickMap.cpp:194-271)*/,
            Axis kd, cudahptr_t cudah)
        : SourceMap(in, out, cudah), _kickdirection(kd), _it(it), _offset(_n, 0),
          _offset_buf(_cudah, _n * sizeof(meshaxis_t)) {
        syncOffset();
    }
    const meshaxis_t* getForce() const { return _offset.data(); }
    void swapOffset(std::vector<meshaxis_t>& offset) {  // inc/SM/KickMap.hpp:54
        if (offset.size() != _n) throw std::invalid_argument("KickMap::swapOffset: wrong size");
        _offset.swap(offset);
        syncOffset();
    }
    void apply() override {  // src/SM/KickMap.cpp:194-271
        if (_kickdirection == Axis::x) chk(ivs_kick_x(c(), _in->data_buf.get(), _out->data_buf.get(), _n, _offset_buf.get(), _it));
        else chk(ivs_kick_y(c(), _in->data_buf.get(), _out->data_buf.get(), _n, _offset_buf.get(), _it));
    }
    void applyTo(PhaseSpace::Position& pos) const override {  // src/SM/KickMap.cpp:280-303
        meshaxis_t& par = _kickdirection == Axis::x ? pos.x : pos.y;
        const meshaxis_t perp = _kickdirection == Axis::x ? pos.y : pos.x;
        meshaxis_t ip;
        const meshaxis_t fr = std::modf(perp, &ip);
        const meshindex_t k = static_cast<meshindex_t>(ip);
        if (k + 1 < _n) par -= (1 - fr) * _offset[k] + fr * _offset[k + 1];
        par = std::max(meshaxis_t(1), std::min(par, meshaxis_t(_n - 1)));
    }
protected:
    void syncOffset() { _cudah->upload(_offset_buf.get(), _offset.data(), _n * sizeof(meshaxis_t)); }
    const Axis _kickdirection;
    const InterpolationType _it;
    std::vector<meshaxis_t> _offset;
    DeviceBuffer _offset_buf;
};

/// inc/SM/DriftMap.hpp
class DriftMap : public KickMap {
public:
    DriftMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, const std::vector<meshaxis_t>& slip,
             meshaxis_t E0, InterpolationType it, bool interpol_clamp, cudahptr_t cudah)
        : KickMap(in, out, it, interpol_clamp, Axis::x, cudah) {
        ivs_drift_slip(_n, in->getMin(0), in->getMax(0), in->getMin(1), in->getMax(1),
                       in->getScale(1, "ElectronVolt"), slip.data(), (uint32_t)slip.size(), E0, _offset.data());
        syncOffset();
    }
};

/// inc/SM/RFKickMap.hpp
class RFKickMap : public KickMap {
public:
    RFKickMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, meshaxis_t angle, frequency_t f_RF,
              InterpolationType it, bool interpol_clamp, cudahptr_t cudah)
        : KickMap(in, out, it, interpol_clamp, Axis::y, cudah) {
        ivs_rf_kick_linear(_n, in->getMin(0), in->getMax(0), in->getScale(0, "Meter"), angle, f_RF, _offset.data());
        syncOffset();
    }
    RFKickMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, timeaxis_t revolutionpart,
              meshaxis_t V_RF, frequency_t f_RF, meshaxis_t V0, InterpolationType it, bool interpol_clamp,
              cudahptr_t cudah)
        : KickMap(in, out, it, interpol_clamp, Axis::y, cudah) {
        ivs_rf_kick_sine(_n, in->getMin(0), in->getMax(0), in->getMin(1), in->getMax(1), in->getScale(0, "Meter"),
                         in->getScale(1, "ElectronVolt"), revolutionpart, V_RF, f_RF, V0, _offset.data());
        syncOffset();
    }
};

/// inc/SM/WakeKickMap.hpp + inc/SM/WakePotentialMap.hpp
class WakeKickMap : public KickMap {
public:
    WakeKickMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, InterpolationType it,
                bool interpol_clamp, cudahptr_t cudah)
        : KickMap(in, out, it, interpol_clamp, Axis::y, cudah) {}
    virtual void update() = 0;
};
class WakePotentialMap : public WakeKickMap {
public:
    WakePotentialMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, ElectricField* field,
                     InterpolationType it, bool interpol_clamp, cudahptr_t cudah)
        : WakeKickMap(in, out, it, interpol_clamp, cudah), _field(field) {}
    void update() override {  // src/SM/WakePotentialMap.cpp:34-50: device to device, host mirror refreshed
        const meshaxis_t* w = _field->wakePotential();
        std::copy(w, w + _n, _offset.begin());
        _cudah->copy(_offset_buf.get(), _field->wakepotential_buf.get(), _n * sizeof(meshaxis_t));
    }
private:
    ElectricField* _field;
};

/// inc/SM/FokkerPlanckMap.hpp
class FokkerPlanckMap : public SourceMap {
public:
    enum class FPType : uint_fast8_t { none = 0, damping_only = 1, diffusion_only = 2, full = 3 };
    enum class FPTracking : uint_fast8_t { none = 0, approximation1 = 1, approximation2 = 2, stochastic = 3 };
    enum DerivationType : uint_fast8_t { two_sided = 3, cubic = 4 };
    FokkerPlanckMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, meshindex_t /*xsize*/,
                    meshindex_t /*ysize*/, FPType fptype, FPTracking fptrack, timeaxis_t e1, DerivationType dt,
                    cudahptr_t cudah, uint64_t seed = 0)
        : SourceMap(in, out, cudah), _dampdecr(e1), _fptrack(fptrack), _w5(size_t(_n) * 5),
          _w5_buf(_cudah, size_t(_n) * 5 * sizeof(float)), _prng(seed),
          _normdist(0, std::sqrt(2 * e1) / in->getDelta(1)) {
        chk(ivs_fokker_planck_weights(_n, in->getMin(1), in->getMax(1), e1, (uint32_t)fptype, (uint32_t)dt, _w5.data()));
        _cudah->upload(_w5_buf.get(), _w5.data(), _w5.size() * sizeof(float));
    }
    void apply() override {  // src/SM/FokkerPlanckMap.cpp:170-213
        chk(ivs_fokker_planck(c(), _in->data_buf.get(), _out->data_buf.get(), _n, _w5_buf.get()));
    }
    void applyTo(PhaseSpace::Position& pos) const override {  // src/SM/FokkerPlanckMap.cpp:214-268
        switch (_fptrack) {
        case FPTracking::none: break;
        case FPTracking::approximation1: {
            meshindex_t yi = std::min<meshindex_t>(static_cast<meshindex_t>(std::floor(pos.y)), _n - 1);
            meshaxis_t offset = 0;
            for (int k = 0; k < 5; k++) {
                const float w = _w5[size_t(yi) * 5 + k];
                if (w != 0.0f) offset += float(2 - k) * w;
            }
            pos.y = std::max(meshaxis_t(1), std::min(pos.y + offset, meshaxis_t(_n - 1)));
            break;
        }
        case FPTracking::approximation2: {  // reads the input grid's host mirror
            const meshindex_t xi = std::min<meshindex_t>(static_cast<meshindex_t>(std::floor(pos.x)), _n - 1);
            const int yi = std::max(0, std::min<int>(static_cast<int>(std::floor(pos.y)), int(_n) - 1));
            const meshdata_t* row = _in->getData() + size_t(xi) * _n;
            meshaxis_t offset = 0, chg = 0;
            for (int k = 0; k < 5; k++) {
                const float w = _w5[size_t(yi) * 5 + k];
                const int idx = yi + k - 2;
                if (w != 0.0f && idx >= 0 && idx < int(_n)) {
                    const meshaxis_t cw = row[idx] * w;
                    chg += cw;
                    offset += cw * meshaxis_t(idx - yi);
                }
            }
            pos.y = std::max(meshaxis_t(1), std::min(pos.y + offset / chg, meshaxis_t(_n - 1)));
            break;
        }
        case FPTracking::stochastic:  // seeded here (the reference seeds from std::random_device)
            pos.y -= pos.y * _dampdecr + _normdist(_prng);
            break;
        }
    }
private:
    const timeaxis_t _dampdecr;
    const FPTracking _fptrack;
    std::vector<float> _w5;
    DeviceBuffer _w5_buf;
    mutable std::mt19937 _prng;
    mutable std::normal_distribution<meshaxis_t> _normdist;
};

/// inc/SM/RotationMap.hpp
class RotationMap : public SourceMap {
public:
    RotationMap(std::shared_ptr<PhaseSpace> in, std::shared_ptr<PhaseSpace> out, meshindex_t /*xsize*/,
                meshindex_t /*ysize*/, meshaxis_t angle, InterpolationType it, bool interpol_clamped,
                meshindex_t rotmapsize, cudahptr_t cudah)
        : SourceMap(in, out, cudah), _angle(angle), _it(it), _clamp(interpol_clamped),
          _cos_dt(std::cos(-angle)), _sin_dt(std::sin(-angle)) {
        // src/SM/RotationMap.cpp:83-87: clamping needs the full map and cubic interpolation
        if (interpol_clamped && (rotmapsize == 0 || it != cubic))
            throw std::invalid_argument("Clamped interpolation is only implemented for cubic interpolation with a full rotation map.");
    }
    void apply() override {  // src/SM/RotationMap.cpp:97-176
        chk(ivs_rotate(c(), _in->data_buf.get(), _out->data_buf.get(), _n, _angle, _in->getMin(0), _in->getMax(0),
                       _in->getMin(1), _in->getMax(1), _it, _clamp));
    }
    void applyTo(PhaseSpace::Position& pos) const override {  // src/SM/RotationMap.cpp:178-188
        const meshaxis_t zx = _in->getAxis(0)->zerobin(), zy = _in->getAxis(1)->zerobin();
        const meshaxis_t dx = pos.x - zx, dy = pos.y - zy;
        pos.x = _cos_dt * dx + _sin_dt * dy + zx;
        pos.y = _cos_dt * dy - _sin_dt * dx + zy;
    }
private:
    const meshaxis_t _angle;
    const InterpolationType _it;
    const bool _clamp;
    const meshaxis_t _cos_dt, _sin_dt;
};

}  // namespace vfps
#endif  // VFPS_B200_HPP
