// ivs_field.cu -- ElectricField on the device: hand-written shared-memory FFT convolution of
// the bunch profile with the impedance (wake potential) and the CSR spectrum / power.
//
// Replaces the reference's clFFT / FFTW plans (src/PS/ElectricField.cpp:73-78,300-306;
// src/FFTWWrapper.cpp:61-107) and the kernels `wakeloss` (:308-316) and `scalewp` (:324-332):
// pad -> r2c -> x Z -> c2r -> scale runs as ONE kernel, one CTA per profile, everything in
// shared memory.  A real transform of length nmax is done as a complex Stockham radix-2 FFT
// of length nmax/2 plus the usual even/odd split, so nmax = 16384 needs 2 x 64 KB of smem.
// Twiddles come from a table computed once per context in double precision.
#include <cmath>
#include <vector>

#include "ivs_common.cuh"

namespace {

constexpr int kFftThreads = 1024;

__device__ __forceinline__ float2 cmul_ref(float2 a, float2 b) {
    // std::complex<float> product as the reference's `Z[i]*F[i]` (src/PS/ElectricField.cpp:230)
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)),
                       __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// Complex FFT of length m (power of two) in shared memory, Stockham autosort radix 2.
// tw[i] = exp(-2 pi i * i / (2m)), i < m.  inverse => conjugated twiddles.  Returns the
// buffer that holds the result (a or b).  All threads of the CTA must call it.
__device__ float2* smem_fft(float2* a, float2* b, uint32_t m, const float2* __restrict__ tw,
                            bool inverse) {
    const uint32_t half = m >> 1;
    for (uint32_t p = 1; p < m; p <<= 1) {
        const uint32_t tstride = half / p;  // twiddle index step: k * (2m / (2p)) / 2 ... see below
        for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
            const uint32_t k = i & (p - 1);
            const uint32_t j = ((i - k) << 1) + k;
            // exp(-i pi k / p) = exp(-2 pi i * (k * m / p) / (2m))
            float2 w = __ldg(&tw[k * tstride * 2]);
            if (inverse) w.y = -w.y;
            const float2 u0 = a[i];
            const float2 u1 = cmul(w, a[i + half]);
            b[j] = cadd(u0, u1);
            b[j + p] = csub(u0, u1);
        }
        __syncthreads();
        float2* t = a; a = b; b = t;
    }
    return a;
}

struct field_args {
    const float* profile;      // [members][n]
    const float2* impedance;   // [nmax] (shared by all members)
    const float2* tw;          // [nmax/2]
    const float* scaling;      // per member (device) or nullptr -> scaling_scalar
    float scaling_scalar;
    float* wake;               // [members][n] or nullptr
    float2* formfactor;        // [members][nmax/2+1] or nullptr
    // CSR power
    float* spectrum;           // [members][nmax] or nullptr
    float* power;              // [members] or nullptr
    float dq, f_scale, cutoff;
    uint32_t n, nmax;
    int mode;                  // 0 = wake potential, 1 = CSR power
};

__global__ void __launch_bounds__(kFftThreads)
field_kernel(field_args g) {
    extern __shared__ float2 fsm[];
    __shared__ float red[32];
    const uint32_t n = g.n, nmax = g.nmax, m = nmax >> 1;
    float2* a = fsm;
    float2* b = fsm + m;
    const uint32_t member = blockIdx.x;
    const float* prof = g.profile + (size_t)member * n;

    // padBunchProfiles (src/PS/ElectricField.cpp:255-263): profile in bins 0..n-1, zeros after;
    // packed as z[j] = x[2j] + i x[2j+1]
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) {
        const uint32_t i0 = 2 * j, i1 = 2 * j + 1;
        a[j] = make_float2(i0 < n ? prof[i0] : 0.0f, i1 < n ? prof[i1] : 0.0f);
    }
    __syncthreads();
    float2* z = smem_fft(a, b, m, g.tw, false);
    float2* other = (z == a) ? b : a;

    // even/odd split: X[k] = E[k] + W^k O[k], E = (Z[k]+conj Z[m-k])/2, O = (Z[k]-conj Z[m-k])/(2i)
    auto spectrum_at = [&](uint32_t k, float2 zk, float2 zmk) -> float2 {
        const float2 e = make_float2(0.5f * (zk.x + zmk.x), 0.5f * (zk.y - zmk.y));
        const float2 o = make_float2(0.5f * (zk.y + zmk.y), -0.5f * (zk.x - zmk.x));
        const float2 w = (k < m) ? __ldg(&g.tw[k]) : make_float2(-1.0f, 0.0f);
        return cadd(e, cmul(w, o));
    };

    if (g.mode == 1) {
        // ElectricField::updateCSR (src/PS/ElectricField.cpp:150-180): only bins 0..nmax/2 of the
        // form factor are non-zero (r2c output), so only they contribute
        const float delta_f = __fdiv_rn(__fdiv_rn(1.0f, g.dq), (float)(nmax - 1));
        const float renorm0 = __fmul_rn(g.dq, g.dq);
        float acc = 0.0f;
        float* spec = g.spectrum ? g.spectrum + (size_t)member * nmax : nullptr;
        for (uint32_t k = threadIdx.x; k < nmax; k += blockDim.x) {
            float s = 0.0f;
            if (k <= m) {
                const float2 zk = z[k % m], zmk = z[(m - k) % m];
                const float2 x = spectrum_at(k, zk, zmk);
                float renorm = renorm0;
                if (g.cutoff > 0.0f) {
                    const float fk = __fmul_rn((float)k, delta_f);  // Ruler::at(k), min = 0
                    const double r = (double)__fdiv_rn(__fmul_rn(g.f_scale, fk), g.cutoff);
                    renorm = (float)((double)renorm * (1.0 - exp(-(r * r))));
                }
                const float nrm = __fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y));
                s = __fmul_rn(__fmul_rn(renorm, g.impedance[k].x), nrm);
            }
            if (spec) spec[k] = s;
            acc = __fadd_rn(acc, __fmul_rn(delta_f, s));
        }
        // block reduction (fixed order)
        for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
            for (int o = 16; o > 0; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
            if (threadIdx.x == 0 && g.power) g.power[member] = t;
        }
        return;
    }

    // wake potential: Y[k] = Z[k] F[k] for k < nmax/2 (src/PS/ElectricField.cpp:229-231), Nyquist
    // bin 0; then the inverse split  z'[k] = (Y[k]+conj Y[m-k]) + i W^-k (Y[k]-conj Y[m-k])
    float2* ff = g.formfactor ? g.formfactor + (size_t)member * (m + 1) : nullptr;
    for (uint32_t k = threadIdx.x; k <= m / 2; k += blockDim.x) {
        const uint32_t mk = m - k;
        const float2 zk = z[k], zmk = z[mk % m];
        float2 xk = spectrum_at(k, zk, zmk);
        float2 xmk = spectrum_at(mk, zmk, zk);
        if (ff) {
            ff[k] = xk;
            ff[mk] = xmk;
        }
        float2 yk = cmul_ref(g.impedance[k], xk);
        float2 ymk = (mk < m) ? cmul_ref(g.impedance[mk], xmk) : make_float2(0.0f, 0.0f);
        if (k == 0) yk.y = 0.0f;  // c2r ignores the imaginary part of bin 0 (and Nyquist is 0)
        // z'[k]
        {
            const float2 ee = make_float2(yk.x + ymk.x, yk.y - ymk.y);
            const float2 d = make_float2(yk.x - ymk.x, yk.y + ymk.y);
            float2 w = (k < m) ? __ldg(&g.tw[k]) : make_float2(-1.0f, 0.0f);
            w.y = -w.y;
            const float2 oo = cmul(w, d);
            other[k] = make_float2(ee.x - oo.y, ee.y + oo.x);
        }
        if (mk != k && mk < m) {
            const float2 ee = make_float2(ymk.x + yk.x, ymk.y - yk.y);
            const float2 d = make_float2(ymk.x - yk.x, ymk.y + yk.y);
            float2 w = __ldg(&g.tw[mk]);
            w.y = -w.y;
            const float2 oo = cmul(w, d);
            other[mk] = make_float2(ee.x - oo.y, ee.y + oo.x);
        }
    }
    __syncthreads();
    float2* r = smem_fft(other, z, m, g.tw, true);
    const float scal = g.scaling ? g.scaling[member] : g.scaling_scalar;
    float* wake = g.wake + (size_t)member * n;
    for (uint32_t x = threadIdx.x; x < n; x += blockDim.x) {
        const float2 v = r[x >> 1];
        wake[x] = __fmul_rn(scal, (x & 1) ? v.y : v.x);  // src/PS/ElectricField.cpp:236-241
    }
}

// twiddle table exp(-2 pi i k / nmax), k < nmax/2, computed in double
ivs_status get_twiddles(ivs_ctx* ctx, uint32_t nmax, const float2** out) {
    for (int i = 0; i < 16; i++) {
        if (ctx->twiddle_n[i] == nmax) {
            *out = (const float2*)ctx->twiddle[i];
            return IVS_OK;
        }
    }
    int slot = -1;
    for (int i = 0; i < 16; i++) if (ctx->twiddle_n[i] == 0) { slot = i; break; }
    if (slot < 0) return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "too many distinct FFT sizes on one context");
    std::vector<float2> tw(nmax / 2);
    for (uint32_t k = 0; k < nmax / 2; k++) {
        const double ph = -2.0 * 3.141592653589793238462643383279502884 * (double)k / (double)nmax;
        tw[k] = make_float2((float)std::cos(ph), (float)std::sin(ph));
    }
    void* d = nullptr;
    IVS_CUDA(ctx, cudaMalloc(&d, sizeof(float2) * (nmax / 2)));
    IVS_CUDA(ctx, cudaMemcpyAsync(d, tw.data(), sizeof(float2) * (nmax / 2), cudaMemcpyHostToDevice, ctx->stream));
    IVS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->twiddle[slot] = d;
    ctx->twiddle_n[slot] = nmax;
    *out = (const float2*)d;
    return IVS_OK;
}

bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

}  // namespace

ivs_status ivs_field_prepare(ivs_ctx* ctx, uint32_t nmax) {
    if (!is_pow2(nmax) || nmax < 4 || nmax > 16384)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "FFT length must be a power of two in [4, 16384]");
    const float2* tw = nullptr;
    ivs_status st = get_twiddles(ctx, nmax, &tw);
    if (st != IVS_OK) return st;
    IVS_CUDA(ctx, cudaFuncSetAttribute(field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * sizeof(float2)));
    return IVS_OK;
}

// batched entry points used by the simulation object
ivs_status ivs_field_launch(ivs_ctx* ctx, const float* profile, uint32_t n, uint32_t members,
                            const float* impedance, uint32_t nmax, const float* d_scaling,
                            float scaling, float* wake, float* formfactor, float* spectrum,
                            float* power, float dq, float f_scale, float cutoff, int mode) {
    if (!is_pow2(nmax) || nmax < 4 || nmax > 16384)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "FFT length must be a power of two in [4, 16384]");
    if (n > nmax) return ivs_fail(ctx, IVS_ERR_INVALID_ARGUMENT, "profile longer than the padded length");
    field_args g;
    g.profile = profile; g.impedance = (const float2*)impedance;
    ivs_status st = get_twiddles(ctx, nmax, &g.tw);
    if (st != IVS_OK) return st;
    g.scaling = d_scaling; g.scaling_scalar = scaling;
    g.wake = wake; g.formfactor = (float2*)formfactor; g.spectrum = spectrum; g.power = power;
    g.dq = dq; g.f_scale = f_scale; g.cutoff = cutoff; g.n = n; g.nmax = nmax; g.mode = mode;
    const size_t smem = sizeof(float2) * nmax;  // two buffers of nmax/2
    if (smem > 48 * 1024)
        IVS_CUDA(ctx, cudaFuncSetAttribute(field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * sizeof(float2)));
    const int threads = (nmax / 4 >= (uint32_t)kFftThreads) ? kFftThreads : (nmax / 4 < 32 ? 32 : (int)(nmax / 4));
    field_kernel<<<members, threads, smem, ctx->stream>>>(g);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}

extern "C" {

ivs_status ivs_wake_potential(ivs_ctx* ctx, const void* d_profile, uint32_t n,
                              const void* d_impedance, uint32_t nmax, float wakescaling,
                              void* d_wake, void* d_formfactor) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_profile && d_impedance && d_wake && n >= 1, "ivs_wake_potential: null pointer");
    return ivs_field_launch(ctx, (const float*)d_profile, n, 1, (const float*)d_impedance, nmax,
                            nullptr, wakescaling, (float*)d_wake, (float*)d_formfactor, nullptr,
                            nullptr, 0, 0, 0, 0);
}

ivs_status ivs_csr_power(ivs_ctx* ctx, const void* d_profile, uint32_t n, const void* d_impedance,
                         uint32_t nmax, float dq, float f_scale, float cutoff, void* d_spectrum,
                         void* d_power) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_profile && d_impedance && d_power && n >= 1, "ivs_csr_power: null pointer");
    return ivs_field_launch(ctx, (const float*)d_profile, n, 1, (const float*)d_impedance, nmax,
                            nullptr, 0, nullptr, nullptr, (float*)d_spectrum, (float*)d_power, dq,
                            f_scale, cutoff, 1);
}

}  // extern "C"
