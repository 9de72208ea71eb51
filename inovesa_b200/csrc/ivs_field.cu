// ivs_field.cu -- ElectricField on the device: hand-written shared-memory FFT convolution of
// the bunch profile with the impedance (wake potential) and the CSR spectrum / power.
//
// Replaces the reference's clFFT / FFTW plans (src/PS/ElectricField.cpp:73-78,300-306;
// src/FFTWWrapper.cpp:61-107) and the kernels `wakeloss` (:308-316) and `scalewp` (:324-332):
// pad -> r2c -> x Z -> c2r -> scale runs as ONE kernel, one CTA per profile, everything in
// shared memory.  A real transform of length nmax is a complex FFT of length m = nmax/2 plus the
// usual even/odd split.  The complex FFT is a Stockham autosort with radix-8 / radix-4 butterflies
// in registers (4 stages at nmax = 2048, 5 at 16384; one block barrier per stage), twiddles from a
// table in shared memory that is computed once per context in double precision.  Two prunings use
// the zero padding (nmax >= 8 n): the first radix-8 stage of the forward transform sees one
// non-zero input per butterfly and is a broadcast folded into the load, and the last stage of the
// inverse transform only forms the n outputs that are kept.
// In CSR mode the same kernel can also form the out-step observables of src/main.cpp:961-1033
// (integral, mean / variance / rms of both projections), so that an out-step is one launch after
// the Y projection.
#include <cmath>
#include <vector>

#include "ivs_common.cuh"
#include "ivs_internal.h"

namespace {

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
// multiplication by -i (forward transforms) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// exp(-2 pi i q / nmax) for q < nmax from the half-circle table tw[q], q < m = nmax/2
template <bool INV>
__device__ __forceinline__ float2 twiddle(const float2* tw, uint32_t q, uint32_t m) {
    float2 w = tw[q & (m - 1)];
    if (q & m) { w.x = -w.x; w.y = -w.y; }
    if (INV) w.y = -w.y;
    return w;
}

template <bool INV>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), b3 = rot90<INV>(csub(a1, a3));
    a0 = cadd(b0, b2);
    a1 = cadd(b1, b3);
    a2 = csub(b0, b2);
    a3 = csub(b1, b3);
}

// One Stockham stage of radix R on a complex sequence of length m held in shared memory:
// sub-transforms of length p become sub-transforms of length p R.  Butterfly i < m / R reads
// a[i + t m/R], t < R, scales by exp(-+ 2 pi i t k / (p R)), k = i mod p, and writes b[j + u p],
// j = (i - k) R + k.  `olimit`: outputs at or beyond this index are not needed (inverse pruning).
template <int R, bool INV>
__device__ __forceinline__ void stage(const float2* __restrict__ a, float2* __restrict__ b, uint32_t m, uint32_t p,
                                      const float2* tw, uint32_t olimit) {
    const uint32_t nb = m / R, qs = (2 * m) / (p * R);
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
        const uint32_t k = i & (p - 1);
        const uint32_t j = (i - k) * R + k;
        if (j >= olimit) continue;
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; t++) v[t] = a[i + t * nb];
        if (p > 1) {
#pragma unroll
            for (int t = 1; t < R; t++) v[t] = cmul(twiddle<INV>(tw, (uint32_t)t * k * qs, m), v[t]);
        }
        if (R == 2) {
            const float2 u0 = v[0];
            v[0] = cadd(u0, v[1]);
            v[1] = csub(u0, v[1]);
        } else if (R == 4) {
            dft4<INV>(v[0], v[1], v[2], v[3]);
        } else {
            // radix 8 = two radix-4 transforms of the even / odd inputs and the w8 twiddles
            dft4<INV>(v[0], v[2], v[4], v[6]);
            dft4<INV>(v[1], v[3], v[5], v[7]);
            const float h = 0.70710678118654752440f;
            // w8^1 = h (1 -+ i), w8^2 = -+ i, w8^3 = h (-1 -+ i)
            const float2 o1 = INV ? make_float2(h * (v[3].x - v[3].y), h * (v[3].x + v[3].y))
                                  : make_float2(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));
            const float2 o2 = rot90<INV>(v[5]);
            const float2 o3 = INV ? make_float2(-h * (v[7].x + v[7].y), h * (v[7].x - v[7].y))
                                  : make_float2(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));
            const float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
            v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
            v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
            v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
            v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
        }
#pragma unroll
        for (int u = 0; u < R; u++)
            if (j + u * p < olimit) b[j + u * p] = v[u];
    }
}

// Complex FFT of length m (power of two >= 8) in shared memory; the stages already done cover
// sub-transforms of length p0 (1: none).  Radix 8 while it fits, then 4 / 2 for the rest.
// Returns the buffer that holds the result (a or b).  All threads of the CTA must call it.
template <bool INV>
__device__ float2* smem_fft(float2* a, float2* b, uint32_t m, uint32_t p0, const float2* tw, uint32_t olimit) {
    uint32_t p = p0;
    while (p < m) {
        const uint32_t left = m / p;
        // only the last stage may prune its outputs
        if (left >= 64 || left == 8) {
            stage<8, INV>(a, b, m, p, tw, left == 8 ? olimit : 0xffffffffu);
            p *= 8;
        } else if (left >= 4) {  // 32 -> 4 . 8, 16 -> 4 . 4, 4 -> 4
            stage<4, INV>(a, b, m, p, tw, left == 4 ? olimit : 0xffffffffu);
            p *= 4;
        } else {
            stage<2, INV>(a, b, m, p, tw, olimit);
            p *= 2;
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
    // out-step observables (mode 1, optional): obs[member*8 + {integral, mean/var/rms q, mean/var/rms p, power}]
    float* obs;
    const float* projy;        // [members][n]
    const float* ws;           // [n]
    float qmin, pmin, dp;
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_dd(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sums (fixed tree), valid in every thread; red: 32 entries
__device__ float block_sum_f(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum_f(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return warp_sum_f(lane < nw ? red[lane] : 0.0f);
}
__device__ double block_sum_dd(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum_dd(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return warp_sum_dd(lane < nw ? red[lane] : 0.0);
}

// PhaseSpace::average + variance (src/PS/PhaseSpace.cpp:241-289) of one projection, by the CTA
__device__ void moments_block(const float* __restrict__ p, uint32_t n, float amin, float delta, float filling,
                              float* red, double* redd, float* out3) {
    float acc = 0.0f;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float q = __fadd_rn(amin, __fmul_rn((float)i, delta));  // Ruler::at(i)
        acc = __fadd_rn(acc, __fmul_rn(p[i], q));
    }
    acc = block_sum_f(acc, red);
    const float scale = __fdiv_rn(delta, filling);
    const float mean = __fmul_rn(acc, scale);
    double vacc = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float q = __fadd_rn(amin, __fmul_rn((float)i, delta));
        const double d = (double)__fsub_rn(q, mean);  // std::pow(float,int) works in double
        vacc += (double)p[i] * (d * d);               // accumulated in double (the reference rounds `var` to float per term)
    }
    vacc = block_sum_dd(vacc, redd);
    if (threadIdx.x == 0) {
        const float var = __fmul_rn((float)vacc, scale);
        out3[0] = mean;
        out3[1] = var;
        out3[2] = sqrtf(var);
    }
}

__global__ void __launch_bounds__(512)
field_kernel(field_args g) {
    extern __shared__ float2 fsm[];
    __shared__ float red[32];
    __shared__ double redd[32];
    const uint32_t n = g.n, nmax = g.nmax, m = nmax >> 1;
    float2* a = fsm;
    float2* b = fsm + m;
    float2* tw = fsm + 2 * m;  // exp(-2 pi i q / nmax), q < m
    const uint32_t member = blockIdx.x;
    const float* prof = g.profile + (size_t)member * n;

    for (uint32_t q = threadIdx.x; q < m; q += blockDim.x) tw[q] = __ldg(&g.tw[q]);
    // padBunchProfiles (src/PS/ElectricField.cpp:255-263): profile in bins 0..n-1, zeros after;
    // packed as z[j] = x[2j] + i x[2j+1].  With nmax >= 8 n only the first m/8 packed entries can be
    // non-zero, and the first radix-8 stage (inputs m/8 apart, no twiddles) copies each of them to
    // its 8 outputs: it is done here.
    const bool bcast = m >= 64 && n <= m / 4;
    if (bcast) {
        for (uint32_t i = threadIdx.x; i < m / 8; i += blockDim.x) {
            const uint32_t i0 = 2 * i, i1 = 2 * i + 1;
            const float2 v = make_float2(i0 < n ? prof[i0] : 0.0f, i1 < n ? prof[i1] : 0.0f);
            float4* dst = reinterpret_cast<float4*>(b + 8 * i);
#pragma unroll
            for (int u = 0; u < 4; u++) dst[u] = make_float4(v.x, v.y, v.x, v.y);
        }
    } else {
        for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) {
            const uint32_t i0 = 2 * j, i1 = 2 * j + 1;
            a[j] = make_float2(i0 < n ? prof[i0] : 0.0f, i1 < n ? prof[i1] : 0.0f);
        }
    }
    __syncthreads();
    float2* z = bcast ? smem_fft<false>(b, a, m, 8, tw, 0xffffffffu) : smem_fft<false>(a, b, m, 1, tw, 0xffffffffu);
    float2* other = (z == a) ? b : a;

    // even/odd split: X[k] = E[k] + W^k O[k], E = (Z[k]+conj Z[m-k])/2, O = (Z[k]-conj Z[m-k])/(2i)
    auto spectrum_at = [&](uint32_t k, float2 zk, float2 zmk) -> float2 {
        const float2 e = make_float2(0.5f * (zk.x + zmk.x), 0.5f * (zk.y - zmk.y));
        const float2 o = make_float2(0.5f * (zk.y + zmk.y), -0.5f * (zk.x - zmk.x));
        const float2 w = (k < m) ? tw[k] : make_float2(-1.0f, 0.0f);
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
        acc = block_sum_f(acc, red);
        if (threadIdx.x == 0 && g.power) g.power[member] = acc;
        if (g.obs) {
            // out-step block of src/main.cpp:961-1033: integrate (src/PS/PhaseSpace.cpp:229-238),
            // variance(0), variance(1) on the two projections
            float* obs = g.obs + (size_t)member * 8;
            float q = 0.0f;
            for (uint32_t x = threadIdx.x; x < n; x += blockDim.x) q = __fadd_rn(q, __fmul_rn(prof[x], g.ws[x]));
            const float filling = block_sum_f(q, red);
            if (threadIdx.x == 0) {
                obs[0] = filling;
                obs[7] = acc;
            }
            moments_block(prof, n, g.qmin, g.dq, filling, red, redd, obs + 1);
            moments_block(g.projy + (size_t)member * n, n, g.pmin, g.dp, filling, red, redd, obs + 4);
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
            float2 w = (k < m) ? tw[k] : make_float2(-1.0f, 0.0f);
            w.y = -w.y;
            const float2 oo = cmul(w, d);
            other[k] = make_float2(ee.x - oo.y, ee.y + oo.x);
        }
        if (mk != k && mk < m) {
            const float2 ee = make_float2(ymk.x + yk.x, ymk.y - yk.y);
            const float2 d = make_float2(ymk.x - yk.x, ymk.y + yk.y);
            float2 w = tw[mk];
            w.y = -w.y;
            const float2 oo = cmul(w, d);
            other[mk] = make_float2(ee.x - oo.y, ee.y + oo.x);
        }
    }
    __syncthreads();
    // only the first n real samples = (n + 1) / 2 packed entries of the result are kept
    float2* r = smem_fft<true>(other, z, m, 1, tw, (n + 1) / 2);
    const float scal = g.scaling ? g.scaling[member] : g.scaling_scalar;
    float* wake = g.wake + (size_t)member * n;
    for (uint32_t x = threadIdx.x; x < n; x += blockDim.x) {
        const float2 v = r[x >> 1];
        wake[x] = __fmul_rn(scal, (x & 1) ? v.y : v.x);  // src/PS/ElectricField.cpp:236-241
    }
}

// small transforms (nmax < 16: unit tests of the reference use n = 4) keep the plain radix-2 path
// of smem_fft: stage<2> handles any power of two

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

constexpr size_t kFieldMaxSmem = 3 * 8192 * sizeof(float2);  // two data buffers and the twiddles at nmax = 16384

}  // namespace

ivs_status ivs_field_prepare(ivs_ctx* ctx, uint32_t nmax) {
    if (!is_pow2(nmax) || nmax < 4 || nmax > 16384)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "FFT length must be a power of two in [4, 16384]");
    const float2* tw = nullptr;
    ivs_status st = get_twiddles(ctx, nmax, &tw);
    if (st != IVS_OK) return st;
    IVS_CUDA(ctx, cudaFuncSetAttribute(field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFieldMaxSmem));
    return IVS_OK;
}

// batched entry points used by the simulation object
ivs_status ivs_field_launch(ivs_ctx* ctx, const float* profile, uint32_t n, uint32_t members,
                            const float* impedance, uint32_t nmax, const float* d_scaling,
                            float scaling, float* wake, float* formfactor, float* spectrum,
                            float* power, float dq, float f_scale, float cutoff, int mode,
                            const ivs_field_obs* o) {
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
    g.obs = nullptr; g.projy = nullptr; g.ws = nullptr; g.qmin = g.pmin = g.dp = 0.0f;
    if (o) { g.obs = o->obs; g.projy = o->projy; g.ws = o->ws; g.qmin = o->qmin; g.pmin = o->pmin; g.dp = o->dp; }
    const size_t smem = sizeof(float2) * (nmax / 2) * 3;  // two buffers of nmax/2 and the twiddle table
    if (smem > 40 * 1024 && !(ctx->optin & 2u)) {  // the kernel's static shared memory counts against the 48 KB default as well
        IVS_CUDA(ctx, cudaFuncSetAttribute(field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFieldMaxSmem));
        ctx->optin |= 2u;  // once per context (the call costs more than the kernel)
    }
    // one radix-8 butterfly per thread and stage up to nmax = 8192, two at 16384
    uint32_t threads = nmax / 16;
    threads = threads < 32 ? 32 : (threads > 512 ? 512 : threads);
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
