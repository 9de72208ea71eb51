// ivs_track.cu -- passive particle tracking: SourceMap::applyToAll (src/SM/SourceMap.cpp:111-117)
// with the applyTo() of each map, one thread per particle.  Positions are PhaseSpace::Position
// {x, y} in grid units.  std::min / std::max are spelled out because their NaN behaviour
// (return the first argument) is part of the reference's results.
#include "ivs_common.cuh"

namespace {

__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ uint32_t f2u(float v) {
    return (v == v && fabsf(v) < 9.0e18f) ? (uint32_t)(long long)v : 0u;
}

// index of the table entry below a coordinate: `meshindex_t yi = yif` with yif = trunc(pos) in
// KickMap::applyTo (src/SM/KickMap.cpp:284-287), a float -> uint32_t conversion that C++ defines only
// for [0, 2^32).  There it is exact here too; outside (NaN, negative, huge: a particle the caller put
// off the grid) the conversion saturates (NaN -> 0), and the callers test `k < n - 1` rather than
// `k + 1 < n` so that a saturated index can never wrap into the table.
__device__ __forceinline__ uint32_t f2idx(float v) { return __float2uint_rz(v); }

// KickMap::applyTo (src/SM/KickMap.cpp:280-303): the kick is interpolated linearly in the
// perpendicular coordinate and SUBTRACTED (forward map), then clamped to [1, n-1].
__global__ void track_kick_kernel(float2* __restrict__ particles, uint32_t np,
                                  const float* __restrict__ offset, uint32_t n, int axis) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    float2 p = particles[i];
    const float perp = (axis == 0) ? p.y : p.x;
    float par = (axis == 0) ? p.x : p.y;
    const float ip = truncf(perp);
    const float fr = IVS_SUB(perp, ip);
    const uint32_t k = f2idx(perp);
    if (k < n - 1) {
        const float s = IVS_ADD(IVS_MUL(IVS_SUB(1.0f, fr), offset[k]), IVS_MUL(fr, offset[k + 1]));
        par = IVS_SUB(par, s);
    }
    par = std_max(1.0f, std_min(par, (float)(n - 1)));
    if (axis == 0) p.x = par; else p.y = par;
    particles[i] = p;
}

// counter-based normal deviates for FPTracking::stochastic: splitmix64 of (seed, step, particle)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Box-Muller on the special-function unit: one 64-bit hash gives two independent N(0,1) deviates.
// u1 = (k + 1) / (2^24 + 1) in (0, 1), the angle 2 pi u2 with u2 in [-1/2, 1/2) (where sin.approx /
// cos.approx are at their best: absolute error ~2^-21).  lg2.approx has an ABSOLUTE error of ~2^-22, which
// next to u1 = 1 exceeds |log u1| and could turn the radicand negative: it is clamped at 0 (the radius
// is then off by at most 1e-3 sigma, with probability ~1e-5 per draw).  The reference draws from an
// unseeded std::mt19937 + std::normal_distribution (src/SM/FokkerPlanckMap.cpp:262-266), so there is
// no stream to reproduce, only the distribution (tests: first two moments, reproducibility per seed).
__device__ __forceinline__ void normal_pair(uint64_t h, float& g0, float& g1) {
    const float u1 = ((float)(uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777217.0f);
    const float a = ((float)(uint32_t)((h >> 8) & 0xFFFFFFu) - 8388608.0f) * (6.283185307179586f / 16777216.0f);
    float l, r, sn, cs;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
    const float q = fmaxf(l * -1.3862943611198906f, 0.0f);  // -2 ln u1
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(a));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(a));
    g0 = r * cs;
    g1 = r * sn;
}
__device__ __forceinline__ float normal01(uint64_t seed, uint64_t step, uint32_t i) {
    float g0, g1;
    normal_pair(mix64(mix64(seed ^ mix64(step)) + i), g0, g1);
    return g0;
}

// FokkerPlanckMap::applyTo (src/SM/FokkerPlanckMap.cpp:214-268); w5 = stencil weights on
// y-2..y+2 in the reference's tap order (ascending index)
__global__ void track_fp_kernel(float2* __restrict__ particles, uint32_t np,
                                const float* __restrict__ grid, uint32_t n,
                                const float* __restrict__ w5, uint32_t fptrack, float e1, float dp,
                                uint64_t seed, uint64_t step) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    float2 p = particles[i];
    if (fptrack == 1) {
        uint32_t yi = f2u(floorf(p.y));
        if (yi > n - 1) yi = n - 1;  // (the reference would read past its table for yi == n)
        float offset = 0.0f;
        for (int k = 0; k < 5; k++) {
            const float w = w5[yi * 5 + k];
            if (w != 0.0f) offset = IVS_ADD(IVS_MUL((float)(2 - k), w), offset);
        }
        p.y = std_max(1.0f, std_min(IVS_ADD(p.y, offset), (float)(n - 1)));
    } else if (fptrack == 2) {
        uint32_t xu = f2u(floorf(p.x));
        if (xu > n - 1) xu = n - 1;
        int yi = (int)floorf(p.y);
        if (yi > (int)(n - 1)) yi = (int)(n - 1);
        if (yi < 0) yi = 0;
        const float* row = grid + (size_t)xu * n;
        float offset = 0.0f, charge = 0.0f;
        for (int k = 0; k < 5; k++) {
            const float w = w5[yi * 5 + k];
            const int idx = yi + k - 2;
            if (w != 0.0f && (uint32_t)idx < n) {
                const float cw = IVS_MUL(row[idx], w);
                charge = IVS_ADD(cw, charge);
                offset = IVS_ADD(IVS_MUL(cw, (float)(idx - yi)), offset);
            }
        }
        offset = IVS_DIV(offset, charge);
        p.y = std_max(1.0f, std_min(IVS_ADD(p.y, offset), (float)(n - 1)));
    } else if (fptrack == 3) {
        // pos.y -= pos.y*e1 + N(0, sqrt(2 e1)/dp)   (:264-266, std::normal_distribution)
        const float sigma = IVS_DIV(sqrtf(IVS_MUL(2.0f, e1)), dp);
        const float g = IVS_MUL(sigma, normal01(seed, step, i));
        p.y = IVS_SUB(p.y, IVS_ADD(IVS_MUL(p.y, e1), g));
    }
    particles[i] = p;
}

// RotationMap::applyTo (src/SM/RotationMap.cpp:178-188)
__global__ void track_rotate_kernel(float2* __restrict__ particles, uint32_t np, float cos_dt,
                                    float sin_dt, float zb0, float zb1) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const float2 p = particles[i];
    const float dx = IVS_SUB(p.x, zb0), dy = IVS_SUB(p.y, zb1);
    float2 r;
    r.x = IVS_ADD(IVS_ADD(IVS_MUL(cos_dt, dx), IVS_MUL(sin_dt, dy)), zb0);
    r.y = IVS_ADD(IVS_SUB(IVS_MUL(cos_dt, dy), IVS_MUL(sin_dt, dx)), zb1);
    particles[i] = r;
}

// the four applyTo() of one simulation step (src/main.cpp:1034-1041) for one particle:
// wake kick (y), RF kick (y), drift (x), Fokker-Planck.  grid_fp_in = input grid of the
// Fokker-Planck map (grid_t3), only read by FPTracking::approximation2.
// Fokker-Planck applyTo of one particle for the deterministic tracking modes 1 and 2
__device__ __forceinline__ void fp_track_one(float2& p, uint32_t n, const float* __restrict__ w5,
                                             const float* __restrict__ grid_member, uint32_t fptrack) {
    if (fptrack == 1) {
        uint32_t yi = f2u(floorf(p.y));
        if (yi > n - 1) yi = n - 1;
        float offset = 0.0f;
        for (int k = 0; k < 5; k++) {
            const float w = w5[yi * 5 + k];
            if (w != 0.0f) offset = IVS_ADD(IVS_MUL((float)(2 - k), w), offset);
        }
        p.y = std_max(1.0f, std_min(IVS_ADD(p.y, offset), (float)(n - 1)));
    } else {
        uint32_t xu = f2u(floorf(p.x));
        if (xu > n - 1) xu = n - 1;
        int yi = (int)floorf(p.y);
        if (yi > (int)(n - 1)) yi = (int)(n - 1);
        if (yi < 0) yi = 0;
        const float* row = grid_member + (size_t)xu * n;
        float offset = 0.0f, charge = 0.0f;
        for (int k = 0; k < 5; k++) {
            const float w = w5[yi * 5 + k];
            const int idx = yi + k - 2;
            if (w != 0.0f && (uint32_t)idx < n) {
                const float cw = IVS_MUL(row[idx], w);
                charge = IVS_ADD(cw, charge);
                offset = IVS_ADD(IVS_MUL(cw, (float)(idx - yi)), offset);
            }
        }
        offset = IVS_DIV(offset, charge);
        p.y = std_max(1.0f, std_min(IVS_ADD(p.y, offset), (float)(n - 1)));
    }
}

// KickMap::applyTo split into its two halves so that the wake kick and the RF kick of a particle (same
// perpendicular coordinate x, unchanged between them) share one index / fraction computation:
// track_cell = modf + index + range test, track_kick = the interpolated kick and the clamp.
// The clamp std::max(1, std::min(par, n - 1)) is written fminf(fmaxf(par, 1), n - 1): the same value for
// every input including NaN (both give 1; n - 1 >= 1), two FMNMX instead of two compare / select pairs.
struct track_cell_t { uint32_t k; float fr; bool ok; };
__device__ __forceinline__ track_cell_t track_cell(float perp, uint32_t n) {
    track_cell_t c;
    c.fr = IVS_SUB(perp, truncf(perp));
    c.k = f2idx(perp);
    c.ok = c.k < n - 1;
    return c;
}
__device__ __forceinline__ void track_kick(float& par, const track_cell_t& c, const float* __restrict__ offset,
                                           float nm1) {
    if (c.ok) {
        const float* q = offset + c.k;  // one address for both taps
        const float s = IVS_ADD(IVS_MUL(IVS_SUB(1.0f, c.fr), q[0]), IVS_MUL(c.fr, q[1]));
        par = IVS_SUB(par, s);
    }
    par = fminf(fmaxf(par, 1.0f), nm1);
}

// One thread advances FOUR particles (i + k * quarter, k = 0..3) through the four maps: their
// dependent table look-ups interleave (the kernel is latency bound: three gathers in a row per
// particle), the launch is less than one wave of CTAs, and in the stochastic mode one Box-Muller
// transform serves two particles.
// The stochastic mode keys its generator on a step counter that lives on the DEVICE
// (counter[0..1] = step as 64 bit, counter[2] = blocks finished), so that a replayed CUDA graph
// still draws fresh numbers every step: every block reads the step when it starts, the last block
// to finish advances it.
#ifndef IVS_TRACK_P
#define IVS_TRACK_P 4
#endif
#ifndef IVS_TRACK_T
#define IVS_TRACK_T 256
#endif
constexpr int kTrackPerThread = IVS_TRACK_P;
constexpr int kTrackThreads = IVS_TRACK_T;
__global__ void __launch_bounds__(kTrackThreads) track_step_kernel(float2* __restrict__ particles, uint32_t np, uint32_t quarter,
                                  uint32_t n, const float* __restrict__ wake, const float* __restrict__ rf,
                                  const float* __restrict__ drift, const float* __restrict__ w5,
                                  const float* __restrict__ grid, uint32_t fptrack, float e1,
                                  float sigma, uint64_t seed, unsigned int* counter,
                                  const int* sim_status, int wb) {
    constexpr int P = kTrackPerThread;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t member = blockIdx.y;
    const bool stochastic = w5 && fptrack == 3;
    // The particles are requested first; the run's status words and the step counter are read by ONE
    // thread per block while they travel (every thread reading those three words -- the same lines for
    // 650 000 threads -- costs more than the block barrier that hands them round)
    float2* base = particles + (size_t)member * np;
    float2 p[P];
    bool on[P];
    const bool mine = i < quarter;
#pragma unroll
    for (int k = 0; k < P; k++) {
        const uint32_t j = i + k * quarter;
        on[k] = mine && j < np;
        if (mine) p[k] = base[on[k] ? j : i];
    }
    __shared__ unsigned long long skey_s;
    __shared__ int stand_down;
    if (threadIdx.x == 0) {
        // the fused step this launch belongs to (or an earlier one) left its staged pad and will be
        // redone by the host: the particles stay where they are (ivs_fused.cu, step_ctl_begin)
        // (wb >= 0: this launch may overlap the NEXT fused step; it stands down only if the step that
        // filled wake buffer wb, or an earlier one, is the one that overflowed)
        int down = 0;
        if (sim_status) {
            const volatile int* st = sim_status;
            const int first_bad = st[0];
            down = wb < 0 ? first_bad != 0x7fffffff : first_bad <= st[2 + wb];
        }
        stand_down = down;
        if (stochastic) {
            const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(counter);
            skey_s = mix64(seed ^ mix64(step));
        }
    }
    __syncthreads();
    if (stand_down) return;
    const unsigned long long skey = stochastic ? skey_s : 0ull;
    if (mine) {
        const float nm1 = (float)(n - 1);
        {
            track_cell_t cx[P];
#pragma unroll
            for (int k = 0; k < P; k++) cx[k] = track_cell(p[k].x, n);
            if (wake) {
                const float* wm = wake + (size_t)member * n;
#pragma unroll
                for (int k = 0; k < P; k++) track_kick(p[k].y, cx[k], wm, nm1);
            }
#pragma unroll
            for (int k = 0; k < P; k++) track_kick(p[k].y, cx[k], rf, nm1);
        }
#pragma unroll
        for (int k = 0; k < P; k++) track_kick(p[k].x, track_cell(p[k].y, n), drift, nm1);
        if (w5 && (fptrack == 1 || fptrack == 2)) {
            const float* gm = grid ? grid + (size_t)member * n * n : nullptr;
#pragma unroll
            for (int k = 0; k < P; k++) fp_track_one(p[k], n, w5, gm, fptrack);
        } else if (stochastic) {
            // pos.y -= pos.y*e1 + N(0, sqrt(2 e1)/dp)   (src/SM/FokkerPlanckMap.cpp:264-266); the
            // reference draws from an unseeded mt19937, here a counter-based generator keyed on
            // (seed, step, member, thread, pair): reproducible run to run
#pragma unroll
            for (int k = 0; k < P; k += 2) {
                const uint64_t h = mix64(skey + ((uint64_t)member << 32) + ((uint64_t)k << 28) + i);
                float g0, g1;
                normal_pair(h, g0, g1);
                p[k].y = IVS_SUB(p[k].y, IVS_ADD(IVS_MUL(p[k].y, e1), IVS_MUL(sigma, g0)));
                p[k + 1].y = IVS_SUB(p[k + 1].y, IVS_ADD(IVS_MUL(p[k + 1].y, e1), IVS_MUL(sigma, g1)));
            }
        }
#pragma unroll
        for (int k = 0; k < P; k++)
            if (on[k]) base[i + k * quarter] = p[k];
    }
    if (stochastic) {
        __syncthreads();  // every thread of the block has read skey_s
        if (threadIdx.x == 0) {
            const unsigned int total = gridDim.x * gridDim.y;
            if (atomicAdd(counter + 2, 1u) == total - 1) {  // the last block: everybody has read the step
                counter[2] = 0;
                unsigned long long* st = reinterpret_cast<unsigned long long*>(counter);
                *st = *st + 1ull;
            }
        }
    }
}


// The same step with the three kick tables STAGED IN SHARED MEMORY (the benchmark's launch: 26 x 10^5
// particles at N = 2048).  track_step_kernel above is bound by the L1's gather wavefronts -- six 4-byte
// gathers per particle into 8 KB tables, about six wavefronts each for the spread of a bunch -- not by
// instructions.  Here a block copies its member's wake potential and the RF and drift tables (3 N floats)
// into shared memory once and walks a CHUNK of the member's particles in batches of 4 per thread, the
// next batch travelling while the current one is computed; the gathers become LDS (about 3.5 wavefronts
// for 32 random banks, no tag stage).  A thread loads and stores its particles as adjacent pairs (16 bytes),
// which are also the pairs of the Box-Muller transform: pairing and generator key (member, pair) do not
// depend on how many chunks the launch uses.
#ifndef IVS_TRACK_ST
#define IVS_TRACK_ST 512
#endif
constexpr int kStagedThreads = IVS_TRACK_ST;
#ifndef IVS_TRACK_STAGED_MIN
#define IVS_TRACK_STAGED_MIN 1
#endif
constexpr uint32_t kStagedMinParticles = IVS_TRACK_STAGED_MIN;
#ifndef IVS_TRACK_SMIN
#define IVS_TRACK_SMIN 3
#endif
__global__ void __launch_bounds__(kStagedThreads, IVS_TRACK_SMIN) track_step_staged_kernel(
    float2* __restrict__ particles, uint32_t np, uint32_t chunk, uint32_t n, const float* __restrict__ wake,
    const float* __restrict__ rf, const float* __restrict__ drift, const float* __restrict__ w5,
    uint32_t fptrack, float e1, float sigma, uint64_t seed, unsigned int* counter, const int* sim_status,
    int wb) {
    constexpr int P = 4;  // particles per thread and batch, as two adjacent pairs (16-byte loads)
    constexpr uint32_t T = kStagedThreads;
    extern __shared__ __align__(16) float tab[];  // [wake n][rf n][drift n]
    float* const t_wake = tab;
    float* const t_rf = tab + n;
    float* const t_drift = tab + 2 * n;
    const uint32_t tid = threadIdx.x;
    const uint32_t member = blockIdx.y;
    const bool stochastic = w5 && fptrack == 3;
    float4* base = reinterpret_cast<float4*>(particles + (size_t)member * np);  // np is even
    const uint32_t h0 = blockIdx.x * (chunk / 2);          // in pairs of particles
    const uint32_t h1 = min(np / 2, h0 + chunk / 2);
    // first batch on its way before anything else
    float4 pp[P / 2];
#pragma unroll
    for (int k = 0; k < P / 2; k++) {
        const uint32_t h = h0 + k * T + tid;
        if (h < h1) pp[k] = base[h];
    }
    __shared__ unsigned long long skey_s;
    __shared__ int stand_down;
    if (tid == 0) {  // see track_step_kernel
        int down = 0;
        if (sim_status) {
            const volatile int* st = sim_status;
            const int first_bad = st[0];
            down = wb < 0 ? first_bad != 0x7fffffff : first_bad <= st[2 + wb];
        }
        stand_down = down;
        if (stochastic) {
            const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(counter);
            skey_s = mix64(seed ^ mix64(step)) + ((uint64_t)member << 32);
        }
    }
    {
        const float4* g_rf = reinterpret_cast<const float4*>(rf);
        const float4* g_dr = reinterpret_cast<const float4*>(drift);
        float4* s_rf = reinterpret_cast<float4*>(t_rf);
        float4* s_dr = reinterpret_cast<float4*>(t_drift);
        if (wake) {
            const float4* g_w = reinterpret_cast<const float4*>(wake + (size_t)member * n);
            float4* s_w = reinterpret_cast<float4*>(t_wake);
            for (uint32_t q = tid; q < n / 4; q += T) s_w[q] = g_w[q];
        }
        for (uint32_t q = tid; q < n / 4; q += T) { s_rf[q] = g_rf[q]; s_dr[q] = g_dr[q]; }
    }
    __syncthreads();
    if (stand_down) return;
    const unsigned long long skey = stochastic ? skey_s : 0ull;
    const float nm1 = (float)(n - 1);
    for (uint32_t hb = h0; hb < h1; hb += (P / 2) * T) {
        float4 nx[P / 2];
#pragma unroll
        for (int k = 0; k < P / 2; k++) {  // the next batch
            const uint32_t h = hb + (P / 2 + k) * T + tid;
            if (h < h1) nx[k] = base[h];
        }
        float2 p[P];
#pragma unroll
        for (int k = 0; k < P / 2; k++) {
            p[2 * k] = make_float2(pp[k].x, pp[k].y);
            p[2 * k + 1] = make_float2(pp[k].z, pp[k].w);
        }
        {
            track_cell_t cx[P];
#pragma unroll
            for (int k = 0; k < P; k++) cx[k] = track_cell(p[k].x, n);
            if (wake) {
#pragma unroll
                for (int k = 0; k < P; k++) track_kick(p[k].y, cx[k], t_wake, nm1);
            }
#pragma unroll
            for (int k = 0; k < P; k++) track_kick(p[k].y, cx[k], t_rf, nm1);
        }
#pragma unroll
        for (int k = 0; k < P; k++) track_kick(p[k].x, track_cell(p[k].y, n), t_drift, nm1);
        if (w5 && fptrack == 1) {
#pragma unroll
            for (int k = 0; k < P; k++) fp_track_one(p[k], n, w5, nullptr, 1);
        } else if (stochastic) {
            // pos.y -= pos.y*e1 + N(0, sqrt(2 e1)/dp)   (src/SM/FokkerPlanckMap.cpp:264-266); one
            // Box-Muller transform per pair of adjacent particles, keyed on (seed, step, member, pair)
#pragma unroll
            for (int k = 0; k < P; k += 2) {
                float g0, g1;
                normal_pair(mix64(skey + (hb + (k / 2) * T + tid)), g0, g1);
                p[k].y = IVS_SUB(p[k].y, IVS_ADD(IVS_MUL(p[k].y, e1), IVS_MUL(sigma, g0)));
                p[k + 1].y = IVS_SUB(p[k + 1].y, IVS_ADD(IVS_MUL(p[k + 1].y, e1), IVS_MUL(sigma, g1)));
            }
        }
#pragma unroll
        for (int k = 0; k < P / 2; k++) {
            const uint32_t h = hb + k * T + tid;
            if (h < h1) base[h] = make_float4(p[2 * k].x, p[2 * k].y, p[2 * k + 1].x, p[2 * k + 1].y);
            pp[k] = nx[k];
        }
    }
    if (stochastic) {
        // every thread of the block read skey_s before the loop; the barrier above orders that read
        if (tid == 0) {
            const unsigned int total = gridDim.x * gridDim.y;
            if (atomicAdd(counter + 2, 1u) == total - 1) {  // the last block: everybody has read the step
                counter[2] = 0;
                unsigned long long* st = reinterpret_cast<unsigned long long*>(counter);
                *st = *st + 1ull;
            }
        }
    }
}

}  // namespace

ivs_status ivs_track_step_launch(ivs_ctx* ctx, float* particles, uint32_t np, uint32_t members,
                                 uint32_t n, const float* wake, const float* rf, const float* drift,
                                 const float* fpw, const float* grid_fp_in, uint32_t fptrack, float e1,
                                 float dp, uint64_t seed, unsigned int* step_counter, const int* sim_status,
                                 int wb, cudaStream_t stream) {
    if (np == 0) return IVS_OK;
    if (!stream) stream = ctx->stream;
    const float sigma = sqrtf(2.0f * e1) / dp;  // uniform, in the reference's fp32 operations
    // tables staged in shared memory where they fit the default 48 KB and FPTracking does not read the grid
    const size_t tab_bytes = (size_t)3 * n * sizeof(float);
    if (n % 4 == 0 && np % 2 == 0 && tab_bytes + 64 <= 48 * 1024 && fptrack != 2 && np >= kStagedMinParticles) {
        const uint32_t batch = 4 * kStagedThreads;
        if (ctx->track_slots_n != n) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, track_step_staged_kernel, kStagedThreads,
                                                              tab_bytes) != cudaSuccess || per_sm < 1)
                per_sm = 1;
            ctx->track_slots = (uint32_t)per_sm * ctx->sm_count;
            ctx->track_slots_n = n;
        }
        // one wave of blocks: as many chunks per member as the machine holds at once, in whole batches
        const uint32_t batches = ivs_div_up(np, batch);
        const uint32_t most = std::max<uint32_t>(1, std::min<uint32_t>(batches, ctx->track_slots / members));
        const uint32_t chunk = ivs_div_up(batches, most) * batch;
        const uint32_t chunks = ivs_div_up(np, chunk);
        track_step_staged_kernel<<<dim3(chunks, members), kStagedThreads, tab_bytes, stream>>>(
            (float2*)particles, np, chunk, n, wake, rf, drift, fpw, fptrack, e1, sigma, seed, step_counter,
            sim_status, wb);
        IVS_LAUNCHED(ctx);
        return IVS_OK;
    }
    const uint32_t quarter = ivs_div_up(np, kTrackPerThread);
    track_step_kernel<<<dim3(ivs_div_up(quarter, kTrackThreads), members), kTrackThreads, 0, stream>>>(
        (float2*)particles, np, quarter, n, wake, rf, drift, fpw, grid_fp_in, fptrack, e1, sigma, seed,
        step_counter, sim_status, wb);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}

ivs_status ivs_track_kick_launch(ivs_ctx* ctx, float* particles, uint32_t np, const float* offset,
                                 uint32_t n, int axis) {
    if (np == 0) return IVS_OK;
    track_kick_kernel<<<ivs_div_up(np, 256), 256, 0, ctx->stream>>>((float2*)particles, np, offset, n, axis);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
ivs_status ivs_track_fp_launch(ivs_ctx* ctx, float* particles, uint32_t np, const float* grid,
                               uint32_t n, const float* w5, uint32_t fptrack, float e1, float dp,
                               uint64_t seed, uint64_t step) {
    if (np == 0 || fptrack == 0) return IVS_OK;
    track_fp_kernel<<<ivs_div_up(np, 256), 256, 0, ctx->stream>>>((float2*)particles, np, grid, n, w5,
                                                                 fptrack, e1, dp, seed, step);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}

extern "C" {

ivs_status ivs_track_kick(ivs_ctx* ctx, void* d_particles, uint32_t np, const void* d_offset,
                          uint32_t n, int axis) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, (d_particles || np == 0) && d_offset && n >= 2, "ivs_track_kick: null pointer");
    IVS_REQUIRE(ctx, axis == 0 || axis == 1, "ivs_track_kick: axis must be 0 (x) or 1 (y)");
    return ivs_track_kick_launch(ctx, (float*)d_particles, np, (const float*)d_offset, n, axis);
}

ivs_status ivs_track_fokker_planck(ivs_ctx* ctx, void* d_particles, uint32_t np,
                                   const void* d_grid, uint32_t n, const void* d_weights,
                                   uint32_t fptrack, float e1, float dp, uint64_t seed,
                                   uint64_t step) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, fptrack <= 3, "ivs_track_fokker_planck: FPTracking must be 0..3");
    IVS_REQUIRE(ctx, (d_particles || np == 0) && n >= 4, "ivs_track_fokker_planck: null pointer");
    IVS_REQUIRE(ctx, fptrack == 0 || fptrack == 3 || d_weights, "ivs_track_fokker_planck: weights needed");
    IVS_REQUIRE(ctx, fptrack != 2 || d_grid, "ivs_track_fokker_planck: grid needed for approximation2");
    return ivs_track_fp_launch(ctx, (float*)d_particles, np, (const float*)d_grid, n,
                               (const float*)d_weights, fptrack, e1, dp, seed, step);
}

ivs_status ivs_track_rotate(ivs_ctx* ctx, void* d_particles, uint32_t np, float angle,
                            float zerobin_x, float zerobin_y) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_particles || np == 0, "ivs_track_rotate: null pointer");
    if (np == 0) return IVS_OK;
    track_rotate_kernel<<<ivs_div_up(np, 256), 256, 0, ctx->stream>>>(
        (float2*)d_particles, np, cosf(-angle), sinf(-angle), zerobin_x, zerobin_y);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}

}  // extern "C"
