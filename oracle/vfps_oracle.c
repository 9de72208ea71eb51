/* vfps_oracle.c -- CPU restatement of the reference's per-step Vlasov-Fokker-Planck update.
 *
 * TEST INFRASTRUCTURE ONLY: see vfps_oracle.h for who may use it and how its parity is pinned.
 * Every function cites the reference file:line (relative to /root/reference/) it follows.
 * All grid arithmetic is float, in the reference's operation order.  MADD(a,b,c) marks the
 * multiply-adds GCC contracts in the reference at -O3 on FMA hardware: with -DORACLE_FMA they
 * are fused (fmaf), otherwise a*b is rounded before the add.  Build with -ffp-contract=off.
 */
#include "vfps_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_FMA
#define MADD(a, b, c) fmaf((a), (b), (c))
#define MADD_D(a, b, c) fma((a), (b), (c))
int orc_is_fma(void) { return 1; }
#else
#define MADD(a, b, c) ((a) * (b) + (c))
#define MADD_D(a, b, c) ((a) * (b) + (c))
int orc_is_fma(void) { return 0; }
#endif

static const double kC = 2.99792458e8;            /* inc/defines.hpp:48 */
static const double kEps0 = 8.854187817e-12;      /* :51 */
static const double kE = 1.602e-19;               /* :57 */
static const double kMe = 510998.9;               /* :63 */
static const double kPi = 3.141592653589793238462643383279502884;
static const double kTwoPi = 6.283185307179586476925286766559005768;
static const double kPiSqr = 9.869604401089358618834490999876151135;
static const double kZ0wall = 376.730313461;      /* inc/Z/Impedance.hpp:94 */

/* float -> uint32 conversion of a possibly negative value the way x86-64 GCC does it for
 * the reference's `meshindex_t jd = qp_int;` (cvttss2si to 64 bit, low half kept) */
static uint32_t f2u(float v) { return (uint32_t)(int64_t)v; }

/* std::min / std::max exactly as libstdc++ defines them (matters when an argument is NaN) */
static float std_min(float a, float b) { return (b < a) ? b : a; }
static float std_max(float a, float b) { return (a < b) ? b : a; }

/* PLAIN(a,b,c): a multiply-add that stays unfused even in the FMA flavour, because GCC 13 at
 * -O3 -march=x86-64-v3 vectorises the enclosing reduction loop as rounded products followed
 * by in-order adds (verified bit-for-bit against oracle/_ref/libinovesa_ref_fma.so) */
#define PLAIN(a, b, c) ((a) * (b) + (c))

/* ------------------------------------------------------------------ Ruler */

/* inc/PS/Ruler.hpp:38 */
float orc_ruler_delta(uint32_t steps, float min, float max) {
    return (max - min) / (float)(steps - 1);
}
/* inc/PS/Ruler.hpp:40 */
float orc_ruler_zerobin(uint32_t steps, float min, float max) {
    return ((min + max) / (min - max) + 1) * (float)(steps - 1) / 2;
}
/* inc/PS/Ruler.hpp:47-49 */
void orc_ruler_axis(uint32_t steps, float min, float max, float* axis) {
    const float delta = orc_ruler_delta(steps, min, max);
    for (uint32_t i = 0; i < steps; i++) axis[i] = MADD((float)i, delta, min);
}

/* ------------------------------------------------------------------ interpolation */

/* src/SM/SourceMap.cpp:142-170 */
void orc_calc_coefficients(float f, uint32_t it, float* ic) {
    switch (it) {
    case 1:
        ic[0] = 1;
        break;
    case 2:
        ic[0] = 1.0f - f;
        ic[1] = f;
        break;
    case 3:
        ic[0] = f * (f - 1.0f) / 2.0f;
        ic[1] = MADD(-f, f, 1.0f); /* 1 - f*f */
        ic[2] = f * (f + 1.0f) / 2.0f;
        break;
    case 4:
        ic[0] = (f - 1.0f) * (f - 2.0f) * f * (float)(-1. / 6.);
        ic[1] = (f + 1.0f) * (f - 1.0f) * (f - 2.0f) / 2.0f;
        ic[2] = (2.0f - f) * f * (f + 1.0f) / 2.0f;
        ic[3] = f * (f + 1.0f) * (f - 1.0f) * (float)(1. / 6.);
        break;
    default:
        break;
    }
}

/* src/SM/KickMap.cpp:339-389 */
void orc_kick_update_sm(const float* offset, uint32_t n_entries, uint32_t n_kd, uint32_t it,
                        uint32_t* index, float* weight) {
    float smc[4];
    for (uint32_t i = 0; i < n_entries; i++) {
        const float poffs = (float)(n_kd / 2) + offset[i];
        float qp_int;
        const float xip = modff(poffs, &qp_int);
        const uint32_t jd = f2u(qp_int);
        if (jd < n_kd) {
            orc_calc_coefficients(xip, it, smc);
            for (uint32_t j1 = 0; j1 < it; j1++) {
                const uint32_t j0 = jd + j1 - (it - 1) / 2;
                if (j0 < n_kd) {
                    index[i * it + j1] = j0;
                    weight[i * it + j1] = smc[j1];
                } else {
                    index[i * it + j1] = n_kd / 2;
                    weight[i * it + j1] = 0;
                }
            }
        } else {
            for (uint32_t j1 = 0; j1 < it; j1++) {
                index[i * it + j1] = n_kd / 2;
                weight[i * it + j1] = 0;
            }
        }
    }
}

/* src/SM/KickMap.cpp:221-266 with a prepared tap table */
static void kick_apply_table(uint32_t n, uint32_t ip, int axis, const uint32_t* index,
                             const float* weight, const float* in, float* out) {
    if (axis == 0) {
        for (uint32_t x = 0; x < n; x++) {
            for (uint32_t y = 0; y < n; y++) {
                float value = 0;
                for (uint32_t j = 0; j < ip; j++) {
                    const uint32_t xs =
                        (uint32_t)((int32_t)(x + index[y * ip + j]) - (int32_t)(n / 2));
                    if (xs < n) value = MADD(in[xs * n + y], weight[y * ip + j], value);
                }
                out[x * n + y] = value;
            }
        }
    } else {
        for (uint32_t x = 0; x < n; x++) {
            const uint32_t offs = x * n;
            for (uint32_t y = 0; y < n; y++) {
                float value = 0;
                for (uint32_t j = 0; j < ip; j++) {
                    const uint32_t ys =
                        (uint32_t)((int32_t)(y + index[x * ip + j]) - (int32_t)(n / 2));
                    if (ys < n) value = MADD(in[offs + ys], weight[x * ip + j], value);
                }
                out[offs + y] = value;
            }
        }
    }
}

/* src/SM/KickMap.cpp:194-271 (swapOffset -> updateSM -> apply) */
void orc_kick_apply(uint32_t n, uint32_t it, int axis, const float* offset, const float* in,
                    float* out) {
    uint32_t* index = (uint32_t*)malloc(sizeof(uint32_t) * n * it);
    float* weight = (float*)malloc(sizeof(float) * n * it);
    orc_kick_update_sm(offset, n, n, it, index, weight);
    kick_apply_table(n, it, axis, index, weight, in, out);
    free(index);
    free(weight);
}

/* src/SM/KickMap.cpp:280-303 */
void orc_kick_track(uint32_t n, int axis, const float* offset, float* particles, uint32_t np) {
    for (uint32_t i = 0; i < np; i++) {
        float* px = &particles[2 * i];
        float* py = &particles[2 * i + 1];
        if (axis == 0) {
            float yif;
            const float yf = modff(*py, &yif);
            const uint32_t yi = f2u(yif);
            if (yi + 1 < n) *px -= MADD(1 - yf, offset[yi], yf * offset[yi + 1]);
            *px = std_max(1.0f, std_min(*px, (float)(n - 1)));
        } else {
            float xif;
            const float xf = modff(*px, &xif);
            const uint32_t xi = f2u(xif);
            if (xi + 1 < n) *py -= MADD(1 - xf, offset[xi], xf * offset[xi + 1]);
            *py = std_max(1.0f, std_min(*py, (float)(n - 1)));
        }
    }
}

/* ------------------------------------------------------------------ RF kick, drift */

/* src/SM/RFKickMap.cpp:15-34 (linear constructor) + :59-68 */
void orc_rf_offsets_linear(uint32_t n, float qmin, float qmax, float bl_scale, float angle,
                           float f_RF, float* offset) {
    const float delta0 = orc_ruler_delta(n, qmin, qmax);
    const float bl2phase = (float)((double)bl_scale / kC * (double)f_RF * kTwoPi);
    const float syncphase = 0, phase = 0, ampl = 1;
    const float phaseoffs = (syncphase - phase);
    const float xcenter = orc_ruler_zerobin(n, qmin, qmax);
    for (uint32_t x = 0; x < n; x++) {
        offset[x] = tanf(angle) * (xcenter - (float)x);
        offset[x] += tanf(angle) * phaseoffs / bl2phase / delta0;
        offset[x] *= ampl;
    }
}

/* src/SM/RFKickMap.cpp:36-57 (non-linear constructor) + :69-75 */
void orc_rf_offsets_nonlinear(uint32_t n, float qmin, float qmax, float pmin, float pmax,
                              float bl_scale, float dE_scale, float revolutionpart, float V_RF,
                              float f_RF, float V0, float* offset) {
    float* axis0 = (float*)malloc(sizeof(float) * n);
    orc_ruler_axis(n, qmin, qmax, axis0);
    const float delta1 = orc_ruler_delta(n, pmin, pmax);
    const float bl2phase = (float)((double)bl_scale / kC * (double)f_RF * kTwoPi);
    const float syncphase = asinf(V0 / V_RF);
    const float phase = syncphase, ampl = 1;
    for (uint32_t x = 0; x < n; x++) {
        const float arg = MADD(axis0[x], bl2phase, phase);
        offset[x] = revolutionpart * MADD(-ampl * V_RF, sinf(arg), V0) / delta1 / dE_scale;
    }
    free(axis0);
}

/* RFKickMap::_calcKick(phase, ampl) (src/SM/RFKickMap.cpp:59-83) as DynamicRFKickMap::apply calls it
 * every step (src/SM/DynamicRFKickMap.cpp:96-105) */
void orc_rf_calc_kick(int linear, uint32_t n, float qmin, float qmax, float pmin, float pmax,
                      float bl_scale, float dE_scale, float angle, float revolutionpart, float V_RF,
                      float f_RF, float V0, float phase, float ampl, float* offset) {
    const float bl2phase = (float)((double)bl_scale / kC * (double)f_RF * kTwoPi);
    if (linear) {
        const float delta0 = orc_ruler_delta(n, qmin, qmax);
        const float syncphase = 0; /* src/SM/RFKickMap.cpp:30 */
        const float phaseoffs = (syncphase - phase);
        const float xcenter = orc_ruler_zerobin(n, qmin, qmax);
        for (uint32_t x = 0; x < n; x++) {
            offset[x] = tanf(angle) * (xcenter - (float)x);
            offset[x] += tanf(angle) * phaseoffs / bl2phase / delta0;
            offset[x] *= ampl;
        }
    } else {
        float* axis0 = (float*)malloc(sizeof(float) * n);
        orc_ruler_axis(n, qmin, qmax, axis0);
        const float delta1 = orc_ruler_delta(n, pmin, pmax);
        for (uint32_t x = 0; x < n; x++) {
            const float arg = MADD(axis0[x], bl2phase, phase);
            offset[x] = revolutionpart * MADD(-ampl * V_RF, sinf(arg), V0) / delta1 / dE_scale;
        }
        free(axis0);
    }
}

/* DynamicRFKickMap::__calcModulation (src/SM/DynamicRFKickMap.cpp:80-94) with the two normal
 * deviates per step supplied by the caller (gauss[2 i], gauss[2 i + 1]; the reference draws them
 * from an unseeded mt19937): out[2 i] = phase, out[2 i + 1] = amplitude factor */
void orc_rf_modulation(uint32_t steps, float phasespread, float amplspread, float modampl,
                       double modtimeincrement, double revolutionpart, float syncphase,
                       const float* gauss, float* out) {
    const float phasenoise_w = (float)(phasespread / sqrt(revolutionpart));
    const float amplnoise_w = (float)(amplspread / sqrt(revolutionpart));
    const float modtimedelta = (float)(kTwoPi * modtimeincrement);
    for (uint32_t i = 0; i < steps; i++) {
        const float phasenoise = gauss[2 * i] * phasenoise_w;
        const float amplnoise = gauss[2 * i + 1] * amplnoise_w;
        const float phasemod = modampl * sinf(modtimedelta * (float)i);
        out[2 * i] = syncphase + phasenoise + phasemod;
        out[2 * i + 1] = 1 + amplnoise;
    }
}

/* src/SM/DriftMap.cpp:21-30 */
void orc_drift_offsets(uint32_t n, float qmin, float qmax, float pmin, float pmax,
                       float dE_scale, const float* slip, uint32_t nslip, float E0,
                       float* offset) {
    float* axis1 = (float*)malloc(sizeof(float) * n);
    orc_ruler_axis(n, pmin, pmax, axis1);
    const float delta0 = orc_ruler_delta(n, qmin, qmax);
    for (uint32_t y = 0; y < n; y++) {
        offset[y] = 0;
        for (uint32_t i = 0; i < nslip; i++) {
            offset[y] = MADD(slip[i] * axis1[y], powf(axis1[y] * dE_scale / E0, (float)i),
                             offset[y]);
        }
        offset[y] /= delta0;
    }
    free(axis1);
}

/* ------------------------------------------------------------------ Fokker-Planck */

/* src/SM/FokkerPlanckMap.cpp:31-118 */
void orc_fp_table(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                  uint32_t derivation, uint32_t* index, float* weight) {
    const uint32_t ip = derivation;
    float* axis1 = (float*)malloc(sizeof(float) * n);
    orc_ruler_axis(n, pmin, pmax, axis1);
    const float d1 = orc_ruler_delta(n, pmin, pmax);
    const float e1_2d = e1 / (2.0f * d1);
    const float e1_6d = e1 / (6.0f * d1);
    const float e1_d2 = e1 / (d1 * d1);
    const float ycenter = orc_ruler_zerobin(n, pmin, pmax);
    const int damp = (fptype != 0 && fptype != 2); /* not none, not diffusion_only */
    const int diff = (fptype != 0 && fptype != 1); /* not none, not damping_only */
    memset(index, 0, sizeof(uint32_t) * n * ip);
    memset(weight, 0, sizeof(float) * n * ip);
#define H(j, k, idx, w) do { index[(j) * ip + (k)] = (idx); weight[(j) * ip + (k)] = (w); } while (0)
#define W(j, k) weight[(j) * ip + (k)]
    if (derivation == 3) {
        for (uint32_t j = 1; j < n - 1; j++) {
            H(j, 0, j - 1, 0);
            H(j, 1, j, 1);
            H(j, 2, j + 1, 0);
            if (damp) {
                const float pos = axis1[j];
                W(j, 0) += -e1_2d * pos;
                W(j, 1) += e1;
                W(j, 2) += e1_2d * pos;
            }
            if (diff) {
                W(j, 0) += e1_d2;
                W(j, 1) += -2.0f * e1_d2;
                W(j, 2) += e1_d2;
            }
        }
        H(n - 1, 0, 0, 0); H(n - 1, 1, 0, 0); H(n - 1, 2, 0, 0);
    } else {
        for (uint32_t j = 2; (float)j < ycenter; j++) {
            const float pos = axis1[j];
            H(j, 0, j - 2, 0);
            H(j, 1, j - 1, 0);
            H(j, 2, j, 1);
            H(j, 3, j + 1, 0);
            if (damp) {
                W(j, 0) += e1_6d * 1.0f * pos;
                W(j, 1) += e1_6d * -6.0f * pos;
                W(j, 2) += MADD(e1_6d * 3.0f, pos, e1);
                W(j, 3) += e1_6d * 2.0f * pos;
            }
            if (diff) {
                W(j, 1) += e1_d2;
                W(j, 2) += -2.0f * e1_d2;
                W(j, 3) += e1_d2;
            }
        }
        for (uint32_t j = f2u(ycenter); j < n - 2; j++) {
            const float pos = axis1[j];
            H(j, 0, j - 1, 0);
            H(j, 1, j, 1);
            H(j, 2, j + 1, 0);
            H(j, 3, j + 2, 0);
            if (damp) {
                W(j, 0) += e1_6d * -2.0f * pos;
                W(j, 1) += MADD(e1_6d * -3.0f, pos, e1);
                W(j, 2) += e1_6d * 6.0f * pos;
                W(j, 3) += e1_6d * -1.0f * pos;
            }
            if (diff) {
                W(j, 0) += e1_d2;
                W(j, 1) += -2.0f * e1_d2;
                W(j, 2) += e1_d2;
            }
        }
        for (uint32_t k = 0; k < 4; k++) {
            H(n - 2, k, 0, 0);
            H(n - 1, k, 0, 0);
        }
    }
#undef H
#undef W
    free(axis1);
}

void orc_fp_weights5(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                     uint32_t derivation, float* w5) {
    const uint32_t ip = derivation;
    uint32_t* index = (uint32_t*)malloc(sizeof(uint32_t) * n * ip);
    float* weight = (float*)malloc(sizeof(float) * n * ip);
    orc_fp_table(n, pmin, pmax, e1, fptype, derivation, index, weight);
    for (uint32_t y = 0; y < n; y++) {
        for (int k = 0; k < 5; k++) w5[y * 5 + k] = 0;
        for (uint32_t j = 0; j < ip; j++) {
            if (weight[y * ip + j] == 0) continue;
            w5[y * 5 + ((int)index[y * ip + j] - (int)y + 2)] = weight[y * ip + j];
        }
    }
    free(index);
    free(weight);
}

/* src/SM/FokkerPlanckMap.cpp:197-211 */
static void fp_apply_table(uint32_t n, uint32_t ip, const uint32_t* index, const float* weight,
                           const float* in, float* out) {
    for (uint32_t x = 0; x < n; x++) {
        const uint32_t offs = x * n;
        for (uint32_t y = 0; y < n; y++) {
            float value = 0;
            for (uint32_t j = 0; j < ip; j++) {
                value = PLAIN(in[offs + index[y * ip + j]], weight[y * ip + j], value);
            }
            out[offs + y] = value;
        }
    }
}

void orc_fp_apply(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                  uint32_t derivation, const float* in, float* out) {
    const uint32_t ip = derivation;
    uint32_t* index = (uint32_t*)malloc(sizeof(uint32_t) * n * ip);
    float* weight = (float*)malloc(sizeof(float) * n * ip);
    orc_fp_table(n, pmin, pmax, e1, fptype, derivation, index, weight);
    fp_apply_table(n, ip, index, weight, in, out);
    free(index);
    free(weight);
}

/* src/SM/FokkerPlanckMap.cpp:214-268 (deterministic modes) */
static void fp_track_table(uint32_t n, uint32_t ip, const uint32_t* index, const float* weight,
                           uint32_t fptrack, const float* grid, float* particles, uint32_t np) {
    for (uint32_t i = 0; i < np; i++) {
        float* px = &particles[2 * i];
        float* py = &particles[2 * i + 1];
        if (fptrack == 1) {
            uint32_t yi = f2u(floorf(*py));
            if (yi > n) yi = n;
            float offset = 0;
            for (uint32_t j = 0; j < ip; j++) {
                const float dy = (float)yi - (float)index[yi * ip + j];
                offset = MADD(dy, weight[yi * ip + j], offset);
            }
            *py = std_max(1.0f, std_min(*py + offset, (float)(n - 1)));
        } else if (fptrack == 2) {
            uint32_t xu = f2u(floorf(*px));
            if (xu > n - 1) xu = n - 1;
            const int32_t xi = (int32_t)xu;
            int32_t yi = (int32_t)floorf(*py);
            if (yi > (int32_t)(n - 1)) yi = (int32_t)(n - 1);
            const uint32_t offs = (uint32_t)xi * n;
            float offset = 0, charge = 0;
            for (uint32_t j = 0; j < ip; j++) {
                const uint32_t idx = index[(uint32_t)yi * ip + j];
                const float w = weight[(uint32_t)yi * ip + j];
                charge = MADD(grid[offs + idx], w, charge);
                offset = MADD(grid[offs + idx] * w, (float)((int32_t)idx - yi), offset);
            }
            offset /= charge;
            *py = std_max(1.0f, std_min(*py + offset, (float)(n - 1)));
        }
    }
}

void orc_fp_track(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                  uint32_t derivation, uint32_t fptrack, const float* grid, float* particles,
                  uint32_t np) {
    const uint32_t ip = derivation;
    uint32_t* index = (uint32_t*)malloc(sizeof(uint32_t) * n * ip);
    float* weight = (float*)malloc(sizeof(float) * n * ip);
    orc_fp_table(n, pmin, pmax, e1, fptype, derivation, index, weight);
    fp_track_table(n, ip, index, weight, fptrack, grid, particles, np);
    free(index);
    free(weight);
}

/* ------------------------------------------------------------------ rotation */

/* src/SM/RotationMap.cpp:190-260 */
static void rot_hinfo(uint32_t n, uint32_t it, float cos_dt, float sin_dt, const float* axis0,
                      const float* axis1, float d0, float d1, float zb0, float zb1, uint32_t x0,
                      uint32_t y0, uint32_t* index, float* weight) {
    const uint32_t ip = it * it;
    float icq[4], icp[4], smc[16];
    float x1r = MADD(cos_dt, axis0[x0], -(sin_dt * axis1[y0])) / d0 + zb0;
    float y1r = MADD(sin_dt, axis0[x0], cos_dt * axis1[y0]) / d1 + zb1;
    const float xf = modff(x1r, &x1r);
    const float yf = modff(y1r, &y1r);
    const uint32_t x1 = f2u(x1r);
    const uint32_t y1 = f2u(y1r);
    if (x1 < n && y1 < n) {
        orc_calc_coefficients(xf, it, icq);
        orc_calc_coefficients(yf, it, icp);
        for (uint32_t smq = 0; smq < it; smq++) {
            for (uint32_t smp = 0; smp < it; smp++) smc[smp * it + smq] = icq[smp] * icp[smq];
        }
        for (uint32_t j1 = 0; j1 < it; j1++) {
            const uint32_t j0 = y1 + j1 - (it - 1) / 2;
            for (uint32_t i1 = 0; i1 < it; i1++) {
                const uint32_t i0 = x1 + i1 - (it - 1) / 2;
                if (i0 < n && j0 < n) {
                    index[i1 * it + j1] = i0 * n + j0;
                    weight[i1 * it + j1] = smc[i1 * it + j1];
                } else {
                    index[i1 * it + j1] = 0;
                    weight[i1 * it + j1] = 0;
                }
            }
        }
    } else {
        for (uint32_t i = 0; i < ip; i++) {
            index[i] = 0;
            weight[i] = 0;
        }
    }
}

/* src/SM/RotationMap.cpp:12-88 (argument check), :142-174 (apply, CPU path) */
int orc_rotation_apply(uint32_t n, float qmin, float qmax, float pmin, float pmax, float angle,
                       uint32_t it, int clamp, uint32_t rotmapsize, const float* in, float* out) {
    if (clamp && !(it == 4 && rotmapsize > 0)) return 1;
    const uint32_t ip = it * it;
    const float cos_dt = cosf(-angle), sin_dt = sinf(-angle);
    float* axis0 = (float*)malloc(sizeof(float) * n);
    float* axis1 = (float*)malloc(sizeof(float) * n);
    orc_ruler_axis(n, qmin, qmax, axis0);
    orc_ruler_axis(n, pmin, pmax, axis1);
    const float d0 = orc_ruler_delta(n, qmin, qmax), d1 = orc_ruler_delta(n, pmin, pmax);
    const float zb0 = orc_ruler_zerobin(n, qmin, qmax), zb1 = orc_ruler_zerobin(n, pmin, pmax);
    uint32_t index[16];
    float weight[16];
    for (uint32_t q_i = 0; q_i < n; q_i++) {
        for (uint32_t p_i = 0; p_i < n; p_i++) {
            const uint32_t i = q_i * n + p_i;
            rot_hinfo(n, it, cos_dt, sin_dt, axis0, axis1, d0, d1, zb0, zb1, q_i, p_i, index,
                      weight);
            float v = 0;
            for (uint32_t j = 0; j < ip; j++) v = PLAIN(in[index[j]], weight[j], v);
            if (clamp) {
                float ceil = FLT_MIN, flor = FLT_MAX;
                for (uint32_t x = 1; x <= 2; x++) {
                    for (uint32_t y = 1; y <= 2; y++) {
                        ceil = std_max(ceil, in[index[x * it + y]]);
                        flor = std_min(flor, in[index[x * it + y]]);
                    }
                }
                v = std_max(std_min(ceil, v), flor);
            }
            out[i] = v;
        }
    }
    free(axis0);
    free(axis1);
    return 0;
}

/* src/SM/RotationMap.cpp:178-188 */
void orc_rotation_track(uint32_t n, float qmin, float qmax, float pmin, float pmax, float angle,
                        float* particles, uint32_t np) {
    const float cos_dt = cosf(-angle), sin_dt = sinf(-angle);
    const float zb0 = orc_ruler_zerobin(n, qmin, qmax), zb1 = orc_ruler_zerobin(n, pmin, pmax);
    for (uint32_t i = 0; i < np; i++) {
        const float x = particles[2 * i], y = particles[2 * i + 1];
        particles[2 * i] = MADD(sin_dt, (y - zb1), MADD(cos_dt, (x - zb0), 0.0f)) + zb0;
        particles[2 * i + 1] = MADD(-sin_dt, (x - zb0), MADD(cos_dt, (y - zb1), 0.0f)) + zb1;
    }
}

/* ------------------------------------------------------------------ PhaseSpace */

/* src/PS/PhaseSpace.cpp:486-501 */
void orc_simpson_weights(uint32_t n, float delta, float* ws) {
    const float ca = 3.f;
    float dc = 1;
    const float h03 = delta / 3.0f;
    ws[0] = h03;
    for (uint32_t x = 1; x + 1 < n; x++) {
        ws[x] = h03 * (ca + dc);
        dc = -dc;
    }
    ws[n - 1] = h03;
}

/* src/PS/PhaseSpace.cpp:313-321 (std::inner_product, sequential) */
void orc_project_x(uint32_t n, const float* grid, const float* ws, float* proj) {
    for (uint32_t x = 0; x < n; x++) {
        float acc = 0;
        for (uint32_t y = 0; y < n; y++) acc = PLAIN(grid[x * n + y], ws[y], acc);
        proj[x] = acc;
    }
}

/* src/PS/PhaseSpace.cpp:325-340 */
void orc_project_y(uint32_t n, const float* grid, const float* ws, float* proj) {
    for (uint32_t y = 0; y < n; y++) {
        float acc = 0;
        for (uint32_t x = 0; x < n; x++) acc = PLAIN(grid[x * n + y], ws[x], acc);
        proj[y] = acc;
    }
}

/* src/PS/PhaseSpace.cpp:229-238 */
float orc_integrate(uint32_t n, const float* proj, const float* ws) {
    float acc = 0;
    for (uint32_t x = 0; x < n; x++) acc = PLAIN(proj[x], ws[x], acc);
    return acc;
}

/* src/PS/PhaseSpace.cpp:342-372 */
void orc_normalize(uint32_t n, float* grid, float filling_set, float filling) {
    for (size_t i = 0; i < (size_t)n * n; i++) grid[i] *= filling_set / filling;
}

/* src/PS/PhaseSpace.cpp:241-289 */
void orc_moments(uint32_t n, const float* proj, const float* axis, float delta, float filling,
                 float* mean, float* var, float* rms) {
    float avg = 0;
    for (uint32_t i = 0; i < n; i++) avg = PLAIN(proj[i], axis[i], avg);
    avg *= delta / filling;
    float v = 0;
    for (uint32_t i = 0; i < n; i++) {
        /* std::pow(float,int) promotes to double (C++11) */
        const double d = (double)(axis[i] - avg);
        v = (float)((double)proj[i] * (d * d) + (double)v);
    }
    v *= delta / filling;
    *mean = avg;
    *var = v;
    *rms = sqrtf(v);
}

/* src/PS/PhaseSpace.cpp:52-65 (constructor without data), :457-484, then src/main.cpp:528-532 */
void orc_initial_grid(uint32_t n, float qmin, float qmax, float pmin, float pmax, double zoom,
                      float* grid) {
    float* a0 = (float*)malloc(sizeof(float) * n);
    float* a1 = (float*)malloc(sizeof(float) * n);
    float* p0 = (float*)malloc(sizeof(float) * n);
    float* p1 = (float*)malloc(sizeof(float) * n);
    float* ws = (float*)malloc(sizeof(float) * n);
    float* proj = (float*)malloc(sizeof(float) * n);
    orc_ruler_axis(n, qmin, qmax, a0);
    orc_ruler_axis(n, pmin, pmax, a1);
    orc_simpson_weights(n, orc_ruler_delta(n, qmin, qmax), ws);
    const double zoom2 = zoom * zoom;
    const float norm = (float)0.398942280401432677939946059934381868L;
    for (uint32_t i = 0; i < n; i++) {
        p0[i] = (float)(norm * exp((-0.5) * a0[i] * a0[i] / zoom2));
        p1[i] = (float)(norm * exp((-0.5) * a1[i] * a1[i] / zoom2));
    }
    for (uint32_t x = 0; x < n; x++) {
        for (uint32_t y = 0; y < n; y++) grid[x * n + y] = p0[x] * p1[y];
    }
    /* createFromProjections: updateXProjection, integrate, normalize */
    orc_project_x(n, grid, ws, proj);
    float filling = orc_integrate(n, proj, ws);
    orc_normalize(n, grid, 1.0f, filling);
    /* constructor tail: updateXProjection, updateYProjection, integrate */
    orc_project_x(n, grid, ws, proj);
    filling = orc_integrate(n, proj, ws);
    /* main(): updateXProjection, normalize */
    orc_normalize(n, grid, 1.0f, filling);
    free(a0); free(a1); free(p0); free(p1); free(ws); free(proj);
}

/* ------------------------------------------------------------------ FFT (FFTW semantics) */

typedef struct { double re, im; } cd;

static void dft_pow2(cd* a, size_t n, int sign) {
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cd t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t k = 0; k < len / 2; k++) {
            const double ph = sign * kTwoPi * (double)k / (double)len;
            const double wr = cos(ph), wi = sin(ph);
            for (size_t i = k; i < n; i += len) {
                cd* u = &a[i];
                cd* v = &a[i + len / 2];
                const double tr = v->re * wr - v->im * wi, ti = v->re * wi + v->im * wr;
                v->re = u->re - tr; v->im = u->im - ti;
                u->re += tr; u->im += ti;
            }
        }
    }
}

static void dft(cd* a, size_t n, int sign) {
    if ((n & (n - 1)) == 0) { dft_pow2(a, n, sign); return; }
    cd* out = (cd*)malloc(sizeof(cd) * n);
    for (size_t k = 0; k < n; k++) {
        double sr = 0, si = 0;
        for (size_t j = 0; j < n; j++) {
            const double ph = sign * kTwoPi * (double)((k * j) % n) / (double)n;
            const double wr = cos(ph), wi = sin(ph);
            sr += a[j].re * wr - a[j].im * wi;
            si += a[j].re * wi + a[j].im * wr;
        }
        out[k].re = sr; out[k].im = si;
    }
    memcpy(a, out, sizeof(cd) * n);
    free(out);
}

/* fftwf_plan_dft_r2c_1d: bins 0..n/2 written (interleaved re,im), rest untouched */
static void fft_r2c(const float* in, size_t n, float* out) {
    cd* a = (cd*)malloc(sizeof(cd) * n);
    for (size_t i = 0; i < n; i++) { a[i].re = in[i]; a[i].im = 0; }
    dft(a, n, -1);
    for (size_t k = 0; k <= n / 2; k++) { out[2 * k] = (float)a[k].re; out[2 * k + 1] = (float)a[k].im; }
    free(a);
}

/* fftwf_plan_dft_c2r_1d: reads bins 0..n/2, imaginary parts of bins 0 and n/2 ignored */
static void fft_c2r(const float* in, size_t n, float* out) {
    cd* a = (cd*)malloc(sizeof(cd) * n);
    a[0].re = in[0]; a[0].im = 0;
    for (size_t k = 1; k <= n / 2; k++) {
        double re = in[2 * k], im = in[2 * k + 1];
        if (2 * k == n) im = 0;
        a[k].re = re; a[k].im = im;
        a[n - k].re = re; a[n - k].im = -im;
    }
    dft(a, n, +1);
    for (size_t i = 0; i < n; i++) out[i] = (float)a[i].re;
    free(a);
}

/* ------------------------------------------------------------------ ElectricField */

/* src/PS/ElectricField.cpp:186-263 (single bunch, bucket 0) */
void orc_wake_potential(uint32_t n, const float* profile, const float* z, uint32_t nmax,
                        float wakescaling, float* wake, float* formfactor, float* padded_wake) {
    float* bp = (float*)calloc(nmax, sizeof(float));            /* _bp_padded (zero filled) */
    float* ff = (float*)calloc(2 * (size_t)nmax, sizeof(float)); /* _formfactor */
    float* wl = (float*)calloc(2 * (size_t)nmax, sizeof(float)); /* _wakelosses */
    float* wp = (float*)calloc(nmax, sizeof(float));            /* _wakepotential_padded */
    memcpy(bp, profile, sizeof(float) * n);                     /* padBunchProfiles :255-263 */
    fft_r2c(bp, nmax, ff);                                      /* :227 */
    for (uint32_t i = 0; i < nmax / 2; i++) {                   /* :229-231, complex<float> product */
        const float a = z[2 * i], b = z[2 * i + 1], c = ff[2 * i], d = ff[2 * i + 1];
        wl[2 * i] = MADD(-b, d, a * c);
        wl[2 * i + 1] = MADD(b, c, a * d);
    }
    fft_c2r(wl, nmax, wp);                                      /* :234 */
    for (uint32_t x = 0; x < n; x++) wake[x] = wakescaling * wp[x]; /* :236-241 */
    if (formfactor) memcpy(formfactor, ff, sizeof(float) * 2 * (nmax / 2 + 1));
    if (padded_wake) memcpy(padded_wake, wp, sizeof(float) * nmax);
    free(bp); free(ff); free(wl); free(wp);
}

/* The same for a TRAIN of nb bunches (src/PS/ElectricField.cpp:186-263 as written, PhaseSpace::nb > 1):
 * padBunchProfiles() copies the profile of bunch b to _bp_padded + bucket[b] * spacing_bins (:255-263), ONE
 * transform pair of nmax = impedance->nFreqs() bins couples all bunches (circularly), and bunch b reads its
 * wake potential back at the same offset (:236-241).  profiles / wake are [nb][n].  Returns 0, or -1 if a
 * bunch does not fit the padded array (the reference would write past it). */
int orc_train_wake_potential(uint32_t n, uint32_t nb, const float* profiles, const uint32_t* bucket,
                             uint32_t spacing_bins, const float* z, uint32_t nmax, float wakescaling,
                             float* wake, float* padded_wake) {
    for (uint32_t b = 0; b < nb; b++)
        if ((uint64_t)bucket[b] * spacing_bins + n > nmax) return -1;
    float* bp = (float*)calloc(nmax, sizeof(float));
    float* ff = (float*)calloc(2 * (size_t)nmax, sizeof(float));
    float* wl = (float*)calloc(2 * (size_t)nmax, sizeof(float));
    float* wp = (float*)calloc(nmax, sizeof(float));
    for (uint32_t b = 0; b < nb; b++) /* padBunchProfiles :255-263 (later bunches overwrite overlaps) */
        memcpy(bp + (size_t)bucket[b] * spacing_bins, profiles + (size_t)b * n, sizeof(float) * n);
    fft_r2c(bp, nmax, ff);
    for (uint32_t i = 0; i < nmax / 2; i++) {
        const float a = z[2 * i], bb = z[2 * i + 1], c = ff[2 * i], d = ff[2 * i + 1];
        wl[2 * i] = MADD(-bb, d, a * c);
        wl[2 * i + 1] = MADD(bb, c, a * d);
    }
    fft_c2r(wl, nmax, wp);
    for (uint32_t b = 0; b < nb; b++)
        for (uint32_t x = 0; x < n; x++)
            wake[(size_t)b * n + x] = wakescaling * wp[(size_t)bucket[b] * spacing_bins + x];
    if (padded_wake) memcpy(padded_wake, wp, sizeof(float) * nmax);
    free(bp); free(ff); free(wl); free(wp);
    return 0;
}

/* src/PS/ElectricField.cpp:133-184 with the frequency ruler of the constructor (:28-31) */
void orc_csr_power(uint32_t n, const float* profile, const float* z, uint32_t nmax, float dq,
                   float f_scale, float cutoff, float* spectrum, float* power) {
    float* bp = (float*)calloc(nmax, sizeof(float));
    float* ff = (float*)calloc(2 * (size_t)nmax, sizeof(float));
    memcpy(bp, profile, sizeof(float) * n);
    fft_r2c(bp, nmax, ff);
    const float fmax_axis = 1 / dq;
    const float delta_f = orc_ruler_delta(nmax, 0, fmax_axis);
    const float formfactorrenorm = dq * dq;
    float intensity = 0;
    for (uint32_t i = 0; i < nmax; i++) {
        float renorm = formfactorrenorm;
        if (cutoff > 0) {
            const float fi = MADD((float)i, delta_f, 0.0f);
            const double r = (double)(f_scale * fi / cutoff);
            renorm = (float)((double)renorm * (1 - exp(-(r * r))));
        }
        const float nrm = MADD(ff[2 * i], ff[2 * i], ff[2 * i + 1] * ff[2 * i + 1]);
        const float s = renorm * z[2 * i] * nrm;
        if (spectrum) spectrum[i] = s;
        intensity = MADD(delta_f, s, intensity);
    }
    *power = intensity;
    free(bp); free(ff);
}

/* ------------------------------------------------------------------ impedances */

/* src/Z/FreeSpaceCSR.cpp:18-43 */
void orc_z_free_space(size_t n, float f_rev, float f_max, float* z) {
    const float delta = f_max / f_rev / (float)(n - 1);
    for (size_t i = 0; i <= n / 2; i++) {
        const float s = powf((float)i * delta, (float)(1.0 / 3.0));
        z[2 * i] = 306.3f * s;
        z[2 * i + 1] = 176.9f * s;
    }
    for (size_t i = n / 2 + 1; i < n; i++) z[2 * i] = z[2 * i + 1] = 0;
}

/* modified Bessel functions of fractional order by quadrature (independent of any library):
 *   K_v(z)           = int_0^inf exp(-z cosh t) cosh(v t) dt
 *   exp(-z) I_v(z)   = 1/pi int_0^pi exp(z (cos th - 1)) cos(v th) dth
 *                      - sin(v pi)/pi int_0^inf exp(-z (cosh t + 1) - v t) dt          */
static double bessel_k(double v, double z) {
    /* trapezoid rule, double-exponentially decaying integrand */
    const double h = 1.0 / 64.0;
    double s = 0.5 * exp(-z);
    for (int k = 1; k < 64 * 40; k++) {
        const double t = k * h;
        const double e = -z * cosh(t);
        if (e < -745.0) break;
        s += exp(e) * cosh(v * t);
    }
    return s * h;
}
static double bessel_i_scaled(double v, double z) {
    /* composite Simpson on [0,pi] (smooth, non-periodic) */
    const int m = 20000;
    const double h = kPi / m;
    double s = 0;
    for (int k = 0; k <= m; k++) {
        const double th = k * h;
        const double f = exp(z * (cos(th) - 1.0)) * cos(v * th);
        s += f * ((k == 0 || k == m) ? 1.0 : ((k & 1) ? 4.0 : 2.0));
    }
    const double first = s * h / 3.0 / kPi;
    /* second integrand is not even in t: composite Simpson (trapezoid would be O(h^2)) */
    const double h2 = 1.0 / 256.0;
    int m2 = 2;
    while (m2 < 256 * 80 && -z * (cosh(m2 * h2) + 1.0) - v * m2 * h2 > -745.0) m2 += 2;
    double s2 = 0;
    for (int k = 0; k <= m2; k++) {
        const double t = k * h2;
        const double f = exp(-z * (cosh(t) + 1.0) - v * t);
        s2 += f * ((k == 0 || k == m2) ? 1.0 : ((k & 1) ? 4.0 : 2.0));
    }
    s2 /= 3.0;
    return first - sin(v * kPi) / kPi * s2 * h2;
}

int orc_airy(double x, double* ai, double* aip, double* bi, double* bip) {
    const double sqrt3 = 1.732050807568877293527446341505872367;
    if (x == 0) {
        *ai = 0.355028053887817239260063186004183176;
        *aip = -0.258819403792806798405183560189203963;
        *bi = *ai * sqrt3;
        *bip = -*aip * sqrt3;
        return 0;
    }
    const double zt = 2.0 / 3.0 * x * sqrt(x);
    const double k13 = (zt > 740) ? 0 : bessel_k(1.0 / 3.0, zt);
    const double k23 = (zt > 740) ? 0 : bessel_k(2.0 / 3.0, zt);
    *ai = sqrt(x / 3.0) / kPi * k13;
    *aip = -x / (kPi * sqrt3) * k23;
    if (zt + 0.25 * log(x) - 0.5723649429247001 > 709.78) {
        *bi = *bip = INFINITY;
        return 1;
    }
    const double ez = exp(zt);
    const double i13 = bessel_i_scaled(1.0 / 3.0, zt) * ez;
    const double i23 = bessel_i_scaled(2.0 / 3.0, zt) * ez;
    /* I_{-v} = I_v + 2/pi sin(v pi) K_v */
    *bi = sqrt(x / 3.0) * (2 * i13 + 2.0 / kPi * (sqrt3 / 2) * k13);
    *bip = x / sqrt3 * (2 * i23 + 2.0 / kPi * (sqrt3 / 2) * k23);
    return 0;
}

/* src/Z/ParallelPlatesCSR.cpp:25-70 */
void orc_z_parallel_plates(size_t nfreqs, float f0, float f_max, double g, float* z) {
    const double Z0 = 1 / (kEps0 * kC); /* physcons::Z0, inc/defines.hpp:54 */
    const double delta = f_max / f0 / (nfreqs - 1.0);
    const double r_bend = kC / (2 * kPi * f0);
    memset(z, 0, sizeof(float) * 2 * nfreqs);
    for (size_t i = 1; i <= nfreqs / 2; i++) {
        double zr = 0, zi = 0;
        const double n = i * delta;
        const double m = n * pow(g / r_bend, 3. / 2.);
        const uint32_t maxp = (uint32_t)(2 * m * pow(r_bend / g, 3. / 2.) * f0 * g / kC);
        const double b = pow(m, -4. / 3.);
        for (uint32_t p = 1; p <= maxp; p += 2) {
            const double u = kPiSqr * p * p / pow(2, 2. / 3.) * b;
            double ai, aip, bi, bip;
            if (orc_airy(u, &ai, &aip, &bi, &bip)) break; /* overflow: stop summing (:54-57) */
            /* Ai'(Ai' - j Bi') + u Ai (Ai - j Bi) */
            zr += aip * aip + (u * ai) * ai;
            zi += -(aip * bip) - (u * ai) * bi;
        }
        const double f = 4.0 * b * n * g / r_bend * kPiSqr * pow(2, 1. / 3.) * Z0;
        z[2 * i] = (float)(zr * f);
        z[2 * i + 1] = (float)(zi * f);
    }
}

/* src/Z/ResistiveWall.cpp:24-57 */
void orc_z_resistive_wall(size_t n, float f0, float f_max, double L, double s, double xi,
                          double b, float* z) {
    const double mu_r = (1 + xi);
    const float z1 = (float)(sqrt(kZ0wall * mu_r * f0 / s / kPi / kC) * L / 2 / b);
    const float delta = (float)(f_max / f0 / (n - 1.0));
    for (size_t i = 0; i <= n / 2; i++) {
        const float r = sqrtf((float)i * delta);
        z[2 * i] = z1 * r;
        z[2 * i + 1] = -z1 * r;
    }
    for (size_t i = n / 2 + 1; i < n; i++) z[2 * i] = z[2 * i + 1] = 0;
}

/* src/Z/ConstImpedance.cpp:18-28 */
void orc_z_const(size_t n, float re, float im, float* z) {
    for (size_t i = 0; i < n; i++) {
        z[2 * i] = (i < n / 2) ? re : 0;
        z[2 * i + 1] = (i < n / 2) ? im : 0;
    }
}

/* src/Z/CollimatorImpedance.cpp:12-25 */
void orc_z_collimator(size_t n, double outer, double inner, float* z) {
    orc_z_const(n, (float)(kZ0wall / kPi * log(outer / inner)), 0.0f, z);
}

/* src/Z/ImpedanceFactory.cpp:21-84 (no impedance file); returns 0 for "no impedance is used" */
int orc_make_impedance(size_t nfreqs, float fmax, double R_bend, double f_rev, double gap,
                       int use_csr, double s, double xi, double collimator, float* z) {
    const double f0 = kC / (kTwoPi * R_bend);
    float* t = (float*)malloc(sizeof(float) * 2 * nfreqs);
    int changed = 0;
    memset(z, 0, sizeof(float) * 2 * nfreqs);
    if (gap != 0) {
        if (use_csr) {
            changed = 1;
            if (gap > 0) orc_z_parallel_plates(nfreqs, (float)f0, fmax, gap, t);
            else orc_z_free_space(nfreqs, (float)f0, fmax, t);
            for (size_t i = 0; i < 2 * nfreqs; i++) z[i] += t[i];
        }
        const double radius = fabs(gap / 2);
        if (s > 0 && xi >= -1) {
            changed = 1;
            orc_z_resistive_wall(nfreqs, (float)f_rev, fmax, kC / f_rev, s, xi, radius, t);
            for (size_t i = 0; i < 2 * nfreqs; i++) z[i] += t[i];
        }
        if (0 < collimator && collimator < radius) {
            changed = 1;
            orc_z_collimator(nfreqs, radius, collimator, t);
            for (size_t i = 0; i < 2 * nfreqs; i++) z[i] += t[i];
        }
    }
    free(t);
    return changed;
}

/* ------------------------------------------------------------------ scalar derivation */

static uint64_t upper_power_of_two(uint64_t v) { /* src/HelperFunctions.cpp:178-189 */
    v--;
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v |= v >> 32;
    v++;
    return v;
}

/* src/main.cpp:163-355,663-664,682 and the constructor-derived members listed in ivs_derived */
void orc_derive(const ivs_ring_params* o, ivs_derived* d) {
    memset(d, 0, sizeof(*d));
    const uint32_t ps_bins = o->grid_size;
    const float pqsize = o->pq_size;
    const float qcenter = -o->shift_x * pqsize / (float)(ps_bins - 1);
    const float pcenter = -o->shift_y * pqsize / (float)(ps_bins - 1);
    const float pqhalf = pqsize / 2;
    d->qmax = qcenter + pqhalf;
    d->qmin = qcenter - pqhalf;
    d->pmax = pcenter + pqhalf;
    d->pmin = pcenter - pqhalf;
    const double sE = o->sE, E0 = o->E0;
    const double dE = sE * E0;
    const double f_rev = (double)o->f_rev;
    const double R_bend = (o->R_bend > 0) ? o->R_bend : kC / (kTwoPi * f_rev);
    const double harmonic_number = (double)o->harmonic;
    const double f_RF = f_rev * harmonic_number;
    const double bunchspacing = 1.0 / f_RF;
    const double V_RF = o->V_RF;
    const double lorentzgamma = E0 / kMe;
    const double V0 = kE * pow(lorentzgamma, 4) / (3 * kEps0 * R_bend);
    const double W0 = V0 * kE;
    const double V_eff = sqrt(V_RF * V_RF - V0 * V0);
    double fs = (double)o->f_s;
    double alpha0 = (double)o->alpha0;
    if (fs == 0) {
        fs = f_rev * sqrt(alpha0 * harmonic_number * V_eff / (kTwoPi * E0));
    } else {
        const int sgn = (fs == 0) ? 0 : ((fs < 0) ? -1 : 1);
        alpha0 = sgn * kTwoPi * E0 / (harmonic_number * V_eff) * pow(fs / f_rev, 2);
    }
    const float alpha[3] = {(float)alpha0, o->alpha1, o->alpha2};
    const double bl = kC * dE / harmonic_number / pow(f_rev, 2.0) / V_eff * fs;
    const double Ib = 0.0 + (double)o->bunch_current;
    const double Qb = Ib / f_rev;
    const double steps = (o->steps_per_Trev > 0) ? o->steps_per_Trev * f_rev / fs
                                                 : (double)(o->steps_per_Ts > 1u ? o->steps_per_Ts : 1u);
    const double calc_damp = E0 * kE / W0 / f_rev;
    const double t_damp = (o->t_damp < 0) ? calc_damp : o->t_damp;
    const double dt = 1.0 / (fs * steps);
    const double revolutionpart = f_rev * dt;
    const double spacing_ps = (bunchspacing * kC / bl / pqsize);
    const uint32_t spacing_bins = (uint32_t)round(ps_bins * spacing_ps);
    const float fmax = (float)(ps_bins * kC / (pqsize * bl));
    const double padding = (o->padding > 1.0) ? o->padding : 1.0;
    size_t padded_bins = (size_t)ceil(ps_bins * padding);
    if (o->round_padding) padded_bins = upper_power_of_two(padded_bins);
    const float angle = (float)(kTwoPi / steps);
    const float e1 = (t_damp > 0) ? (float)(2.0 / (fs * t_damp * steps)) : 0;

    d->dE = dE; d->R_bend = R_bend; d->f_rev = f_rev; d->f_RF = f_RF; d->V0 = V0;
    d->V_eff = V_eff; d->fs = fs; d->alpha0 = alpha0; d->bl = bl; d->Ib = Ib; d->Qb = Qb;
    d->steps = steps; d->t_damp = t_damp; d->dt = dt; d->revolutionpart = revolutionpart;
    d->padded_bins = (uint32_t)padded_bins;
    d->spacing_bins = spacing_bins;
    d->laststep_per_Ts = (uint32_t)ceil(steps);
    d->angle = angle; d->fmax = fmax; d->e1 = e1;
    d->slip[0] = angle;
    d->slip[1] = alpha[1] / alpha[0] * angle;
    d->slip[2] = alpha[2] / alpha[0] * angle;
    d->dq = orc_ruler_delta(ps_bins, d->qmin, d->qmax);
    d->dp = orc_ruler_delta(ps_bins, d->pmin, d->pmax);
    d->zerobin_x = orc_ruler_zerobin(ps_bins, d->qmin, d->qmax);
    d->zerobin_y = orc_ruler_zerobin(ps_bins, d->pmin, d->pmax);
    d->bl_f = (float)bl;   /* Ruler scale map holds meshaxis_t (src/PS/PhaseSpace.cpp:151-166) */
    d->dE_f = (float)dE;
    /* inc/PS/ElectricField.hpp:99-100, src/PS/ElectricField.cpp:58 */
    const float raw = (float)(Ib * dt * kC / (double)d->bl_f / ((double)d->dp * sE * E0));
    d->wakescaling = raw / (float)padded_bins;
}

/* ------------------------------------------------------------------ whole simulation */

struct orc_sim {
    ivs_ring_params p;
    ivs_derived d;
    uint32_t n, it, fp_ip, np, nmax;
    int has_wake, has_fp;
    float *t1, *t2, *t3, *proj, *ws, *wake;
    float *rf, *drift, *fp5;
    uint32_t *rf_idx, *dr_idx, *wk_idx, *fp_idx;
    float *rf_w, *dr_w, *wk_w, *fp_w;
    float *zwake, *zrad;
    float* particles;
    float *axis0, *axis1;
    float filling;
    float* rf_seq;             /* DynamicRFKickMap: (phase, amplitude) per step, NULL = static RF */
    uint32_t rf_seq_len, rf_seq_pos;
};

static void sim_rebuild_tables(orc_sim* s) {
    orc_kick_update_sm(s->rf, s->n, s->n, s->it, s->rf_idx, s->rf_w);
    orc_kick_update_sm(s->drift, s->n, s->n, s->it, s->dr_idx, s->dr_w);
}

/* object wiring of src/main.cpp:478-763.  zwake / zrad: impedance tables computed earlier by
 * orc_make_impedance with the same parameters (the Airy quadrature of the parallel-plates model
 * takes minutes; tests keep the tables in tests/golden/), NULL = compute them here */
orc_sim* orc_sim_create_z(const ivs_ring_params* p, const float* grid, const float* particles,
                          uint32_t np, const float* zwake, const float* zrad) {
    orc_sim* s = (orc_sim*)calloc(1, sizeof(orc_sim));
    s->p = *p;
    orc_derive(p, &s->d);
    const ivs_derived* d = &s->d;
    const uint32_t n = p->grid_size;
    s->n = n;
    s->it = p->interpolation;
    s->nmax = d->padded_bins;
    const size_t nn = (size_t)n * n;
    s->t1 = (float*)malloc(sizeof(float) * nn);
    s->t2 = (float*)malloc(sizeof(float) * nn);
    s->t3 = (float*)malloc(sizeof(float) * nn);
    s->proj = (float*)malloc(sizeof(float) * n);
    s->ws = (float*)malloc(sizeof(float) * n);
    s->wake = (float*)calloc(n, sizeof(float));
    s->rf = (float*)malloc(sizeof(float) * n);
    s->drift = (float*)malloc(sizeof(float) * n);
    s->fp5 = (float*)malloc(sizeof(float) * n * 5);
    s->rf_idx = (uint32_t*)malloc(sizeof(uint32_t) * n * 4);
    s->dr_idx = (uint32_t*)malloc(sizeof(uint32_t) * n * 4);
    s->wk_idx = (uint32_t*)malloc(sizeof(uint32_t) * n * 4);
    s->fp_idx = (uint32_t*)malloc(sizeof(uint32_t) * n * 4);
    s->rf_w = (float*)malloc(sizeof(float) * n * 4);
    s->dr_w = (float*)malloc(sizeof(float) * n * 4);
    s->wk_w = (float*)malloc(sizeof(float) * n * 4);
    s->fp_w = (float*)malloc(sizeof(float) * n * 4);
    s->zwake = (float*)calloc(2 * (size_t)s->nmax, sizeof(float));
    s->zrad = (float*)calloc(2 * (size_t)s->nmax, sizeof(float));
    s->axis0 = (float*)malloc(sizeof(float) * n);
    s->axis1 = (float*)malloc(sizeof(float) * n);
    orc_ruler_axis(n, d->qmin, d->qmax, s->axis0);
    orc_ruler_axis(n, d->pmin, d->pmax, s->axis1);
    orc_simpson_weights(n, d->dq, s->ws);
    if (grid) memcpy(s->t1, grid, sizeof(float) * nn);
    else orc_initial_grid(n, d->qmin, d->qmax, d->pmin, d->pmax, p->zoom, s->t1);
    memcpy(s->t2, s->t1, sizeof(float) * nn);
    memcpy(s->t3, s->t1, sizeof(float) * nn);
    /* RF map, src/main.cpp:640-649 */
    if (p->linear_rf) {
        orc_rf_offsets_linear(n, d->qmin, d->qmax, d->bl_f, d->angle, (float)d->f_RF, s->rf);
    } else {
        orc_rf_offsets_nonlinear(n, d->qmin, d->qmax, d->pmin, d->pmax, d->bl_f, d->dE_f,
                                 (float)d->revolutionpart, (float)d->V_eff, (float)d->f_RF,
                                 (float)d->V0, s->rf);
    }
    /* drift map, :663-679 */
    orc_drift_offsets(n, d->qmin, d->qmax, d->pmin, d->pmax, d->dE_f, d->slip, 3, (float)p->E0,
                      s->drift);
    sim_rebuild_tables(s);
    /* Fokker-Planck or Identity, :682-707 */
    s->has_fp = d->e1 > 0;
    s->fp_ip = p->derivation;
    if (s->has_fp) {
        orc_fp_table(n, d->pmin, d->pmax, d->e1, p->fp_type, p->derivation, s->fp_idx, s->fp_w);
        orc_fp_weights5(n, d->pmin, d->pmax, d->e1, p->fp_type, p->derivation, s->fp5);
    } else {
        for (uint32_t y = 0; y < n; y++) for (int k = 0; k < 5; k++) s->fp5[y * 5 + k] = (k == 2);
    }
    /* impedances, :717-727 */
    if (zwake) {
        memcpy(s->zwake, zwake, sizeof(float) * 2 * s->nmax);
        s->has_wake = 1;
    } else {
        s->has_wake = orc_make_impedance(s->nmax, d->fmax, d->R_bend, d->f_rev, p->gap, p->use_csr,
                                         p->wall_conductivity, p->wall_susceptibility, p->collimator,
                                         s->zwake);
    }
    if (zrad) memcpy(s->zrad, zrad, sizeof(float) * 2 * s->nmax);
    else orc_make_impedance(s->nmax, d->fmax, d->R_bend, d->f_rev, (p->gap > 0) ? p->gap : -1, 1, 0, 0,
                            0, s->zrad);
    s->np = np;
    if (np) {
        s->particles = (float*)malloc(sizeof(float) * 2 * np);
        memcpy(s->particles, particles, sizeof(float) * 2 * np);
    }
    orc_project_x(n, s->t1, s->ws, s->proj); /* :893-894 */
    s->filling = orc_integrate(n, s->proj, s->ws);
    return s;
}

orc_sim* orc_sim_create(const ivs_ring_params* p, const float* grid, const float* particles,
                        uint32_t np) {
    return orc_sim_create_z(p, grid, particles, np, NULL, NULL);
}

void orc_sim_destroy(orc_sim* s) {
    if (!s) return;
    free(s->t1); free(s->t2); free(s->t3); free(s->proj); free(s->ws); free(s->wake);
    free(s->rf); free(s->drift); free(s->fp5);
    free(s->rf_idx); free(s->dr_idx); free(s->wk_idx); free(s->fp_idx);
    free(s->rf_w); free(s->dr_w); free(s->wk_w); free(s->fp_w);
    free(s->zwake); free(s->zrad); free(s->particles); free(s->axis0); free(s->axis1); free(s->rf_seq);
    free(s);
}

/* loop body of src/main.cpp:948-1053 */
void orc_sim_step(orc_sim* s, uint32_t nsteps) {
    const uint32_t n = s->n;
    const size_t nn = (size_t)n * n;
    for (uint32_t i = 0; i < nsteps; i++) {
        if (s->has_wake) { /* wkm->update(), src/SM/WakePotentialMap.cpp:34-50 */
            orc_wake_potential(n, s->proj, s->zwake, s->nmax, s->d.wakescaling, s->wake, NULL, NULL);
            orc_kick_update_sm(s->wake, n, n, s->it, s->wk_idx, s->wk_w);
        }
        s->filling = orc_integrate(n, s->proj, s->ws);
        if (s->has_wake) {
            kick_apply_table(n, s->it, 1, s->wk_idx, s->wk_w, s->t1, s->t2);
            if (s->np) orc_kick_track(n, 1, s->wake, s->particles, s->np);
        } else {
            memcpy(s->t2, s->t1, sizeof(float) * nn); /* Identity */
        }
        if (s->rf_seq && s->rf_seq_pos < s->rf_seq_len) {
            /* DynamicRFKickMap::apply: _calcKick(next modulation), then KickMap::apply */
            const ivs_derived* d = &s->d;
            orc_rf_calc_kick((int)s->p.linear_rf, n, d->qmin, d->qmax, d->pmin, d->pmax, d->bl_f, d->dE_f,
                             d->angle, (float)d->revolutionpart, (float)d->V_eff, (float)d->f_RF,
                             (float)d->V0, s->rf_seq[2 * s->rf_seq_pos], s->rf_seq[2 * s->rf_seq_pos + 1], s->rf);
            orc_kick_update_sm(s->rf, n, n, s->it, s->rf_idx, s->rf_w);
            s->rf_seq_pos++;
        }
        kick_apply_table(n, s->it, 1, s->rf_idx, s->rf_w, s->t2, s->t1);
        if (s->np) orc_kick_track(n, 1, s->rf, s->particles, s->np);
        kick_apply_table(n, s->it, 0, s->dr_idx, s->dr_w, s->t1, s->t3);
        if (s->np) orc_kick_track(n, 0, s->drift, s->particles, s->np);
        if (s->has_fp) {
            fp_apply_table(n, s->fp_ip, s->fp_idx, s->fp_w, s->t3, s->t1);
            /* FokkerPlanckMap::applyTo reads _in (= grid_t3) */
            if (s->np) fp_track_table(n, s->fp_ip, s->fp_idx, s->fp_w, s->p.fp_track, s->t3,
                                      s->particles, s->np);
        } else {
            memcpy(s->t1, s->t3, sizeof(float) * nn);
        }
        orc_project_x(n, s->t1, s->ws, s->proj);
    }
}

/* makes the RF map a DynamicRFKickMap with the given (phase, amplitude) queue, one entry per step */
void orc_sim_set_rf_modulation(orc_sim* s, const float* phase_ampl, uint32_t count) {
    free(s->rf_seq);
    s->rf_seq = (float*)malloc(sizeof(float) * 2 * (count ? count : 1));
    memcpy(s->rf_seq, phase_ampl, sizeof(float) * 2 * count);
    s->rf_seq_len = count;
    s->rf_seq_pos = 0;
}

void orc_sim_derived(const orc_sim* s, ivs_derived* d) { *d = s->d; }

void orc_sim_get(const orc_sim* s, float* grid, float* profile, float* wake, float* particles) {
    const uint32_t n = s->n;
    if (grid) memcpy(grid, s->t1, sizeof(float) * (size_t)n * n);
    if (profile) memcpy(profile, s->proj, sizeof(float) * n);
    if (wake) memcpy(wake, s->wake, sizeof(float) * n);
    if (particles && s->np) memcpy(particles, s->particles, sizeof(float) * 2 * s->np);
}

void orc_sim_tables(const orc_sim* s, float* rf, float* drift, float* fp5, float* zwake,
                    float* zrad) {
    const uint32_t n = s->n;
    if (rf) memcpy(rf, s->rf, sizeof(float) * n);
    if (drift) memcpy(drift, s->drift, sizeof(float) * n);
    if (fp5) memcpy(fp5, s->fp5, sizeof(float) * n * 5);
    if (zwake) memcpy(zwake, s->zwake, sizeof(float) * 2 * s->nmax);
    if (zrad) memcpy(zrad, s->zrad, sizeof(float) * 2 * s->nmax);
}

void orc_sim_set_tables(orc_sim* s, const float* rf, const float* drift, const float* fp5,
                        const float* zwake) {
    const uint32_t n = s->n;
    if (rf) memcpy(s->rf, rf, sizeof(float) * n);
    if (drift) memcpy(s->drift, drift, sizeof(float) * n);
    sim_rebuild_tables(s);
    if (fp5) {
        memcpy(s->fp5, fp5, sizeof(float) * n * 5);
        /* rebuild a 5-tap index/weight table (ascending order is the reference's tap order) */
        s->fp_ip = 4;
        for (uint32_t y = 0; y < n; y++) {
            /* lower-half rows carry w[+2] == 0, upper-half rows w[-2] == 0 */
            const int base = (fp5[y * 5 + 4] == 0) ? -2 : -1;
            for (int k = 0; k < 4; k++) {
                const int rel = base + k;
                int idx = (int)y + rel;
                float w = fp5[y * 5 + rel + 2];
                if (idx < 0 || idx >= (int)n) { idx = 0; w = 0; }
                s->fp_idx[y * 4 + k] = (uint32_t)idx;
                s->fp_w[y * 4 + k] = w;
            }
        }
        s->has_fp = 1;
    }
    if (zwake) memcpy(s->zwake, zwake, sizeof(float) * 2 * s->nmax);
}

/* out-step observables, src/main.cpp:961-1033 */
void orc_sim_observe(orc_sim* s, float* obs, float* csr_spectrum, float* projy_out) {
    const uint32_t n = s->n;
    float* projy = (float*)malloc(sizeof(float) * n);
    s->filling = orc_integrate(n, s->proj, s->ws);
    obs[0] = s->filling;
    orc_moments(n, s->proj, s->axis0, s->d.dq, s->filling, &obs[1], &obs[2], &obs[3]);
    orc_project_y(n, s->t1, s->ws, projy);
    orc_moments(n, projy, s->axis1, s->d.dp, s->filling, &obs[4], &obs[5], &obs[6]);
    const float f_scale = (float)(kC / (double)s->d.bl_f);
    orc_csr_power(n, s->proj, s->zrad, s->nmax, s->d.dq, f_scale, s->p.f_cutoff, csr_spectrum,
                  &obs[7]);
    if (projy_out) memcpy(projy_out, projy, sizeof(float) * n);
    free(projy);
}
