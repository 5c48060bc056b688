// exact_math.cuh -- bit-exact scalar building blocks of the per-frame parameter path.
//
// The reference (networks/warping_2dof_alignment.py:35-58, :125-140) derives every frame's
// rotation / homography / canvas scale through a chain of tiny torch ops whose fp32 rounding
// is decided by the libraries torch dispatches to.  One ulp of drift in any of them moves
// most sampling coordinates of the frame by an ulp, which is visible at the 1e-4 level in the
// warped depth (DESIGN.md "Why bit-exact parameters").  So the device code restates those
// roundings exactly:
//
//   * 3-term dot products in the accumulate order of each call site (dot3_*),
//   * atan2f as glibc 2.39 computes it (fdlibm e_atan2f.c / s_atanf.c, no FMA) -- the function
//     behind torch.atan2 on two 0-dim CPU tensors (ref :48),
//   * cosf as MKL VML vmsCos(VML_HA) computes it (fp64 reduction + odd polynomial) -- the
//     function behind torch.cos on a CPU tensor (ref :48).
//
// Everything here is __host__ __device__ so that tests/ can sweep it on the CPU against libm
// and torch without a GPU.  The translation unit MUST be compiled with -fmad=false: every
// a*b+c below is two roundings unless written as fmaf()/fma().
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VIDC_HD __host__ __device__ __forceinline__
#else
#define VIDC_HD static inline
#endif

namespace vidc {

VIDC_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
VIDC_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// (a0 b0 + a1 b1) + a2 b2, every op rounded -- bmm / broadcast matmul call sites (ref :42,:43,:55)
VIDC_HD float dot3_muladd(float a0, float a1, float a2, float b0, float b1, float b2) {
    const float p0 = a0 * b0, p1 = a1 * b1, p2 = a2 * b2;
    const float s = p0 + p1;
    return s + p2;
}
// k-ascending FMA chain -- mm call sites (ref :54 S@S, :55 (KR)@Kinv, :142, :242)
VIDC_HD float dot3_fma(float a0, float a1, float a2, float b0, float b1, float b2) {
    float s = a0 * b0;
    s = fmaf(a1, b1, s);
    return fmaf(a2, b2, s);
}
// (a0 b0 + a2 b2) + a1 b1, no FMA -- mm with the column-major corners operand (ref :18,:125)
VIDC_HD float dot3_021(float a0, float a1, float a2, float b0, float b1, float b2) {
    const float p0 = a0 * b0, p1 = a1 * b1, p2 = a2 * b2;
    const float s = p0 + p2;
    return s + p1;
}

// ---- glibc 2.39 atanf (sysdeps/ieee754/flt-32/s_atanf.c), plain mul/add, no contraction ----
VIDC_HD float glibc_atanf(float x) {
    const float aT0 = u2f(0x3eaaaaabu), aT1 = u2f(0xbe4ccccdu), aT2 = u2f(0x3e124925u),
                aT3 = u2f(0xbde38e38u), aT4 = u2f(0x3dba2e6eu), aT5 = u2f(0xbd9d8795u),
                aT6 = u2f(0x3d886b35u), aT7 = u2f(0xbd6ef16bu), aT8 = u2f(0x3d4bda59u),
                aT9 = u2f(0xbd15a221u), aT10 = u2f(0x3c8569d7u);
    const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
    float hi = 0.0f, lo = 0.0f;
    int id;
    if (ix >= 0x4c000000u) {                       // |x| >= 2^25 (or NaN)
        if (ix > 0x7f800000u) return x + x;
        const float r = u2f(0x3fc90fdau) + u2f(0x33a22168u);
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) {                        // |x| < 0.4375
        if (ix < 0x31000000u) return x;            // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000u) {                    // |x| < 1.1875
            if (ix < 0x3f300000u) { id = 0; x = ((x + x) - 1.0f) / (x + 2.0f); hi = u2f(0x3eed6338u); lo = u2f(0x31ac3769u); }
            else                  { id = 1; x = (x - 1.0f) / (x + 1.0f);       hi = u2f(0x3f490fdau); lo = u2f(0x33222168u); }
        } else {
            if (ix < 0x401c0000u) { id = 2; x = (x - 1.5f) / (x * 1.5f + 1.0f); hi = u2f(0x3f7b985eu); lo = u2f(0x33140fb4u); }
            else                  { id = 3; x = -1.0f / x;                      hi = u2f(0x3fc90fdau); lo = u2f(0x33a22168u); }
        }
    }
    const float z = x * x;
    const float w = z * z;
    float s1 = aT10 * w + aT8;
    s1 = s1 * w + aT6;
    s1 = s1 * w + aT4;
    s1 = s1 * w + aT2;
    s1 = s1 * w + aT0;
    s1 = s1 * z;
    float s2 = aT9 * w + aT7;
    s2 = s2 * w + aT5;
    s2 = s2 * w + aT3;
    s2 = s2 * w + aT1;
    s2 = s2 * w;
    const float t = x * (s1 + s2);
    if (id < 0) return x - t;
    const float r = hi - ((t - lo) - x);
    return (hx >> 31) ? -r : r;
}

// ---- glibc 2.39 atan2f (sysdeps/ieee754/flt-32/e_atan2f.c) ----
VIDC_HD float glibc_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = u2f(0x3f490fdbu), pi_o_2 = u2f(0x3fc90fdbu), pi = u2f(0x40490fdbu),
                pi_lo = u2f(0xb3bbbd2eu);
    const uint32_t hx = f2u(x), hy = f2u(y);
    const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;          // NaN
    if (hx == 0x3f800000u) return glibc_atanf(y);                    // x == 1.0
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);   // 2*sign(x) + sign(y)
    if (iy == 0) {
        switch (m) {
            case 0: case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        }
        switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = glibc_atanf(fabsf(y / x));
    switch (m) {
        case 0: return z;
        case 1: return u2f(f2u(z) ^ 0x80000000u);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

// ---- MKL VML vmsCos, VML_HA accuracy, main path |x| <= 10000 (MKL 2024.2, AVX-512 kernel) ----
//   N = rint((|x| + pi/2)/pi) via the 1.5*2^23 shifter (fp32), r = |x| - (N - 1/2) pi in fp64,
//   cos x = (-1)^N (float)(r + r (r^2 P(r^2)))
VIDC_HD float mkl_cosf_ha(float x) {
    const float HALFPI = u2f(0x3fc90fdbu), INVPI = u2f(0x3ea2f983u), SHIFTER = u2f(0x4b400000u);
    const double PI_HI = 0x1.921fb5444p+1, PI_LO = 0x1.68c234c4c6629p-38;
    const double C3 = -0x1.55554bc836587p-3, C5 = 0x1.110ed3804ca96p-7,
                 C7 = -0x1.9f6ffeea73463p-13, C9 = 0x1.5dbdf0e4c7deep-19;
    const float ax = fabsf(x);
    const float t = ax + HALFPI;
    const float y = fmaf(t, INVPI, SHIFTER);
    float n = y - SHIFTER;
    n = n - 0.5f;
    const double dn = (double)n;
    double r = (double)ax;
    r = fma(-PI_HI, dn, r);
    r = fma(-dn, PI_LO, r);
    const double r2 = r * r;
    double p = fma(C9, r2, C7);
    p = fma(r2, p, C5);
    p = fma(r2, p, C3);
    const double q = p * r2;
    const float f = (float)fma(r, q, r);
    return u2f(f2u(f) ^ (f2u(y) << 31));
}


// ---- MKL VML vmsSin, VML_HA (same reduction / polynomial as the cosine; behind torch.sin on CPU tensors,
//      dataset.py:345,:483) ----
VIDC_HD float mkl_sinf_ha(float x) {
    const float INVPI = u2f(0x3ea2f983u), SHIFTER = u2f(0x4b400000u);
    const double PI_HI = 0x1.921fb5444p+1, PI_LO = 0x1.68c234c4c6629p-38;
    const double C3 = -0x1.55554bc836587p-3, C5 = 0x1.110ed3804ca96p-7,
                 C7 = -0x1.9f6ffeea73463p-13, C9 = 0x1.5dbdf0e4c7deep-19;
    const float ax = fabsf(x);
    const float y = fmaf(ax, INVPI, SHIFTER);
    const float n = y - SHIFTER;
    const double dn = (double)n;
    double r = (double)ax;
    r = fma(-PI_HI, dn, r);
    r = fma(-dn, PI_LO, r);
    const double r2 = r * r;
    double p = fma(C9, r2, C7);
    p = fma(r2, p, C5);
    p = fma(r2, p, C3);
    const double q = p * r2;
    const float f = (float)fma(r, q, r);
    return u2f(f2u(f) ^ (f2u(y) << 31) ^ (f2u(x) & 0x80000000u));
}

// ---- dataset-side gravity conditioning (dataset.py:45-55 rule 1, :334-345 / :472-483 rule 0) ----
VIDC_HD void condition_gravity(const float* raw, int rule, float* g, float* a) {
    float g0 = raw[0], g1 = raw[1], g2 = raw[2];
    if (rule == 0) { g1 = -g1; g2 = -g2; }                       // :473-474
    const float s1 = g1 * g1, s2 = g2 * g2;
    const float psi = s1 + s2;                                   // :475
    a[0] = 0.0f; a[1] = 1.0f; a[2] = 0.0f;
    if (rule == 0) {
        if (!(psi < 1e-4f)) {                                    // :476
            const float pitch = glibc_atan2f(g2, g1);            // :479
            const float c = mkl_cosf_ha(pitch);
            if (!(c > 0.707f)) { a[1] = c; a[2] = mkl_sinf_ha(pitch); }   // :480-483
        }
    } else {
        bool keep_g = psi < 1e-6f;                               // :47-48
        if (!keep_g) keep_g = !(mkl_cosf_ha(glibc_atan2f(g2, g1)) > 0.3f);   // :50-54
        if (keep_g) { a[0] = g0; a[1] = g1; a[2] = g2; }
    }
    g[0] = g0; g[1] = g1; g[2] = g2;
}

}  // namespace vidc
