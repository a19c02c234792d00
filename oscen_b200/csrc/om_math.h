/*
 * om_math.h — libm-faithful f32 transcendentals for the device kernels.
 *
 * Why: the reference (Rust f32::sin/cos/exp/tan/tanh) lowers to the platform libm, and the
 * parity oracle links glibc.  Three of these feed recurrences that accumulate error over
 * thousands of samples (ADSR one-pole coefficient 1-exp(-K/n); the electric-piano rotation
 * multipliers cos/sin(2*pi*f/sr); FM operator phase chains), so "within 1 ulp of libm" is not
 * enough there.  om_expf / om_sinf / om_cosf therefore restate the published algorithms glibc
 * >= 2.28 ships for these functions (ARM optimized-routines: table-driven exp2f core and the
 * double-precision sincosf polynomials) operation by operation in IEEE double, which a GPU
 * evaluates bit-identically.  tests/test_om_math.py checks them against the box's glibc
 * (bit-exact on dense sweeps); DESIGN.md records the exhaustive sweep.
 *
 * om_tanhf / om_tanf restate glibc 2.39's fdlibm-derived pure-f32 routines (tanhf+expm1f are
 * bit-exact against the box's libm; tanf is bit-exact except ~2e-6 of inputs next to pi/2 that
 * differ by 1 ulp, and it only feeds non-accumulating filter coefficients).
 *
 * The header compiles as plain C++ (host, for the CPU check) and as CUDA device code.
 * Nothing here calls CUDA's approximate intrinsics; kernels are built with -fmad=false and
 * every fused operation below is an explicit fma().
 */
#ifndef OSCEN_B200_OM_MATH_H
#define OSCEN_B200_OM_MATH_H

#if defined(__CUDACC_RTC__)
#include "rtc_prelude.h"
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#endif

#if defined(__CUDACC__)
#define OM_HD __host__ __device__ __forceinline__
#else
#define OM_HD static inline
#endif

/* glibc's x86-64 FMA ifunc variants (__expf_fma, __sinf_fma, ...) are the ones selected on
   every CPU with FMA3; gcc contracts each a*b+c of the C source into one fma there. */
#ifndef OM_GLIBC_FMA
#define OM_GLIBC_FMA 1
#endif
#if OM_GLIBC_FMA
#define OM_MADD(a, b, c) fma((a), (b), (c))
#else
#define OM_MADD(a, b, c) ((a) * (b) + (c))
#endif

OM_HD uint32_t om_asuint(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
OM_HD uint64_t om_asuint64(double f) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(f);
#else
    uint64_t u; memcpy(&u, &f, 8); return u;
#endif
}
OM_HD double om_asdouble(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double f; memcpy(&f, &u, 8); return f;
#endif
}

/* 2^(i/32) table: tab[i] = bits(2^(i/32)) - (i << 47)  (exp2f_data.c layout) */
#define OM_EXP2F_TAB                                                                              \
    { 0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
      0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
      0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
      0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
      0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
      0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
      0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
      0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull }

#if defined(__CUDACC__)
static __device__ __constant__ uint64_t om_exp2f_tab_dev[32] = OM_EXP2F_TAB;
#endif
static const uint64_t om_exp2f_tab_host[32] = OM_EXP2F_TAB;

OM_HD uint64_t om_exp2f_tab(uint32_t i) {
#if defined(__CUDA_ARCH__)
    return om_exp2f_tab_dev[i];
#else
    return om_exp2f_tab_host[i];
#endif
}

/* core: exp(x) for |x| < ~88 evaluated in double, returned as double (error ~2^-34 rel.) */
OM_HD double om_exp_core(double xd) {
    const double N = 32.0;
    const double InvLn2N = 0x1.71547652b82fep+0 * N;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / N / N / N;
    const double C1 = 0x1.ebfce50fac4f3p-3 / N / N;
    const double C2 = 0x1.62e42ff0c52d6p-1 / N;
#if OM_GLIBC_FMA
    /* gcc contracts both uses of z = InvLn2N*xd (z+SHIFT and z-kd) in __expf_fma; verified
       exhaustively (0 mismatches over every f32 with |x| < 88) against glibc 2.39 */
    double kd = fma(InvLn2N, xd, SHIFT);
    uint64_t ki = om_asuint64(kd);
    kd -= SHIFT;
    double r = fma(InvLn2N, xd, -kd);
    double z;
#else
    double z = InvLn2N * xd;
    double kd = z + SHIFT;
    uint64_t ki = om_asuint64(kd);
    kd -= SHIFT;
    double r = z - kd;
#endif
    uint64_t t = om_exp2f_tab((uint32_t)(ki & 31u));
    t += ki << (52 - 5);
    double s = om_asdouble(t);
    z = OM_MADD(C0, r, C1);
    double r2 = r * r;
    double y = OM_MADD(C2, r, 1.0);
    y = OM_MADD(z, r2, y);
    return y * s;
}

/* expf: glibc sysdeps/ieee754/flt-32/e_expf.c */
OM_HD float om_expf(float x) {
    uint32_t abstop = (om_asuint(x) >> 20) & 0x7ff;
    if (abstop >= 0x42b) {                       /* |x| >= 88 or NaN */
        if (om_asuint(x) == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8) return x + x;
        if (x > 0x1.62e42ep6f) return INFINITY;
        if (x < -0x1.9fe368p6f) return 0.0f;
    }
    return (float)om_exp_core((double)x);
}

/* sincosf polynomial data: glibc sysdeps/ieee754/flt-32/s_sincosf_data.c */
#define OM_HPI_INV 0x1.45F306DC9C883p+23        /* 2/pi * 2^24 */
#define OM_HPI 0x1.921FB54442D18p0
#define OM_C0 0x1p0
#define OM_C1 (-0x1.ffffffd0c621cp-2)
#define OM_C2 0x1.55553e1068f19p-5
#define OM_C3 (-0x1.6c087e89a359dp-10)
#define OM_C4 0x1.99343027bf8c3p-16
#define OM_S1 (-0x1.555545995a603p-3)
#define OM_S2 0x1.1107605230bc4p-7
#define OM_S3 (-0x1.994eb3774cf24p-13)

/* sinf_poly (sincosf.h): n even -> sine polynomial, n odd -> cosine polynomial; csign = +-1
   selects __sincosf_table[0] / [1] (the cosine coefficients negated). */
OM_HD float om_sinf_poly(double x, double x2, double csign, int n) {
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = OM_MADD(x2, OM_S3, OM_S2);
        double x7 = x3 * x2;
        double s = OM_MADD(x3, OM_S1, x);
        return (float)OM_MADD(x7, s1, s);
    } else {
        double x4 = x2 * x2;
        double c2 = OM_MADD(x2, csign * OM_C4, csign * OM_C3);
        double c1 = OM_MADD(x2, csign * OM_C1, csign * OM_C0);
        double x6 = x4 * x2;
        double c = OM_MADD(x4, csign * OM_C2, c1);
        return (float)OM_MADD(x6, c2, c);
    }
}

/* 4/pi in 24-bit pieces: __inv_pio4 of s_sincosf_data.c, for |x| >= 120 (reduce_large) */
#define OM_INV_PIO4                                                                          \
    { 0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u,         \
      0x4e441529u, 0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u,          \
      0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u, 0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u,          \
      0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u }
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t om_inv_pio4_dev[24] = OM_INV_PIO4;
#endif
static const uint32_t om_inv_pio4_host[24] = OM_INV_PIO4;

/* reduce_large (sincosf.h): Payne-Hanek style reduction with 3 x 32-bit multiplies */
OM_HD double om_reduce_large(uint32_t xi, int *np) {
#if defined(__CUDA_ARCH__)
    const uint32_t *arr = &om_inv_pio4_dev[(xi >> 26) & 15];
#else
    const uint32_t *arr = &om_inv_pio4_host[(xi >> 26) & 15];
#endif
    int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = (uint64_t)(xi * arr[0]);            /* 32-bit wrapping product, as in the source */
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return x * 0x1.921FB54442D18p-62;          /* pi / 2^64 */
}

/* sinf: glibc sysdeps/ieee754/flt-32/s_sinf.c */
OM_HD float om_sinf(float y) {
    double x = (double)y;
    uint32_t top = (om_asuint(y) >> 20) & 0x7ff;
    if (top < 0x3f4) {                           /* abstop12(y) < abstop12(pio4) */
        double s = x * x;
        if (top < 0x398) return y;               /* |y| < 2^-12 */
        return om_sinf_poly(x, s, 1.0, 0);
    } else if (top < 0x42f) {                    /* |y| < 120 */
        double r = x * OM_HPI_INV;
        int n = ((int32_t)r + 0x800000) >> 24;
        x = OM_MADD(-(double)n, OM_HPI, x);
        /* sign[n & 3] = {1,-1,-1,1} in both tables; table[1] (n & 2) negates the cosine coeffs */
        double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        return om_sinf_poly(x * sg, x * x, (n & 2) ? -1.0 : 1.0, n);
    } else if (top < 0x7f8) {
        uint32_t xi = om_asuint(y);
        int sign = (int)(xi >> 31);
        int n;
        x = om_reduce_large(xi, &n);
        int q = (n + sign) & 3;
        double sg = (q == 1 || q == 2) ? -1.0 : 1.0;
        return om_sinf_poly(x * sg, x * x, ((n + sign) & 2) ? -1.0 : 1.0, n);
    }
    return y - y;                                /* inf/nan -> nan */
}

/* cosf: glibc sysdeps/ieee754/flt-32/s_cosf.c */
OM_HD float om_cosf(float y) {
    double x = (double)y;
    uint32_t top = (om_asuint(y) >> 20) & 0x7ff;
    if (top < 0x3f4) {
        double x2 = x * x;
        if (top < 0x398) return 1.0f;
        return om_sinf_poly(x, x2, 1.0, 1);
    } else if (top < 0x42f) {
        double r = x * OM_HPI_INV;
        int n = ((int32_t)r + 0x800000) >> 24;
        x = OM_MADD(-(double)n, OM_HPI, x);
        double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        return om_sinf_poly(x * sg, x * x, (n & 2) ? -1.0 : 1.0, n ^ 1);
    } else if (top < 0x7f8) {
        uint32_t xi = om_asuint(y);
        int sign = (int)(xi >> 31);
        int n;
        x = om_reduce_large(xi, &n);
        int q = (n + sign) & 3;
        double sg = (q == 1 || q == 2) ? -1.0 : 1.0;
        return om_sinf_poly(x * sg, x * x, ((n + sign) & 2) ? -1.0 : 1.0, n ^ 1);
    }
    return y - y;
}

/* ---- branch-free forms for the kernels' inner loops --------------------------------------
 * Same arithmetic as om_sinf / om_cosf (so the same bits as glibc), reorganised so that a warp
 * whose lanes land in different quadrants does not execute three divergent paths:
 *  - |y| < 0.75 takes reduce_fast too: there n = 0 and fma(-0.0, hpi, x) = x exactly, and for
 *    |y| < 2^-12 the polynomial rounds back to y (x^3/6 is < 1/6 ulp);
 *  - table[1] (quadrants 2,3) negates every cosine coefficient and sign[] negates x for the
 *    sine polynomial; both polynomials are odd-symmetric under that negation in round-to-
 *    nearest, so the result is the table[0] polynomial negated iff (n & 2);
 *  - the sine and cosine polynomials share one 2-DMUL + 4-DFMA schedule with per-lane selected
 *    constants (lane-select instead of branch).
 * tests/csrc/check_om_math.cpp sweeps them against glibc (bit-exact). */
OM_HD double om_sel(bool c, double a, double b) { return c ? a : b; }
/* sinf for |y| < 120 (top12 < 0x42f); else the general routine */
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ float om_sinf_slow(float y) { return om_sinf(y); }   /* |y| >= 120: cold */
static __device__ __noinline__ float om_cosf_slow(float y) { return om_cosf(y); }
#else
static inline float om_sinf_slow(float y) { return om_sinf(y); }
static inline float om_cosf_slow(float y) { return om_cosf(y); }
#endif
OM_HD float om_sinf_u(float y) {
    const uint32_t top = (om_asuint(y) >> 20) & 0x7ff;
    if (top >= 0x42f) return om_sinf_slow(y);
    double x = (double)y;
    double r = x * OM_HPI_INV;
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = OM_MADD(-(double)n, OM_HPI, x);
    const bool odd = (n & 1) != 0;
    const double x2 = x * x;
    const double a = om_sel(odd, x2, x) * x2;
    const double p = OM_MADD(x2, om_sel(odd, OM_C4, OM_S3), om_sel(odd, OM_C3, OM_S2));
    const double b = a * x2;
    const double c1 = OM_MADD(x2, OM_C1, OM_C0);
    const double q = om_sel(odd, c1, x);
    const double sc = OM_MADD(a, om_sel(odd, OM_C2, OM_S1), q);
    const float res = (float)OM_MADD(b, p, sc);
    return (n & 2) ? -res : res;
}
OM_HD float om_cosf_u(float y) {
    const uint32_t top = (om_asuint(y) >> 20) & 0x7ff;
    if (top >= 0x42f) return om_cosf_slow(y);
    if (top < 0x398) return 1.0f;
    double x = (double)y;
    double r = x * OM_HPI_INV;
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = OM_MADD(-(double)n, OM_HPI, x);
    const bool odd = ((n ^ 1) & 1) != 0;          /* sinf_poly(.., n ^ 1) */
    const double x2 = x * x;
    const double a = om_sel(odd, x2, x) * x2;
    const double p = OM_MADD(x2, om_sel(odd, OM_C4, OM_S3), om_sel(odd, OM_C3, OM_S2));
    const double b = a * x2;
    const double c1 = OM_MADD(x2, OM_C1, OM_C0);
    const double q = om_sel(odd, c1, x);
    const double sc = OM_MADD(a, om_sel(odd, OM_C2, OM_S1), q);
    const float res = (float)OM_MADD(b, p, sc);
    /* even quadrant (cos poly): table sign from (n & 2); odd quadrant (sine poly on x*sign[n&3]):
       sign[] = {1,-1,-1,1} */
    const bool neg = odd ? ((n & 2) != 0) : (((n & 3) == 1) || ((n & 3) == 2));
    return neg ? -res : res;
}

/* sinf for callers that GUARANTEE |y| < 120 (top12 < 0x42f) — no range test, no branch: the reduction of om_sinf's middle
   path, then BOTH of its polynomials evaluated in IEEE double (independent chains: 2 DMUL + 3 DFMA and 2 DMUL + 4 DFMA) and
   the quadrant's one selected.  No per-lane selection of 64-bit coefficients (10 FSEL and the constant moves that go with
   them in om_sinf_u).  Same operations on the selected path as om_sinf, so the same bits as glibc;
   tests/csrc/check_om_math.cpp sweeps it against the box's libm (0 mismatches over every f32 with |y| < 120).
   Where the constants live matters in a frame loop that evaluates it three times: the six MULTIPLIER constants sit in the
   constant bank (a DFMA takes one operand from there for free; as immediates they cost two moves each per use), the three
   ADDEND constants come in through `k`, which a kernel fills once and pins in registers (om_sin_regs_pinned). */
struct OmSinRegs { double s2, c3, c1; };
#if defined(__CUDACC__)
static __device__ __constant__ double om_sin_mul_dev[6] = { OM_HPI_INV, OM_HPI, OM_S3, OM_S1, OM_C4, OM_C2 };
#endif
#if defined(__CUDA_ARCH__)
#define OM_SIN_MUL(i) om_sin_mul_dev[i]
#else
static const double om_sin_mul_host[6] = { OM_HPI_INV, OM_HPI, OM_S3, OM_S1, OM_C4, OM_C2 };
#define OM_SIN_MUL(i) om_sin_mul_host[i]
#endif
#if defined(__CUDACC__)
static __device__ __constant__ double om_sin_add_dev[3] = { OM_S2, OM_C3, OM_C1 };
#endif
OM_HD OmSinRegs om_sin_regs() {
    OmSinRegs k;
#if defined(__CUDA_ARCH__)
    k.s2 = om_sin_add_dev[0]; k.c3 = om_sin_add_dev[1]; k.c1 = om_sin_add_dev[2];    /* loaded, not immediates: ptxas keeps a load's result live */
#else
    k.s2 = OM_S2; k.c3 = OM_C3; k.c1 = OM_C1;
#endif
    return k;
}
OM_HD float om_sinf_rk(float y, const OmSinRegs &k) {
    double x = (double)y;
    const double r = x * OM_SIN_MUL(0);
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = OM_MADD(-(double)n, OM_SIN_MUL(1), x);
    const double x2 = x * x;
    /* sine polynomial (n even) */
    const double x3 = x * x2;
    const double s1 = OM_MADD(x2, OM_SIN_MUL(2), k.s2);
    const double x5 = x3 * x2;
    const double s = OM_MADD(x3, OM_SIN_MUL(3), x);
    const double rs = OM_MADD(x5, s1, s);
    /* cosine polynomial (n odd) */
    const double x4 = x2 * x2;
    const double c2 = OM_MADD(x2, OM_SIN_MUL(4), k.c3);
    const double c1 = OM_MADD(x2, k.c1, OM_C0);
    const double x6 = x4 * x2;
    const double c = OM_MADD(x4, OM_SIN_MUL(5), c1);
    const double rc = OM_MADD(x6, c2, c);
    const float res = (float)((n & 1) ? rc : rs);
    return (n & 2) ? -res : res;
}
OM_HD float om_sinf_r(float y) { return om_sinf_rk(y, om_sin_regs()); }
/* three independent om_sinf_rk evaluations written step by step across the three arguments: the compiler keeps the source
   order of independent instructions, so the three conversion / FP64 chains (each ~200 cycles deep) overlap instead of
   queueing behind one another.  Same operations per argument as om_sinf_rk. */
OM_HD void om_sinf_rk3(float y0, float y1, float y2, const OmSinRegs &k, float &o0, float &o1, float &o2) {
    double x[3] = { (double)y0, (double)y1, (double)y2 };
    int n[3];
    double x2[3], rs[3], rc[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; ++i) { const double r = x[i] * OM_SIN_MUL(0); n[i] = ((int32_t)r + 0x800000) >> 24; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; ++i) { x[i] = OM_MADD(-(double)n[i], OM_SIN_MUL(1), x[i]); x2[i] = x[i] * x[i]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; ++i) {
        const double x3 = x[i] * x2[i];
        const double s1 = OM_MADD(x2[i], OM_SIN_MUL(2), k.s2);
        const double x5 = x3 * x2[i];
        const double s = OM_MADD(x3, OM_SIN_MUL(3), x[i]);
        rs[i] = OM_MADD(x5, s1, s);
        const double x4 = x2[i] * x2[i];
        const double c2 = OM_MADD(x2[i], OM_SIN_MUL(4), k.c3);
        const double c1 = OM_MADD(x2[i], k.c1, OM_C0);
        const double x6 = x4 * x2[i];
        const double c = OM_MADD(x4, OM_SIN_MUL(5), c1);
        rc[i] = OM_MADD(x6, c2, c);
    }
    float res[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; ++i) { const float t = (float)((n[i] & 1) ? rc[i] : rs[i]); res[i] = (n[i] & 2) ? -t : t; }
    o0 = res[0]; o1 = res[1]; o2 = res[2];
}

OM_HD float om_asfloat(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

/* ---- tanf: glibc 2.39 sysdeps/ieee754/flt-32/{s_tanf,k_tanf,e_rem_pio2f}.c (fdlibm, pure f32,
   built without contraction).  Restated for 0 <= |x| < 3*pi/4, which covers every filter
   argument on the path (TptFilter: pi*fc/sr < pi/2; LP18Filter: pi*[0.001,0.33]). */
OM_HD float om_k_tanf(float x, float y, int iy) {
    const float T0 = 3.3333334327e-01f, T1 = 1.3333334029e-01f, T2 = 5.3968254477e-02f,
                T3 = 2.1869488060e-02f, T4 = 8.8632395491e-03f, T5 = 3.5920790397e-03f,
                T6 = 1.4562094584e-03f, T7 = 5.8804126456e-04f, T8 = 2.4646313977e-04f,
                T9 = 7.8179444245e-05f, T10 = 7.1407252108e-05f, T11 = -1.8558637748e-05f,
                T12 = 2.5907305826e-05f;
    const float pio4 = 7.8539812565e-01f, pio4lo = 3.7748947079e-08f;
    float z, r, v, w, s;
    int32_t hx = (int32_t)om_asuint(x);
    int32_t ix = hx & 0x7fffffff;
    if (ix < 0x39000000) {                              /* |x| < 2^-13 */
        if ((int)x == 0) {
            if ((ix | (iy + 1)) == 0) return 1.0f / fabsf(x);
            else if (iy == 1) return x;
            else return -1.0f / x;
        }
    }
    if (ix >= 0x3f2ca140) {                             /* |x| >= 0.6744 */
        if (hx < 0) { x = -x; y = -y; }
        z = pio4 - x;
        w = pio4lo - y;
        x = z + w; y = 0.0f;
        if (fabsf(x) < 0x1p-13f) return (float)((1 - ((hx >> 30) & 2)) * iy) * (1.0f - (float)(2 * iy) * x);
    }
    z = x * x;
    w = z * z;
    r = T1 + w * (T3 + w * (T5 + w * (T7 + w * (T9 + w * T11))));
    v = z * (T2 + w * (T4 + w * (T6 + w * (T8 + w * (T10 + w * T12)))));
    s = z * x;
    r = y + z * (s * (r + v) + y);
    r += T0 * s;
    w = x + r;
    if (ix >= 0x3f2ca140) {
        v = (float)iy;
        return (float)(1 - ((hx >> 30) & 2)) * (v - 2.0f * (x - (w * w / (w + v) - r)));
    }
    if (iy == 1) return w;
    float a, t;
    z = om_asfloat(om_asuint(w) & 0xfffff000u);
    v = r - (z - x);
    t = a = -1.0f / w;
    t = om_asfloat(om_asuint(t) & 0xfffff000u);
    s = 1.0f + t * z;
    return t + a * (s + t * v);
}

/* om_k_tanf(x, 0, 1) for 2^-13 <= x < 0.6744 (0x39000000 <= bits < 0x3f2ca140, x > 0): the polynomial core without the
   range branches — every TptFilter coefficient refresh below ~10.3 kHz at 48 kHz lands here.  Same operations in the
   same order (y = 0 terms kept: `0 + z*(..+0)` is exact for these positive operands), so the same bits. */
OM_HD bool om_tanf_mid_ok(float x) {
    const uint32_t ix = om_asuint(x);
    return ix >= 0x39000000u && ix < 0x3f2ca140u;
}
OM_HD float om_k_tanf_mid(float x) {
    const float T0 = 3.3333334327e-01f, T1 = 1.3333334029e-01f, T2 = 5.3968254477e-02f,
                T3 = 2.1869488060e-02f, T4 = 8.8632395491e-03f, T5 = 3.5920790397e-03f,
                T6 = 1.4562094584e-03f, T7 = 5.8804126456e-04f, T8 = 2.4646313977e-04f,
                T9 = 7.8179444245e-05f, T10 = 7.1407252108e-05f, T11 = -1.8558637748e-05f,
                T12 = 2.5907305826e-05f;
    const float z = x * x;
    const float w = z * z;
    float r = T1 + w * (T3 + w * (T5 + w * (T7 + w * (T9 + w * T11))));
    const float v = z * (T2 + w * (T4 + w * (T6 + w * (T8 + w * (T10 + w * T12)))));
    const float s = z * x;
    r = 0.0f + z * (s * (r + v) + 0.0f);
    r += T0 * s;
    return x + r;
}

/* tan in double, rounded once (fallback for |x| >= 3*pi/4; within 1 ulp of a faithful libm) */
OM_HD float om_tanf_wide(float y) {
    double x = (double)y;
    if (!(fabs(x) < 1.0e4)) return tanf(y);
    double q = nearbyint(x * 0x1.45F306DC9C883p-1);           /* x * 2/pi */
    int n = (int)q;
    double r = fma(-q, 0x1.921FB54442D18p0, x);
    r = fma(-q, 0x1.1A62633145C07p-54, r);
    double r2 = r * r;
    double s = -1.0 / 1307674368000.0;
    s = fma(s, r2, 1.0 / 6227020800.0);
    s = fma(s, r2, -1.0 / 39916800.0);
    s = fma(s, r2, 1.0 / 362880.0);
    s = fma(s, r2, -1.0 / 5040.0);
    s = fma(s, r2, 1.0 / 120.0);
    s = fma(s, r2, -1.0 / 6.0);
    s = fma(s * r2, r, r);
    double c = 1.0 / 20922789888000.0;
    c = fma(c, r2, -1.0 / 87178291200.0);
    c = fma(c, r2, 1.0 / 479001600.0);
    c = fma(c, r2, -1.0 / 3628800.0);
    c = fma(c, r2, 1.0 / 40320.0);
    c = fma(c, r2, -1.0 / 720.0);
    c = fma(c, r2, 1.0 / 24.0);
    c = fma(c, r2, -0.5);
    c = fma(c, r2, 1.0);
    return (float)((n & 1) ? -c / s : s / c);
}

OM_HD float om_tanf(float x) {
    uint32_t hx = om_asuint(x);
    uint32_t ix = hx & 0x7fffffffu;
    if (ix <= 0x3f490fdau) return om_k_tanf(x, 0.0f, 1);        /* |x| <= pi/4 */
    if (ix < 0x4016cbe4u) {                                      /* |x| < 3*pi/4: n = +-1 */
        const float pio2_1 = 1.5707855225e+00f, pio2_1t = 1.0804334124e-05f,
                    pio2_2 = 1.0804273188e-05f, pio2_2t = 6.0770999344e-11f;
        float z, y0, y1;
        if ((hx >> 31) == 0) {
            z = x - pio2_1;
            if ((ix & 0xfffffff0u) != 0x3fc90fd0u) { y0 = z - pio2_1t; y1 = (z - y0) - pio2_1t; }
            else { z -= pio2_2; y0 = z - pio2_2t; y1 = (z - y0) - pio2_2t; }
        } else {
            z = x + pio2_1;
            if ((ix & 0xfffffff0u) != 0x3fc90fd0u) { y0 = z + pio2_1t; y1 = (z - y0) + pio2_1t; }
            else { z += pio2_2; y0 = z + pio2_2t; y1 = (z - y0) + pio2_2t; }
        }
        return om_k_tanf(y0, y1, -1);
    }
    return om_tanf_wide(x);
}

/* ---- tanhf: glibc 2.39 sysdeps/ieee754/flt-32/{s_tanhf,s_expm1f}.c (fdlibm, pure f32) */
OM_HD float om_expm1f(float x) {
    const float one = 1.0f, ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f,
                invln2 = 1.4426950216e+00f, Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f,
                Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
    float y, hi, lo, c = 0.0f, t, e, hxs, hfx, r1;
    int32_t k;
    uint32_t hx = om_asuint(x);
    uint32_t xsb = hx & 0x80000000u;
    hx &= 0x7fffffffu;
    if (hx >= 0x4195b844u) {                            /* |x| >= 27 ln2 */
        if (hx >= 0x42b17218u) {
            if (hx > 0x7f800000u) return x + x;
            if (hx == 0x7f800000u) return (xsb == 0) ? x : -1.0f;
            if (x > 8.8721679688e+01f) return INFINITY;
        }
        if (xsb != 0) return 1.0e-30f - one;
    }
    if (hx > 0x3eb17218u) {                             /* |x| > 0.5 ln2 */
        if (hx < 0x3F851592u) {                         /* |x| < 1.5 ln2 */
            if (xsb == 0) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
            else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
        } else {
            k = (int32_t)(invln2 * x + ((xsb == 0) ? 0.5f : -0.5f));
            t = (float)k;
            hi = x - t * ln2_hi;
            lo = t * ln2_lo;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (hx < 0x33000000u) {
        return x;
    } else k = 0;
    hfx = 0.5f * x;
    hxs = x * hfx;
    r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    t = 3.0f - r1 * hfx;
    e = hxs * ((r1 - t) / (6.0f - x * t));
    if (k == 0) return x - (x * e - hxs);
    e = (x * (e - c) - c);
    e -= hxs;
    if (k == -1) return 0.5f * (x - e) - 0.5f;
    if (k == 1) {
        if (x < -0.25f) return -2.0f * (e - (x + 0.5f));
        return one + 2.0f * (x - e);
    }
    if (k <= -2 || k > 56) {
        y = one - (e - x);
        y = om_asfloat(om_asuint(y) + ((uint32_t)k << 23));
        return y - one;
    }
    if (k < 23) {
        t = om_asfloat(0x3f800000u - (0x1000000u >> k));
        y = t - (e - x);
        y = om_asfloat(om_asuint(y) + ((uint32_t)k << 23));
    } else {
        t = om_asfloat((uint32_t)(0x7f - k) << 23);
        y = x - (e + t);
        y += one;
        y = om_asfloat(om_asuint(y) + ((uint32_t)k << 23));
    }
    return y;
}

OM_HD float om_tanhf(float x) {
    const float one = 1.0f, two = 2.0f;
    float t, z;
    uint32_t jx = om_asuint(x), ix = jx & 0x7fffffffu;
    if (ix >= 0x7f800000u) return (jx >> 31) ? one / x - one : one / x + one;
    if (ix < 0x41b00000u) {                             /* |x| < 22 */
        if (ix == 0) return x;
        if (ix < 0x24000000u) return x * (one + x);
        if (ix >= 0x3f800000u) {
            t = om_expm1f(two * fabsf(x));
            z = one - two / (t + two);
        } else {
            t = om_expm1f(-two * fabsf(x));
            z = -t / (t + two);
        }
    } else z = one - 1.0e-30f;
    return (jx >> 31) ? -z : z;
}

/* ---- tanhf, lane-select form for the kernels (same bits as om_tanhf / glibc) ---------------
 * expm1f's argument reduction is one formula for every k (k = 0 and k = +-1 are the general
 * `hi = x - k*ln2_hi, lo = k*ln2_lo` with exact products), so only k itself and the final
 * reconstruction are selected per lane; a warp whose lanes fall in different |x| bands no longer
 * serialises three code paths.  Domain handled here: 2^-55 <= |x| < 22 with the expm1f argument
 * in (-2, 0] (|x| < 1) or [2, 44) (|x| >= 1); everything else defers to om_tanhf. */
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ float om_tanhf_slow(float x) { return om_tanhf(x); }   /* cold ranges */
#else
static inline float om_tanhf_slow(float x) { return om_tanhf(x); }
#endif
OM_HD float om_tanhf_u(float x) {
    const uint32_t jx = om_asuint(x), ix = jx & 0x7fffffffu;
    if (ix >= 0x41b00000u || ix < 0x24000000u) return om_tanhf_slow(x);     /* |x| >= 22, tiny, 0, inf, nan */
    const float one = 1.0f, ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f,
                invln2 = 1.4426950216e+00f, Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f,
                Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
    const bool big = ix >= 0x3f800000u;                                /* |x| >= 1 */
    const float ax = fabsf(x);
    float a = big ? 2.0f * ax : -2.0f * ax;                            /* expm1f argument */
    const uint32_t ha = om_asuint(a) & 0x7fffffffu;
    if (ha >= 0x4195b844u) return om_tanhf_slow(x);                    /* |a| >= 27 ln2: cold */
    /* k */
    int32_t k;
    if (ha > 0x3eb17218u) {
        const int32_t kg = (int32_t)(invln2 * a + (big ? 0.5f : -0.5f));
        k = (ha < 0x3F851592u) ? (big ? 1 : -1) : kg;
    } else k = 0;
    const bool tiny = ha < 0x33000000u;                                /* |a| < 2^-25: expm1f(a) = a */
    const float t0 = (float)k;
    const float hi = a - t0 * ln2_hi;
    const float lo = t0 * ln2_lo;
    const float xr = hi - lo;
    const float c = (hi - xr) - lo;
    const float hfx = 0.5f * xr;
    const float hxs = xr * hfx;
    const float r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    const float tt = 3.0f - r1 * hfx;
    float e = hxs * ((r1 - tt) / (6.0f - xr * tt));
    float em1;
    if (k == 0) {
        em1 = xr - (xr * e - hxs);
    } else {
        e = (xr * (e - c) - c);
        e -= hxs;
        if (k == -1) em1 = 0.5f * (xr - e) - 0.5f;
        else if (k == 1) em1 = (xr < -0.25f) ? -2.0f * (e - (xr + 0.5f)) : one + 2.0f * (xr - e);
        else if (k <= -2 || k > 56) {
            float y = one - (e - xr);
            y = om_asfloat(om_asuint(y) + ((uint32_t)k << 23));
            em1 = y - one;
        } else if (k < 23) {
            const float t1 = om_asfloat(0x3f800000u - (0x1000000u >> k));
            float y = t1 - (e - xr);
            em1 = om_asfloat(om_asuint(y) + ((uint32_t)k << 23));
        } else {
            const float t1 = om_asfloat((uint32_t)(0x7f - k) << 23);
            float y = xr - (e + t1);
            y += one;
            em1 = om_asfloat(om_asuint(y) + ((uint32_t)k << 23));
        }
    }
    if (tiny) em1 = a;
    const float z = big ? one - 2.0f / (em1 + 2.0f) : -em1 / (em1 + 2.0f);
    return (jx >> 31) ? -z : z;
}

#endif /* OSCEN_B200_OM_MATH_H */
