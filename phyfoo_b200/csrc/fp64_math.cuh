// fp64_math.cuh -- exp and log for the mean-field kernels: table-driven, ~1 ulp, half the instructions of the CUDA
// math library's versions (which spend ~35 and ~55 instructions per call; the kernels take 4 exp and 1 log per hidden
// node and sweep, which is most of their arithmetic).  __host__ __device__ so that tests/emul runs the same bits on the
// CPU and tests/test_fp64_math.py can bound the error against the C library.
//
//   phy_exp(x):  x = (64 k + j) ln2/64 + r, |r| <= ln2/128;  exp(x) = 2^k * T[j] * (1 + r + r^2/2 + ... + r^6/720)
//   phy_log(z):  z = 2^k m, m in [0x1.68p-1, 0x1.68p0);  r = fma(m, invc[i], -1) exactly rounded, |r| < 2^-8;
//                log z = k ln2 + logc[i] + log1p(r) (degree 6)            [table layout: scripts/gen_math_tables.py]
// Special cases follow IEEE / java.lang.Math: exp(-inf) = 0, exp(x < -745.2) = 0 through denormals, exp(NaN) = NaN,
// log(0) = -inf, log(z < 0) = NaN, log(inf) = inf, denormal arguments are handled.
#pragma once
#if defined(__CUDACC_RTC__)
// run-time compilation (jit_o0.hpp): no host headers; the fixed-width types and the two macros the device paths use
typedef signed char int8_t; typedef unsigned char uint8_t; typedef int int32_t; typedef unsigned int uint32_t;
typedef long long int64_t; typedef unsigned long long uint64_t;
#define INFINITY __longlong_as_double(0x7FF0000000000000LL)
#define NAN __longlong_as_double(0x7FF8000000000000LL)
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#endif

#include "fp64_math_tables.h"

#if defined(__CUDACC__)
#define PHY_MATH_HD __host__ __device__ __forceinline__
#else
#define PHY_MATH_HD inline
#endif

namespace phyfoo {

#if defined(__CUDACC__)
static __device__ const double d_phy_exp_tab[64] = {PHY_EXP_TABLE_VALUES};
static __device__ const double d_phy_log_tab[256] = {PHY_LOG_TABLE_VALUES};
#endif
static const double h_phy_exp_tab[64] = {PHY_EXP_TABLE_VALUES};
static const double h_phy_log_tab[256] = {PHY_LOG_TABLE_VALUES};

PHY_MATH_HD uint64_t phy_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
PHY_MATH_HD double phy_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

// exp with the 64-entry table 2^(j/64) passed in (a kernel's copy in shared memory); phy_exp below uses the global one
PHY_MATH_HD double phy_exp_tab(double x, const double* tab64) {
    if (!(fabs(x) < 690.0)) {                                   // rare: huge, tiny, inf, NaN
        if (x != x) return x;
        if (x > 709.782712893384) return INFINITY;
        if (x < -745.1332191019412) return 0.0;
    }
    const double t = x * PHY_64_OVER_LN2;
    double kd = t + 0x1.8p52;                                   // round to nearest integer in the low mantissa bits
#if defined(__CUDA_ARCH__)
    const int32_t ki = __double2loint(kd);
#else
    const int32_t ki = (int32_t)(uint32_t)phy_bits(kd);
#endif
    kd -= 0x1.8p52;
    double r = fma(kd, -PHY_LN2_64_HI, x);
    r = fma(kd, -PHY_LN2_64_LO, r);
    const double T = tab64[ki & 63];
    double q = fma(r, 0x1.6c16c16c16c17p-10, 0x1.1111111111111p-7);   // 1/720, 1/120
    q = fma(q, r, 0x1.5555555555555p-5);                               // 1/24
    q = fma(q, r, 0x1.5555555555555p-3);                               // 1/6
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);                                 // expm1(r)
    const double res = fma(T, p, T);
    const int k = ki >> 6;
#if defined(__CUDA_ARCH__)
    // 2^k through the exponent field of the high word (32-bit integer ops)
    if (k < -1000) return __hiloint2double(__double2hiint(res) + ((k + 1000) << 20), __double2loint(res)) * 0x1p-1000;
    return __hiloint2double(__double2hiint(res) + (k << 20), __double2loint(res));
#else
    if (k < -1000) {                                            // result below 2^-1000: finish in two steps (denormals)
        const double s = phy_from_bits(phy_bits(res) + ((uint64_t)(int64_t)(k + 1000) << 52));
        return s * 0x1p-1000;
    }
    return phy_from_bits(phy_bits(res) + ((uint64_t)(int64_t)k << 52));
#endif
}

PHY_MATH_HD double phy_exp(double x) {
#if defined(__CUDA_ARCH__)
    return phy_exp_tab(x, d_phy_exp_tab);
#else
    return phy_exp_tab(x, h_phy_exp_tab);
#endif
}

PHY_MATH_HD double phy_log(double z) {
#if defined(__CUDA_ARCH__)
    int hi = __double2hiint(z), lo = __double2loint(z);
    int kadj = 0;
    if ((unsigned)(hi - 0x00100000) >= 0x7FE00000u) {           // zero, denormal, negative, inf, NaN
        if (((hi << 1) | lo) == 0) return -INFINITY;
        if (hi == 0x7FF00000 && lo == 0) return z;
        if (hi < 0 || (hi & 0x7FF00000) == 0x7FF00000) return z != z ? z : NAN;
        z *= 0x1p52;                                            // denormal: scale into the normal range
        hi = __double2hiint(z); lo = __double2loint(z);
        kadj = -52;
    }
    const int tmp = hi - 0x3FE68000;
    const int i = (tmp >> 13) & 127;
    const int k = (tmp >> 20) + kadj;
    const double m = __hiloint2double(hi - (tmp & (int)0xFFF00000), lo);
    const double invc = d_phy_log_tab[2 * i], logc = d_phy_log_tab[2 * i + 1];
#else
    uint64_t ix = phy_bits(z);
    int kadj = 0;
    if (ix - 0x0010000000000000ULL >= 0x7FE0000000000000ULL) {  // zero, denormal, negative, inf, NaN
        if ((ix << 1) == 0) return -INFINITY;
        if (ix == 0x7FF0000000000000ULL) return z;
        if ((ix >> 63) || (ix & 0x7FF0000000000000ULL) == 0x7FF0000000000000ULL) return z != z ? z : NAN;
        ix = phy_bits(z * 0x1p52);                              // denormal: scale into the normal range
        kadj = -52;
    }
    const uint64_t tmp = ix - 0x3FE6800000000000ULL;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52) + kadj;
    const double m = phy_from_bits(ix - (tmp & 0xFFF0000000000000ULL));
    const double invc = h_phy_log_tab[2 * i], logc = h_phy_log_tab[2 * i + 1];
#endif
    const double r = fma(m, invc, -1.0);
    const double kd = (double)k;
    double q = fma(r, -0x1.5555555555555p-3, 0x1.999999999999ap-3);    // -1/6, 1/5
    q = fma(q, r, -0.25);
    q = fma(q, r, 0x1.5555555555555p-2);                                // 1/3
    q = fma(q, r, -0.5);
    const double l1p = fma(r * r, q, r);                                // log1p(r)
    const double head = fma(kd, PHY_LN2_HI, logc);
    return head + fma(kd, PHY_LN2_LO, l1p);
}

}  // namespace phyfoo
