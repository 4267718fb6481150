// n1n_math.cuh -- arithmetic helpers shared by the node-per-thread network kernel (net1n.cuh), the F81-structured order-0 pass
// (tree_pass_f81.cuh) and the run-time compiled order-0 kernels (jit_o0.hpp embeds this file for NVRTC): four exponentials at
// once, reciprocal without special cases, log z accumulated as mantissa product + exponent sum.  No host includes.
#pragma once
#include "fp64_math.cuh"

#if defined(__CUDACC__)
#define N1N_HD __host__ __device__ __forceinline__
#define N1N_RARE __host__ __device__ __noinline__      // rare paths stay out of the step loop's instruction stream
#else
#define N1N_HD inline
#define N1N_RARE inline
#endif

namespace phyfoo {

struct N1nAcc {
    double Fq;      // sum over this thread's nodes of sum_a q(a) g(a)
    double M;       // product of the mantissas of z
    int E;          // sum of the exponents of z
    double Fslow;   // sum of log z of nodes whose z is not a normal number (zero, denormal, inf, NaN)
};

// rows: a q row is 32 doubles = [half 0: 8 windows x (q0, q1)] [half 1: 8 windows x (q2, q3)]; `reg` already points at this
// thread's window (2 * w doubles in), so its four values sit at +0, +1, +16, +17.  An AB row is 64 doubles = four such halves
// (A0 A1 | A2 A3 | B0 B1 | B2 B3).  16-byte accesses: LDS.128 / STS.128, 8 windows x 16 B = all 32 banks once.
struct alignas(16) N1nD2 { double x, y; };
N1N_HD void n1n_ld4(const double* p, double (&v)[4]) {
    const N1nD2 a = *reinterpret_cast<const N1nD2*>(p), b = *reinterpret_cast<const N1nD2*>(p + 16);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
N1N_HD void n1n_st4(double* p, const double (&v)[4]) {
    N1nD2 a, b; a.x = v[0]; a.y = v[1]; b.x = v[2]; b.y = v[3];
    *reinterpret_cast<N1nD2*>(p) = a; *reinterpret_cast<N1nD2*>(p + 16) = b;
}
N1N_HD double n1n_sum4(const double (&v)[4]) { return ((v[0] + v[1]) + v[2]) + v[3]; }

N1N_HD int n1n_hi(double z) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(z);
#else
    return (int)(phy_bits(z) >> 32);
#endif
}
N1N_HD double n1n_with_hi(double z, int hi) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, __double2loint(z));
#else
    return phy_from_bits(((uint64_t)(uint32_t)hi << 32) | (phy_bits(z) & 0xffffffffULL));
#endif
}
N1N_HD double n1n_rcp(double z) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(z);
#else
    return 1.0 / z;
#endif
}
// reciprocal of a normal number away from the ends of the exponent range: the seed and the two Newton steps of the
// library's __drcp_rn without its special-case tests (the caller has looked at the exponent already)
N1N_HD double n1n_rcp_normal(double z) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z));
    double e = fma(-z, r, 1.0);
    r = fma(r, e, r);
    e = fma(-z, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / z;
#endif
}
struct N1nD4 { double v0, v1, v2, v3; };
N1N_RARE N1nD4 n1n_exp4_rare(double x0, double x1, double x2, double x3, const double* tab64) {
    N1nD4 r;
    r.v0 = phy_exp_tab(x0, tab64); r.v1 = phy_exp_tab(x1, tab64); r.v2 = phy_exp_tab(x2, tab64); r.v3 = phy_exp_tab(x3, tab64);
    return r;
}
// z = sum of the exponentials is zero, denormal, huge, inf or NaN: library reciprocal, log z taken directly
struct N1nZR { double rz, lz; };
N1N_RARE N1nZR n1n_z_rare(double z) {
    N1nZR r;
    r.rz = n1n_rcp(z); r.lz = phy_log(z);
    return r;
}
N1N_HD bool n1n_exp_common(const double (&x)[4]);
// four exponentials at once: the common case (every |x| < 690) is straight-line code, so that the four polynomial chains
// interleave; phy_exp_tab's arithmetic
N1N_HD void n1n_exp4(const double (&x)[4], const double* tab64, double (&e)[4]) {
    if (n1n_exp_common(x)) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double kd = fma(x[a], PHY_64_OVER_LN2, 0x1.8p52);       // (phy_exp_tab rounds the product first: same k except at ties)
#if defined(__CUDA_ARCH__)
            const int32_t ki = __double2loint(kd);
#else
            const int32_t ki = (int32_t)(uint32_t)phy_bits(kd);
#endif
            kd -= 0x1.8p52;
            double r = fma(kd, -PHY_LN2_64_HI, x[a]);
            r = fma(kd, -PHY_LN2_64_LO, r);
            const double T = tab64[ki & 63];
            double q = fma(r, 0x1.6c16c16c16c17p-10, 0x1.1111111111111p-7);
            q = fma(q, r, 0x1.5555555555555p-5);
            q = fma(q, r, 0x1.5555555555555p-3);
            q = fma(q, r, 0.5);
            const double pm = fma(r * r, q, r);
            const double res = fma(T, pm, T);
            e[a] = n1n_with_hi(res, n1n_hi(res) + ((ki >> 6) << 20));       // |x| < 690: the result is a normal number
        }
    } else {
        const N1nD4 r = n1n_exp4_rare(x[0], x[1], x[2], x[3], tab64);
        e[0] = r.v0; e[1] = r.v1; e[2] = r.v2; e[3] = r.v3;
    }
}
// The same with a 16-entry table 2^(j/16) (every fourth entry of the 64-entry one) and one more term of the polynomial
// (|r| <= ln 2 / 32: the first neglected term r^8 / 8! is below 2e-18): 16 doubles are 128 bytes, one bank pair per entry, so a
// warp's look-up is one shared-memory wavefront whatever the 32 indices are (64 entries: 4.6 wavefronts on average in the
// order-1 node kernel, whose busiest unit is the shared-memory data path -- profiles/r02_c3_l2window_summary.md).
N1N_RARE N1nD4 n1n_exp4_rare_g(double x0, double x1, double x2, double x3) {
    N1nD4 r;
    r.v0 = phy_exp(x0); r.v1 = phy_exp(x1); r.v2 = phy_exp(x2); r.v3 = phy_exp(x3);
    return r;
}
N1N_HD void n1n_exp4_t16(const double (&x)[4], const double* tab16, double (&e)[4]) {
    if (n1n_exp_common(x)) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double kd = fma(x[a], 0.25 * PHY_64_OVER_LN2, 0x1.8p52);
#if defined(__CUDA_ARCH__)
            const int32_t ki = __double2loint(kd);
#else
            const int32_t ki = (int32_t)(uint32_t)phy_bits(kd);
#endif
            kd -= 0x1.8p52;
            double r = fma(kd, -4.0 * PHY_LN2_64_HI, x[a]);
            r = fma(kd, -4.0 * PHY_LN2_64_LO, r);
            const double T = tab16[ki & 15];
            double q = fma(r, 0x1.a01a01a01a01ap-13, 0x1.6c16c16c16c17p-10);      // 1/5040, 1/720
            q = fma(q, r, 0x1.1111111111111p-7);
            q = fma(q, r, 0x1.5555555555555p-5);
            q = fma(q, r, 0x1.5555555555555p-3);
            q = fma(q, r, 0.5);
            const double pm = fma(r * r, q, r);
            const double res = fma(T, pm, T);
            e[a] = n1n_with_hi(res, n1n_hi(res) + ((ki >> 4) << 20));       // |x| < 690: the result is a normal number
        }
    } else {
        const N1nD4 r = n1n_exp4_rare_g(x[0], x[1], x[2], x[3]);
        e[0] = r.v0; e[1] = r.v1; e[2] = r.v2; e[3] = r.v3;
    }
}
// log z accumulated as (mantissa product, exponent sum)
N1N_HD void n1n_acc_logz(N1nAcc& acc, double z) {
    const int hi = n1n_hi(z);
    const unsigned ef = ((unsigned)hi >> 20);                    // sign + exponent field: a positive normal number has 1..0x7fe
    if (ef - 1u < 0x7feu) {
        acc.E += (int)ef - 1023;
        acc.M *= n1n_with_hi(z, (hi & 0x000fffff) | 0x3ff00000);
    } else acc.Fslow += phy_log(z);
}
// fold the exponent of the running mantissa product into E (every 64 nodes at the latest: 2^64 is far from overflow)
N1N_HD void n1n_fold(N1nAcc& acc) {
    const int hi = n1n_hi(acc.M);
    acc.E += (hi >> 20) - 1023;
    acc.M = n1n_with_hi(acc.M, (hi & 0x000fffff) | 0x3ff00000);
}
N1N_HD double n1n_free_energy(const N1nAcc& acc) {
    const double lz = fma((double)acc.E, PHY_LN2_HI, fma((double)acc.E, PHY_LN2_LO, phy_log(acc.M)));
    return (acc.Fq - acc.Fslow) - lz;
}

// message of a node to its tree parent: msg(a) = q(a) A(a) + sum_{a' != a} q(a') B(a') = q(a) D(a) + sum_a' q(a') B(a')
N1N_HD void n1n_message(const double (&q)[4], const double (&D)[4], const double (&B)[4], double (&msg)[4]) {
    double sw = q[0] * B[0];
    sw = fma(q[1], B[1], sw); sw = fma(q[2], B[2], sw); sw = fma(q[3], B[3], sw);
#pragma unroll
    for (int a = 0; a < 4; ++a) msg[a] = fma(q[a], D[a], sw);
}

// |x| < 690 for all four: integer compares on the high words (690.0 = 0x4085900000000000)
N1N_HD bool n1n_exp_common(const double (&x)[4]) {
    return ((n1n_hi(x[0]) & 0x7fffffff) < 0x40859000) & ((n1n_hi(x[1]) & 0x7fffffff) < 0x40859000) &
           ((n1n_hi(x[2]) & 0x7fffffff) < 0x40859000) & ((n1n_hi(x[3]) & 0x7fffffff) < 0x40859000);
}

}  // namespace phyfoo
