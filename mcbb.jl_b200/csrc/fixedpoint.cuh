// fixedpoint.cuh — pieces shared by the two tensor-core ensemble integrators (batched_imma.cu: mma.sync,
// batched_umma.cu: tcgen05): the branch-free sincos and the 52-bit fixed-point split / recombination that makes
// the coupling sums  P = A sin(Y), Q = A cos(Y)  (src/systems.jl:119-130) exact integer GEMMs.
#pragma once
#include "common.cuh"

namespace mcbb {

// branch-free sincos: two-term Cody-Waite reduction by pi/2, then the fdlibm kernel polynomials on [-pi/4, pi/4]
__device__ __forceinline__ void sincos_fast2(const double x, double* s, double* c) {
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
    const double qd = fma(x, 0.6366197723675814, magic);
    const int q = __double2loint(qd);
    const double qf = qd - magic;
    double r = fma(-qf, 1.5707963267948966, x);
    r = fma(-qf, 6.123233995736766e-17, r);
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sr = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double a = (q & 1) ? cr : sr;
    const double b = (q & 1) ? sr : cr;
    *s = (q & 2) ? -a : a;
    *c = ((q + 1) & 2) ? -b : b;
}

// x in [-1, 1] -> u = round(x * 2^50) + 2^51 as (low 32 bits, high 20 bits).  The eight little-endian bytes of
// u are the radix-256 digits d_0..d_6 (d_7 = 0) of the fixed-point value.
__device__ __forceinline__ void fixed52(const double x, unsigned& lo, unsigned& hi) {
    const double t = fma(x, 1125899906842624.0 /* 2^50 */, 6755399441055744.0 /* 1.5 * 2^52 */);
    lo = (unsigned)__double2loint(t);
    hi = (unsigned)__double2hiint(t) & 0xFFFFFu;
}

// sum_k v[k] 2^(8k), k = 0..6, for digit sums v[k] < 2^15 (at most 128 non-zero couplings per row)
__device__ __forceinline__ long long recombine7(const unsigned v0, const unsigned v1, const unsigned v2, const unsigned v3,
                                                const unsigned v4, const unsigned v5, const unsigned v6) {
    const unsigned w0 = v0 + (v1 << 8), w1 = v2 + (v3 << 8), w2 = v4 + (v5 << 8);
    return (long long)((((((unsigned long long)v6 << 16) + w2) << 16) + w1) << 16) + (long long)w0;
}

}  // namespace mcbb
