// umma_math.cuh — the scalar arithmetic of the tcgen05 ensemble integrator (batched_umma.cu) that decides its
// exactness claims: the branch-free sincos, the 52-bit fixed-point split and the recombination of the radix-256 digit
// sums.  Kept apart from the kernel so that tests/host_emul/umma_math_emul.cu can compile THIS source for the host and
// check it in the CPU-only suite (accuracy against libm, exactness against integer arithmetic).  On the device the
// constants come from the constant bank and the bit moves are the usual intrinsics; on the host both have plain
// equivalents.
#pragma once
#include <cstring>

#include "common.cuh"

namespace mcbb {

#define MCBB_UM_CONSTANTS                                                                                             \
    {0.6366197723675814,     /* 0  2/pi */                                                                             \
     6755399441055744.0,     /* 1  1.5 * 2^52 */                                                                       \
     1.5707963267948966,     /* 2  pi/2 high */                                                                        \
     6.123233995736766e-17,  /* 3  pi/2 low */                                                                         \
     1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04, \
     8.33333333332248946124e-03, -1.66666666666666324348e-01, /* 4..9  sin kernel */                                   \
     -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,                             \
     2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02, /* 10..15 cos kernel */      \
     1125899906842624.0,     /* 16 2^50 */                                                                             \
     0., 0., 0.}

// save times t_save (eval_ode_run's saveat grid), read with uniform or near-uniform indices: constant bank instead of
// 8 n_t bytes of every CTA's shared memory.  (Declared here, ahead of c_um, to keep the constant-bank layout — and with
// it the generated code — of the kernel unchanged.)
constexpr int UM_MAX_NT = 4096;
static __constant__ double c_um_ts[UM_MAX_NT];

// sincos / fixed-point constants live in the constant bank: DFMA takes them as c[bank][offset] operands, no
// per-use immediate moves
static __constant__ double c_um[20] = MCBB_UM_CONSTANTS;
static const double h_um[20] = MCBB_UM_CONSTANTS;

#ifdef __CUDA_ARCH__
#define UM_C(i) c_um[(i)]
#define UM_LO(x) __double2loint(x)
#define UM_HI(x) __double2hiint(x)
#define UM_HILO(h, l) __hiloint2double((h), (l))
#else
#define UM_C(i) h_um[(i)]
static inline int um_host_lo(double x) {
    long long b;
    std::memcpy(&b, &x, 8);
    return (int)(unsigned)(b & 0xffffffffll);
}
static inline int um_host_hi(double x) {
    long long b;
    std::memcpy(&b, &x, 8);
    return (int)(unsigned)((unsigned long long)b >> 32);
}
static inline double um_host_hilo(int h, int l) {
    const unsigned long long b = ((unsigned long long)(unsigned)h << 32) | (unsigned)l;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}
#define UM_LO(x) um_host_lo(x)
#define UM_HI(x) um_host_hi(x)
#define UM_HILO(h, l) um_host_hilo((h), (l))
#endif

// G independent sincos chains written side by side so that the scheduler interleaves them (one basic block);
// same arithmetic as sincos_fast2 (fixedpoint.cuh)
template <int G>
MCBB_HD void um_sincos_vec(const double (&x)[G], double* s, double* c) {
#pragma unroll
    for (int i = 0; i < G; i++) {
        const double qd = fma(x[i], UM_C(0), UM_C(1));
        const int q = UM_LO(qd);
        const double qf = qd - UM_C(1);
        double r = fma(-qf, UM_C(2), x[i]);
        r = fma(-qf, UM_C(3), r);
        const double z = r * r;
        double ps = fma(z, UM_C(4), UM_C(5));
        ps = fma(z, ps, UM_C(6));
        ps = fma(z, ps, UM_C(7));
        ps = fma(z, ps, UM_C(8));
        ps = fma(z, ps, UM_C(9));
        const double sr = fma(z * r, ps, r);
        double pc = fma(z, UM_C(10), UM_C(11));
        pc = fma(z, pc, UM_C(12));
        pc = fma(z, pc, UM_C(13));
        pc = fma(z, pc, UM_C(14));
        pc = fma(z, pc, UM_C(15));
        const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
        const double a = (q & 1) ? cr : sr;
        const double b = (q & 1) ? sr : cr;
        // quadrant signs on the integer pipe (the FP64 pipe is the scarce one)
        s[i] = UM_HILO(UM_HI(a) ^ ((q & 2) << 30), UM_LO(a));
        c[i] = UM_HILO(UM_HI(b) ^ (((q + 1) & 2) << 30), UM_LO(b));
    }
}

// x in [-1, 1] -> (low 32 bits, high 20 bits) of round(x 2^50) + 2^51, constants from the constant bank
MCBB_HD void um_fixed52(const double x, unsigned& lo, unsigned& hi) {
    const double t = fma(x, UM_C(16), UM_C(1));
    lo = (unsigned)UM_LO(t);
    hi = (unsigned)UM_HI(t) & 0xFFFFFu;
}

// sum_k v[k] 2^(8k) - deg 2^51 as a double (one rounding).  Digit sums are < 2^15 (<= 128 couplings of <= 255), so
// three of them pack into 31 bits; one wide multiply-add joins the two packs, v6 and the degree only touch the
// high word.  degh = deg << 19.
MCBB_HD double um_recombine(const unsigned v0, const unsigned v1, const unsigned v2, const unsigned v3,
                                               const unsigned v4, const unsigned v5, const unsigned v6, const int degh) {
    const unsigned lo3 = v0 + (v1 << 8) + (v2 << 16);
    const unsigned mid3 = v3 + (v4 << 8) + (v5 << 16);
    const unsigned long long w = (unsigned long long)mid3 * 16777216ull + lo3;
    const int hi = (int)(unsigned)(w >> 32) + ((int)(v6 << 16) - degh);
    return (double)(long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)w);
}

}  // namespace mcbb
