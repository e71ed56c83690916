// umma_common.cuh — plumbing shared by the tcgen05 ensemble integrators (batched_umma.cu: kuramoto_network,
// batched_umma_sl.cu: Stuart-Landau): tcgen05 / mbarrier wrappers (PTX ISA 8.7, sm_100a; encodings pinned on hardware by
// profiles/microbench/umma_probe.cu and mbar_probe.cu), the per-trajectory controller block, the Tsit5 dense-output
// polynomials and the KL-histogram measure of eval_ode_run (eval_mc_prob.jl:312-330).
#pragma once
#include "fixedpoint.cuh"
#include "lane_core.cuh"
#include "umma_math.cuh"  // after lane_core.cuh: keeps the constant-bank layout of the kernels

namespace mcbb {

constexpr int UM_NSLOT = 8;  // u, k1..k7
constexpr int UM_A_BYTES = 128 * 128;
constexpr int UM_MAX_KL_BINS = 64;

struct UmSlotCtl {
    long long run, fin_run, iter, nrej, natt;
    double t, dt, dts, lq, al1s, d0, d1, dt0, t_old, dts_old;  // lq = ln(qold) of the PI controller
    double hstep;  // step of the stage inputs: dts while running, the probe step dt0 while measuring the initial dt
    int mode, isave, ret, clipped, accept, fresh, fin, fin_ok, s0, s1;
};

// ---- tcgen05 / mbarrier plumbing (PTX ISA 8.7, sm_100a) ------------------------------------------------------
__device__ __forceinline__ unsigned um_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets in 16-byte units
__device__ __forceinline__ unsigned long long um_desc(const unsigned addr, const unsigned lbo, const unsigned sbo) {
    return (unsigned long long)((addr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo >> 4) & 0x3FFFu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// A operand from tensor memory (lane = row, 32-bit column = four consecutive K bytes)
__device__ __forceinline__ void um_mma_i8_ts(const unsigned d_tmem, const unsigned a_tmem, const unsigned long long db,
                                             const unsigned idesc, const unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void um_tmem_st32(const unsigned taddr, const unsigned (&w)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n\t"
        "tcgen05.wait::st.sync.aligned;" ::"r"(taddr),
        "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]),
        "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]),
        "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]),
        "r"(w[31])
        : "memory");
}

__device__ __forceinline__ void um_commit(const unsigned mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void um_mbar_wait(const unsigned mbar, const unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar), "r"(parity), "r"(20000u)  // suspend-time hint (ns): sleep in hardware until the phase flips
            : "memory");
    }
}

// 32 lanes x 32 columns (two trajectories), load and wait in one statement
__device__ __forceinline__ void um_tmem_ld32(const unsigned taddr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// 32 lanes x 16 columns (one trajectory)
__device__ __forceinline__ void um_tmem_ld16(const unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

// the same loads without the wait: issue several, then um_tmem_wait_ld() once
__device__ __forceinline__ void um_tmem_ld16_nowait(const unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void um_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// after um_tmem_wait_ld(): ties the loaded registers to the wait, so that no consumer is scheduled ahead of it (the
// compiler does not know that tcgen05.ld completes asynchronously)
__device__ __forceinline__ void um_tmem_landed(unsigned (&v)[16]) {
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                      "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// tensor-memory allocation: power of two >= 32 columns
template <int JB>
struct UmCols {
    static constexpr unsigned need = 16 * JB + 32;  // accumulators + the A operand
    static constexpr unsigned value = need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u;
};

// warp-level sum with a fixed tree (deterministic)
__device__ __forceinline__ double um_warp_sum(double p) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
    return p;
}

// Tsit5 dense-output weights b_q(theta) (OrdinaryDiffEq tsit_interpolants): b_0 = th (1 + th (A + th (B + th C))),
// b_q = th^2 (A + th (B + th C)) for q = 1..6; lane q of a warp evaluates polynomial q
static __constant__ double c_um_bth[7][3] = {
    {-2.763706197274826, 2.9132554618219126, -1.0530884977290216},
    {0.13169999999999998, -0.2234, 0.1017},
    {3.9302962368947516, -5.941033872131505, 2.490627285651253},
    {-12.411077166933676, 30.33818863028232, -16.548102889244902},
    {37.50931341651104, -88.1789048947664, 47.37952196281928},
    {-27.896526289197286, 65.09189467479366, -34.87065786149661},
    {1.5, -4., 2.5}};

// a / b for normal, well-scaled b without the compiler's slow-path branch: reciprocal seed, two Newton steps,
// one residual correction (used only inside the step-size error norm)
__device__ __forceinline__ double um_div(const double a, const double b) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);
}

// KL divergence of the saved samples of one state dimension to the normal distribution with their mean and standard
// deviation, on the histogram of empirical_1D_KL_divergence_hist (eval_mc_prob.jl:312-330; same arithmetic as the lane
// kernel, lane_core.cuh).  Sample k of the dimension is smp[k * stride], k = 1 .. n_t - 1.
// The per-thread histogram lives in shared memory when the caller can lend some (hs[bin * 128 + thread]): a local
// array indexed by a data-dependent bin scatters over 32 cache lines per warp access.
static __device__ __noinline__ double um_kl_hist(const double* smp, const long long stride, const int n_t, const double mu, const double sig,
                             const int nb, const double nstd, const double tol, unsigned short* hs) {
    const double NaN = nan("");
    const double kq = (nb - 1) / 2.;
    const double bw = (nstd * sig) / kq;
    const double start = mul_add_2r(-kq, bw, mu), stop = mul_add_2r(kq, bw, mu);
    int nc = 0;
    if (sig >= tol) nc = julia_range_len(start, bw, stop);
    if (nc > nb) nc = nb;
    if (sig != sig) return NaN;  // NaN std (non-finite samples): the reference's KL is NaN
    if (nc <= 0) return 0.;      // sig < sig_tol
    int hist[UM_MAX_KL_BINS];
    if (hs) {
        for (int b = 0; b < nc; b++) hs[b * 128] = 0;
    } else {
        for (int b = 0; b < nc; b++) hist[b] = 0;
    }
    const double hb = bw / 2.;
    const double last_edge = mul_add_2r(kq, bw, hb);  // `mu` missing: eval_mc_prob.jl:328
    const double e1 = start - hb, ibw = 1. / bw;
    auto edge = [&](const int m) { return fma((double)(m - 1), bw, start) - hb; };  // edges 1..nc (ascending)
    // fit(Histogram, u, edges; closed=:left) = searchsortedlast on the vector [edge(1..nc), last_edge].  Its bisection
    // probes last_edge only after x >= edge(nc) (everything before it is ascending), so the result is: r = position
    // among the ascending edges — taken directly and corrected against the very same edge expression — and, when
    // r == nc, bin nc if x < last_edge, else outside.  Eight samples are fetched at a time (the scratch lives in
    // L2 / HBM: keep the loads in flight).
    constexpr int KLC = 8;  // samples in flight per thread (16 measured slower: 357 k vs 376 k traj/s on C5 with KL)
    for (int k0 = 1; k0 < n_t; k0 += KLC) {
        double xs[KLC];
#pragma unroll
        for (int q = 0; q < KLC; q++) xs[q] = (k0 + q < n_t) ? __ldcg(smp + (long long)(k0 + q) * stride) : nan("");
#pragma unroll
        for (int q = 0; q < KLC; q++) {
            const double x = xs[q];
            if (!(x == x)) continue;  // NaN compares false everywhere: the bisection runs off the right end
            const double f = floor((x - e1) * ibw) + 1.;
            int r = f < 0. ? 0 : (f > (double)nc ? nc : (int)f);
            while (r >= 1 && x < edge(r)) r--;
            while (r < nc && x >= edge(r + 1)) r++;
            if (r >= 1 && (r < nc || x < last_edge)) {
                if (hs)
                    hs[(r - 1) * 128] += 1;
                else
                    hist[r - 1] += 1;
            }
        }
    }
    double tot = 0., rs = 0.;
    double qn[UM_MAX_KL_BINS];  // unnormalised normal weights of the bin centres
    for (int b = 0; b < nc; b++) {
        if (hs) hist[b] = hs[b * 128];
        tot += (double)hist[b];
        const double z = (fma((double)b, bw, start) - mu) / sig;
        qn[b] = exp(-(z * z) / 2.);
        rs += qn[b];
    }
    if (tot == 0.) return NaN;  // 0/0 weights -> NaN, as StatsBase would give
    double kl = 0.;
    for (int b = 0; b < nc; b++) {
        const double p = (double)hist[b] / tot;
        if (p > 0.) {
            const double q = qn[b] / rs;
            kl += p * log(p / q);
        }
    }
    return kl;
}


}  // namespace mcbb
