// batched_umma.cu — ensemble Tsit5 for kuramoto_network (src/systems.jl:119-130) with a unit-weight adjacency,
// coupling sums on the 5th-generation tensor cores (tcgen05.mma kind::i8, A operand and accumulators in tensor
// memory), statistics of eval_ode_run (mean / std / KL-hist, eval_mc_prob.jl:191-198, 312-330) fused into the step.
//
// Same exact fixed-point algorithm as batched_imma.cu (see fixedpoint.cuh): per right-hand-side evaluation
//        P = A sin(Y),  Q = A cos(Y)          A: N x N in {0,1},  sin/cos(Y): N x JB
// with sin/cos written as u = round(x 2^50) + 2^51 and split into radix-256 digits.  What changes is the
// machine mapping, which is built around tcgen05 instead of mma.sync:
//
//   * D[128 x 16 JB] (s32, TMEM) = A[128 x 128] (u8, resident in TMEM: row r in lane r, four K bytes per 32-bit
//     column) * B[128 x 16 JB] (u8, shared memory, MN-major, no swizzle).  A in tensor memory instead of shared
//     memory frees 16 KB per CTA: 24 instead of 16 resident trajectories per SM.
//     In the MN-major canonical layout the 16 columns of one trajectory are 16 CONTIGUOUS bytes per oscillator
//     row, and the digits of (sin, cos) are simply the little-endian bytes of the two 64-bit integers
//     {u_sin, u_cos}: the producer writes its element with ONE 16-byte shared-memory store, no byte shuffles.
//   * one thread per oscillator row (thread r <-> TMEM lane r, tcgen05.ld 32x32b): the thread that produced
//     sin/cos of element (r, j) is the one that reads row r of P, Q back, so sin/cos and the Welford state stay
//     in registers and every Runge-Kutta stage vector is thread-private shared memory.  Element ownership
//     never changes => the stage loop needs no CTA-wide barrier (warps publish their digits on a 4-count
//     mbarrier, the last one to arrive issues the MMA, the others sleep on a named barrier until its completion
//     mbarrier flips); 2 CTA barriers per step remain, around the step controller (19 in the mma.sync kernel).
//   * the MMA is asynchronous; four CTAs per SM keep the FP64 pipe busy while another CTA's MMA round trip
//     (~450 clocks measured, profiles/microbench/umma_probe.cu) is in flight.
//
// Operand encodings were pinned bit-exactly on hardware with profiles/microbench/umma_probe.cu (shared-memory
// descriptors, instruction descriptor, A from tensor memory), mbarrier.arrive / pending_count semantics with
// mbar_probe.cu, FP64 latency / issue numbers with dfma_latency.cu.
#include <algorithm>
#include <cmath>
#include <vector>

#include "umma_common.cuh"

namespace mcbb {

struct UmmaArgs {
    int N, KS, n_sms;
    double inv_n, ln_qoldinit;
    const unsigned char* a_img;  // [128][128] A, row-major bytes (row r goes to TMEM lane r)
    const int* deg;              // [128] number of non-zero couplings per row
    const double* bias;          // [128]
    const int* filter_pos;       // [128]
    double unit_weight;
    double scalars[MCBB_N_SCALARS];
    int n_par;
    int par_slot[MCBB_MAX_PAR];
    SolverDev o;
    int n_meas, meas[MCBB_MAX_MEAS], n_filter, cyclic;
    SolveIO io;
    unsigned long long* timing;  // [8] phase clocks (TIMING builds only)
    // empirical_1D_KL_divergence_hist (eval_mc_prob.jl:312-330): the saved samples of the runs of a chunk are kept in
    // a global scratch [run][n_t][N], mean / std / success of every run in `stat` / `okflag`, and binned after the solve by
    // k_umma_kl_post (one thread per run and oscillator) — inside the integrator the binning of a finished run stalled
    // the five other trajectories of its CTA
    int need_kl, kl_bins;
    double kl_nstd, kl_tol;
    double* smp;
    double* stat;                // [runs of the chunk][N][2] mean, std
    int* okflag;                 // [runs of the chunk]
    long long run_begin, run_end;  // KL instantiation: the runs of this launch
    int lw_mode;                 // light-warp kernel only (batched_umma_lw.cuh)
};

// TIMING = true: thread 0 of every CTA accumulates clock64() per phase into a.timing (development aid,
// MCBB_UMMA_TIMING=1; the compiler may move arithmetic across the ticks, so treat the split as approximate)
// KL = true: the instantiation that keeps the saved samples and evaluates the KL-histogram measure (a separate
// instantiation: the call and the sample stores cost the plain mean / std kernel ~5 % through register allocation)
// NW = warps per tensor-memory lane quadrant.  NW = 1: one thread per oscillator row carries all JB trajectories of the
// CTA.  NW = 2: two threads per row (warps w and w + 4 may both address TMEM lanes 32 (w % 4) ..), each carrying JB / 2
// trajectories: twice the warps per scheduler for the same tensor memory and shared memory, half the serial work per
// warp and stage.
template <int JB, int MINB, bool TIMING = false, bool KL = false, int NW = 1>
__global__ void __launch_bounds__(128 * NW, MINB) k_solve_batched_umma(const UmmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NT = 128 * NW;          // threads per CTA
    constexpr int JBT = JB / NW;          // trajectories per thread
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int wq = wid & 3;               // tensor-memory lane quadrant of this warp
    const int j0 = (NW == 1) ? 0 : (wid >> 2) * JBT;  // first trajectory slot of this thread (warp-uniform)
    const int N = a.N;
    const int row = tid & 127;
    const bool valid = row < N;
    const int rr = valid ? row : N;  // idle lanes (rows >= N of the MMA tile) work on a dummy row: no branches
    const SolverDev& o = a.o;
    const int n_t = a.io.n_t;
    const long long n_mc = a.io.n_mc;
    constexpr unsigned NCOL = 16 * JB;            // MMA N: 16 accumulator columns per trajectory (7 + 7 digits, 2 zero)
    constexpr int G = JBT <= 4 ? JBT : (JBT % 3 == 0 ? 3 : 4);  // trajectories whose dependent chains are interleaved (2 measures the same)
    constexpr int RS = UM_NSLOT * JB + 2;         // row stride (doubles) = 2 (mod 4): 16-byte aligned thread-private
                                                  // blocks whose 128-bit accesses are bank-conflict free
    static_assert(JB % NW == 0 && JBT % G == 0, "trajectories per thread must be a multiple of the interleave group");
    static_assert(!KL || NW == 1, "the KL instantiation lends the digit panel to 128 per-thread histograms");

    unsigned char* sB = smem;                                   // [JB][128][16]
    double* sm_k = (double*)(sB + JB * 2048);                   // [N + 1][JB][8] (+2 pad)  u, k1..k7 per element
    const double* const sm_ts = c_um_ts;                        // [n_t] (constant bank)
    double* sm_red0 = sm_k + (size_t)(N + 1) * RS;              // [4][JB]
    double* sm_red1 = sm_red0 + 4 * JB;                         // [4][JB]
    UmSlotCtl* ctl = (UmSlotCtl*)(sm_red1 + 4 * JB);            // [JB]
    unsigned long long* mbar_p = (unsigned long long*)(ctl + JB);
    unsigned* tmem_slot = (unsigned*)(mbar_p + 2);
    unsigned long long* full_p = mbar_p + 1;  // "digits published": one arrival per row warp
    const unsigned mbar_full = um_smem_u32(full_p);
    const unsigned mbar = um_smem_u32(mbar_p);
    double* const kb = sm_k + (size_t)rr * RS;

#define KP(slot, j) kb[(j)*UM_NSLOT + (slot)]

    for (int i = tid; i < JB * 2048 / 16; i += NT) ((uint4*)sB)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < (N + 1) * RS; i += NT) sm_k[i] = 0.;  // stage slots are read before written (x 0)
    // the controller lanes sit in a different warp (scheduler partition) for each CTA resident on an SM
    const int cw = (blockIdx.x / a.n_sms) & (4 * NW - 1);
    const bool is_ctl = (wid == cw) && lane < JB;
    if (is_ctl) {
        UmSlotCtl& c = ctl[lane];
        c.mode = MODE_FETCH;
        c.accept = c.fresh = c.fin = 0;
        c.hstep = 0.;
        c.dts = 0.;
        c.al1s = 0.;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_full), "n"(4 * NW));
    }
    // tensor memory: one power-of-two allocation for the accumulators and the A operand — or, for five CTAs per SM with
    // JB = 4, two (64 accumulator columns + 32 columns of A = 96: five CTAs fit the 512 columns, four of 128 would not)
    constexpr bool SPLIT = (MINB == 5 && NCOL == 64);
    if (wid == 0) {
        if (SPLIT) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(um_smem_u32(tmem_slot)));
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(um_smem_u32(tmem_slot + 1)));
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(tmem_slot)),
                         "n"(UmCols<JB>::value));
        }
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // A image and zeros -> visible to the MMA's proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const unsigned tmem_row = tmem_base + ((unsigned)(32 * wq) << 16);  // this warp's 32 TMEM lanes
    const unsigned tmem_a = SPLIT ? tmem_slot[1] : tmem_base + NCOL;      // A operand: 32 columns behind the accumulators
    {  // row r of A -> lane r: 128 K bytes = 32 columns (the first warp of every quadrant writes it)
        unsigned w[32];
        if (wid < 4) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint4 t4 = reinterpret_cast<const uint4*>(a.a_img + (size_t)row * 128)[q];
            w[4 * q] = t4.x;
            w[4 * q + 1] = t4.y;
            w[4 * q + 2] = t4.z;
            w[4 * q + 3] = t4.w;
        }
        um_tmem_st32(tmem_a + ((unsigned)(32 * wq) << 16), w);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // MMA operands: A from tensor memory (next K step + 8 columns); B MN-major (16-column chunks 2048 B apart, 8-deep K groups 128 B apart, next K step + 512 B)
    const unsigned idesc = (2u << 4) /* D s32 */ | (0u << 7) /* A u8 */ | (0u << 10) /* B u8 */ | (0u << 15) /* A K-major */ |
                           (1u << 16) /* B MN-major */ | ((NCOL >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned sB_addr = um_smem_u32(sB);
    const int KSTEPS = a.KS;
    unsigned char* const sB_row = sB + row * 16;

    const double bias = a.bias[row];
    const int degh = a.deg[row] << 19;  // deg 2^51 in units of the high word
    const int fpos = a.filter_pos[row];
    unsigned phase = 0;

    // thread-private state of the JB elements of this oscillator row
    double sv[JBT], cv[JBT], un[JBT], am[JBT], a2[JBT], as[JBT];
#pragma unroll
    for (int j = 0; j < JBT; j++) sv[j] = cv[j] = un[j] = am[j] = a2[j] = as[j] = 0.;

    auto begin_attempt = [&](UmSlotCtl& c) -> bool {
        if (++c.iter > o.maxiters) {
            c.ret = MCBB_RET_MAXITERS;
            return false;
        }
        c.dts = c.dt;
        c.clipped = 0;
        if (c.dts >= o.t1 - c.t) {
            c.dts = o.t1 - c.t;
            c.clipped = 1;
        }
        if (!c.clipped && c.dts <= o.dtmin) {
            c.ret = MCBB_RET_DTLESSTHANMIN;
            return false;
        }
        c.hstep = c.dts;
        return true;
    };

    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
    if (TIMING) tprev = clock64();
#define UM_TICK(k)                              \
    if (TIMING && tid == 0) {                   \
        const long long tn_ = clock64();        \
        tph[k] += tn_ - tprev;                  \
        tprev = tn_;                            \
    }
    for (;;) {
        // ---- per-trajectory control: PI controller, accept / reject, save window, finish, refill -----------------
        if (is_ctl) {
            UmSlotCtl& c = ctl[lane];
            c.accept = 0;
            c.fin = 0;
            c.fresh = 0;
            int finished = 0;
            if (c.mode == MODE_INIT) {
                c.mode = MODE_RUN;
                if (!begin_attempt(c)) finished = 1;
            } else if (c.mode == MODE_RUN) {
                c.natt++;
                double e2 = 0., nb = 0.;
                for (int w = 0; w < 4; w++) {
                    e2 += sm_red0[w * JB + lane];
                    nb += sm_red1[w * JB + lane];
                }
                // PI controller (OrdinaryDiffEq stepsize_controller!/step_accept_controller! for PIController):
                //   q = EEst^beta1 / qold^beta2 / gamma clamped to [1/qmax, 1/qmin], dt <- dt / q
                // evaluated as dt * clamp(gamma exp(beta2 ln qold - beta1 ln EEst), qmin, qmax) with
                // ln EEst = ln(e2 / N) / 2: one log and one exp on the critical path, no sqrt / pow / division.
                const double m2 = e2 * a.inv_n;  // EEst^2
                const double L = 0.5 * log(m2);
                if (!(m2 > 1.)) {
                    const double r = (m2 == 0.) ? o.qmax : fmin(o.qmax, fmax(o.qmin, o.gamma * exp(fma(o.beta2, c.lq, -o.beta1 * L))));
                    c.lq = fmax(L, a.ln_qoldinit);
                    const double tnew = c.clipped ? o.t1 : c.t + c.dts;
                    int k1 = c.isave;
                    while (k1 < n_t && sm_ts[k1] <= tnew) k1++;
                    c.s0 = c.isave;
                    c.s1 = k1;
                    c.isave = k1;
                    c.t_old = c.t;
                    c.dts_old = c.dts;
                    c.t = tnew;
                    c.dt = fmin(o.dtmax, fmax(o.dtmin, c.dts * r));
                    c.accept = 1;
                    if (nb > 0.) c.ret = MCBB_RET_UNSTABLE;  // NaN in the new state
                } else {
                    c.nrej++;
                    c.dt = c.dts * fmax(o.qmin, o.gamma * exp(-o.beta1 * L));
                }
                if (c.ret != 0 || (c.accept && !(c.t < o.t1)))
                    finished = 1;
                else if (!begin_attempt(c))
                    finished = 1;
            }
            if (finished) {
                c.fin = 1;
                c.fin_run = c.run;
                c.fin_ok = (c.ret == 0) && c.isave >= n_t;
                if (a.io.retcode) a.io.retcode[c.run] = c.ret;
                if (a.io.nsteps) {
                    a.io.nsteps[c.run] = c.natt;
                    a.io.nsteps[n_mc + c.run] = c.nrej;
                }
                c.mode = MODE_FETCH;
            }
            if (c.mode == MODE_FETCH) {
                const long long run = (long long)atomicAdd(a.io.work_counter, 1ULL);
                if (run >= (KL ? a.run_end : n_mc)) {
                    c.mode = MODE_DONE;
                } else {
                    double sc[MCBB_N_SCALARS];
                    for (int q = 0; q < MCBB_N_SCALARS; q++) sc[q] = a.scalars[q];
                    for (int p = 0; p < a.n_par; p++) sc[a.par_slot[p]] = a.io.par[(long long)p * n_mc + run];
                    // K/N (systems.jl:128) times the coupling value, times the 2^-50 of the fixed-point sums
                    c.al1s = ((sc[0] / (double)N) * a.unit_weight) * 8.881784197001252e-16;
                    c.run = run;
                    c.t = o.t0;
                    c.iter = 0;
                    c.nrej = 0;
                    c.natt = 0;
                    c.lq = a.ln_qoldinit;
                    c.hstep = 0.;
                    c.isave = 0;
                    c.ret = 0;
                    c.mode = MODE_INIT;
                    c.fresh = 1;
                }
            }
        }
        const bool all_done = __syncthreads_and(!is_ctl || ctl[lane].mode == MODE_DONE);
        UM_TICK(0)  // error partials + controller + barriers

        // ---- row owners: samples of the accepted step, u <- u_new, results of finished runs, new initial state ----
        bool any_init = false, any_fin = false;
        unsigned runmask = 0u;  // bit j: slot j integrates (its stage slots are live)
#pragma unroll
        for (int j = 0; j < JB; j++) {
            const UmSlotCtl& c = ctl[j];
            any_init |= (c.mode == MODE_INIT);
            runmask |= (c.mode == MODE_RUN ? 1u : 0u) << j;
            any_fin |= (c.fin != 0);
        }
#pragma unroll
        for (int jj = 0; jj < JBT; jj++) {
            const int j = j0 + jj;
            const UmSlotCtl& c = ctl[j];
            if (c.accept) {  // uniform over the warps that carry slot j; idle lanes run along on the dummy row
                const int ks0 = c.s0 < 1 ? 1 : c.s0, ks1 = c.s1;  // sample 0 is dropped (eval_mc_prob.jl:152)
                if (ks0 < ks1) {
                    const double dts = c.dts_old, t_old = c.t_old;
                    double kk[8];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const double2 t2 = reinterpret_cast<const double2*>(kb + j * UM_NSLOT)[q];
                        kk[2 * q] = t2.x;
                        kk[2 * q + 1] = t2.y;
                    }
                    const int lq = lane < 7 ? lane : 0;
                    const double pa = c_um_bth[lq][0], pb = c_um_bth[lq][1], pc = c_um_bth[lq][2];
                    for (int k = ks0; k < ks1; k++) {
                        // dense-output weights: lane q evaluates b_q(theta), a shuffle hands all seven to every row
                        const double th = um_div(sm_ts[k] - t_old, dts);
                        const double inner = fma(th, fma(th, pc, pb), pa);
                        const double bq = (lq == 0) ? th * fma(th, inner, 1.) : (th * th) * inner;
                        double lin = kk[1] * __shfl_sync(0xffffffffu, bq, 0);
#pragma unroll
                        for (int q = 1; q < 7; q++) lin = fma(kk[1 + q], __shfl_sync(0xffffffffu, bq, q), lin);
                        double x = fma(dts, lin, kk[0]);
                        if (a.cyclic) {  // _cyclic_setback!, eval_mc_prob.jl:286-296
                            if (k == 1) {
                                const double twopi = 6.283185307179586, pi = 3.141592653589793;
                                double vv = julia_mod_pos(x, twopi);
                                if (vv > pi) vv -= twopi;
                                as[jj] = x - vv;
                                x = vv;
                            } else {
                                x = x - as[jj];
                            }
                        }
                        // (a run that finished with this step has already handed its slot to the next run: fin_run)
                        if (KL && valid) a.smp[(((size_t)((c.fin ? c.fin_run : c.run) - a.run_begin)) * n_t + k) * N + row] = x;
                        const double delta = x - am[jj];
                        if (fabs(delta) <= 1.7976931348623157e308) {
                            const double mnew = fma(delta, um_div(1., (double)k), am[jj]);
                            a2[jj] = fma(delta, x - mnew, a2[jj]);
                            am[jj] = mnew;
                        } else {
                            am[jj] = nonfinite_mean(am[jj], x);
                            a2[jj] = nan("");
                        }
                    }
                }
                // the accepted state was evaluated with the error estimate and carried in registers across the controller
                // barrier (recomputing it here instead measured 2 % slower: 483 k against 493 k traj/s)
                KP(0, j) = un[jj];
                KP(1, j) = KP(7, j);  // FSAL
            }
            if (c.fin && valid && fpos >= 0) {
                const bool ok = c.fin_ok != 0;
                const double mean = ok ? am[jj] : nan("");
                const double sd = ok ? sqrt((a2[jj] < 0. ? 0. : a2[jj]) / (double)(n_t - 2)) : nan("");
                if (KL) {
                    const size_t rl = (size_t)(c.fin_run - a.run_begin);
                    *reinterpret_cast<double2*>(a.stat + (rl * N + row) * 2) = make_double2(mean, sd);
                    if (row == 0) a.okflag[rl] = ok ? 1 : 0;
                }
                for (int m = 0; m < a.n_meas; m++)
                    if (!KL || a.meas[m] != MCBB_MEAS_KL_HIST)  // the KL measures are written by k_umma_kl_post
                        a.io.feat[((long long)m * a.n_filter + fpos) * n_mc + c.fin_run] = (a.meas[m] == MCBB_MEAS_MEAN) ? mean : sd;
            }
            if (c.fresh) {
                if (valid) {
                    KP(0, j) = a.io.ic[(long long)row * n_mc + c.run];
#pragma unroll
                    for (int q = 1; q < UM_NSLOT; q++) KP(q, j) = 0.;  // no stale values from the slot's previous run
                }
                am[jj] = a2[jj] = as[jj] = 0.;
            }
        }
        if (all_done) break;
        UM_TICK(1)  // samples, update, results, refill

        // ---- one round: 6 right-hand-side evaluations of the whole batch -----------------------------------------
        // Everything below is written without per-trajectory branches: slots that are idle or still measuring their
        // initial step compute harmlessly on their own columns (their stage slots k2..k7 are zero), so that the G
        // dependent chains of a group sit in one basic block and hide each other's latency.
#pragma unroll 1
        for (int s = 0; s < 6; s++) {
            double cs[6];  // stage coefficients with zeros beyond the current stage
#pragma unroll
            for (int q = 0; q < 6; q++) cs[q] = (q <= s) ? TS_A(s, q) : 0.;
            const double ci0 = (s == 1) ? 1. : 0.;  // initial-dt probe: u0 + dt0 f(u0) in the second evaluation
            // (a) stage input Y -> sin, cos -> the 16 digit bytes of this (row, trajectory) element
#pragma unroll
            for (int g0 = 0; g0 < JBT; g0 += G) {
                double y[G];
#pragma unroll
                for (int i = 0; i < G; i++) {
                    const int j = j0 + g0 + i;
                    const double2* kq = reinterpret_cast<const double2*>(kb + j * UM_NSLOT);
                    const double2 k01 = kq[0], k23 = kq[1], k45 = kq[2];  // (u, k1) (k2, k3) (k4, k5)
                    const double k6 = kq[3].x;  // 128-bit like the others: 64-bit accesses at this row stride conflict 2-way
                    const double c0 = ((runmask >> j) & 1u) ? cs[0] : ci0;
                    double lin = c0 * k01.y;
                    lin = fma(cs[1], k23.x, lin);
                    lin = fma(cs[2], k23.y, lin);
                    lin = fma(cs[3], k45.x, lin);
                    lin = fma(cs[4], k45.y, lin);
                    lin = fma(cs[5], k6, lin);
                    y[i] = fma(ctl[j].hstep, lin, k01.x);
                }
                um_sincos_vec<G>(y, &sv[g0], &cv[g0]);
#pragma unroll
                for (int i = 0; i < G; i++) {
                    uint4 dg;
                    um_fixed52(sv[g0 + i], dg.x, dg.y);
                    um_fixed52(cv[g0 + i], dg.z, dg.w);
                    *reinterpret_cast<uint4*>(sB_row + (j0 + g0 + i) * 2048) = dg;  // rows >= N meet zero columns of A
                }
            }
            UM_TICK(2)  // stage input, sincos, digits
            // (b) D = A * digits on the tensor core.  No CTA barrier: every warp publishes its rows (generic -> async
            //     proxy fence, then a release-arrive on a 4 NW-count mbarrier); the lane whose arrival completes the phase
            //     (pending count 1, profiles/microbench/mbar_probe.cu) issues the MMAs; their completion arrives on
            //     the second mbarrier, the one everybody waits on.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // this thread's TMEM reads of the last stage
            __syncwarp();
            int issuer = 0;
            if (lane == 0) {
                unsigned long long st;
                unsigned pend;
                asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(mbar_full) : "memory");
                asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pend) : "l"(st));
                if (pend == 1u) {
                    issuer = 1;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t}" ::"r"(mbar_full),
                                 "r"(phase)
                                 : "memory");  // acquire the other warps' releases (the phase is complete)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int ks = 0; ks < KSTEPS; ks++)
                        um_mma_i8_ts(tmem_base, tmem_a + 8u * ks, um_desc(sB_addr + ks * 512, 128, 2048), idesc, ks > 0 ? 1u : 0u);
                    um_commit(mbar);
                }
            }
            // Only the issuing warp polls the completion mbarrier (a polling warp burns issue slots: ~10 % of all
            // issued instructions when all four polled); the others sleep on hardware barrier 1 until it joins.
            issuer = __shfl_sync(0xffffffffu, issuer, 0);
            UM_TICK(3)  // publish (+ issue)
            if (issuer) {
                um_mbar_wait(mbar, phase);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
            UM_TICK(4)  // MMA wait
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // (c) row r of P, Q from tensor memory, int64 recombination, k = a1 (c P - s Q) + w
            double kv[JBT];
            double* const kst = kb + 2 + s;
            auto to_k = [&](const int jj, const unsigned* w) {
                const int j = j0 + jj;
                const double P = um_recombine(w[0], w[1], w[2], w[3], w[4], w[5], w[6], degh);
                const double Q = um_recombine(w[8], w[9], w[10], w[11], w[12], w[13], w[14], degh);
                kv[jj] = fma(ctl[j].al1s, cv[jj] * P - sv[jj] * Q, bias);
                kst[j * UM_NSLOT] = ((runmask >> j) & 1u) ? kv[jj] : 0.;  // slots that do not integrate keep zeros
            };
#pragma unroll
            for (int g0 = 0; g0 + 1 < JBT; g0 += 2) {  // two trajectories per tensor-memory load
                unsigned v[32];
                um_tmem_ld32(tmem_row + 16u * (unsigned)(j0 + g0), v);
                to_k(g0, v);
                to_k(g0 + 1, v + 16);
            }
            if (JBT & 1) {
                unsigned v[16];
                um_tmem_ld16(tmem_row + 16u * (unsigned)(j0 + JBT - 1), v);
                to_k(JBT - 1, v);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // TMEM reads done before the next MMA
            UM_TICK(5)  // accumulators -> k
            if (any_init && s < 2) {  // initial step size (OrdinaryDiffEq initdt.jl), only in rounds that start a run
#pragma unroll
                for (int jj = 0; jj < JBT; jj++) {
                    const int j = j0 + jj;
                    if (ctl[j].mode == MODE_INIT) {
                        double p0 = 0., p1 = 0.;
                        if (valid) {
                            const double u0 = KP(0, j);
                            const double sk = o.abstol + fabs(u0) * o.reltol;
                            if (s == 0) {
                                KP(1, j) = kv[jj];
                                const double w0 = u0 / sk, w1 = kv[jj] / sk;
                                p0 = w0 * w0;
                                p1 = w1 * w1;
                            } else {
                                const double vv = (kv[jj] - KP(1, j)) / sk;
                                p0 = vv * vv;
                            }
                        }
                        p0 = um_warp_sum(p0);
                        p1 = um_warp_sum(p1);
                        if (lane == 0) {
                            sm_red0[wq * JB + j] = p0;
                            sm_red1[wq * JB + j] = p1;
                        }
                    }
                }
                __syncthreads();
                if (is_ctl && ctl[lane].mode == MODE_INIT) {
                    UmSlotCtl& c = ctl[lane];
                    double r0 = 0., r1 = 0.;
                    for (int w = 0; w < 4; w++) {
                        r0 += sm_red0[w * JB + lane];
                        r1 += sm_red1[w * JB + lane];
                    }
                    if (o.dt_user > 0) {
                        c.dt = o.dt_user;
                        c.dt0 = 0.;
                    } else if (s == 0) {
                        c.d0 = sqrt(r0 / N);
                        c.d1 = sqrt(r1 / N);
                        const double dt0 = (c.d0 < 1e-5 || c.d1 < 1e-5) ? 1e-6 : 0.01 * (c.d0 / c.d1);
                        c.dt0 = fmin(dt0, o.dtmax);
                    } else {
                        const double d2 = sqrt(r0 / N) / c.dt0;
                        const double dm = fmax(c.d1, d2);
                        const double dt1 = (dm <= 1e-15) ? fmax(1e-6, c.dt0 * 1e-3) : pow(10., -(2. + log10(dm)) / 5.);
                        c.dt = fmin(fmin(100. * c.dt0, dt1), o.dtmax);
                    }
                    c.hstep = c.dt0;
                }
                __syncthreads();
            }
        }

        // ---- error estimate partials and the candidate new state (branch-free; the controller ignores idle slots) ---
        {
            double pe[JBT];
            bool bad[JBT];
#pragma unroll
            for (int jj = 0; jj < JBT; jj++) {
                const int j = j0 + jj;
                const double dts = ctl[j].dts;
                double kk[8];  // u, k1..k7: four 128-bit loads
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2 t2 = reinterpret_cast<const double2*>(kb + j * UM_NSLOT)[q];
                    kk[2 * q] = t2.x;
                    kk[2 * q + 1] = t2.y;
                }
                double lin = TS_BT(0) * kk[1];
                double lu = TS_A(5, 0) * kk[1];
#pragma unroll
                for (int q = 1; q < 7; q++) {
                    lin = fma(TS_BT(q), kk[1 + q], lin);
                    if (q < 6) lu = fma(TS_A(5, q), kk[1 + q], lu);
                }
                const double u0 = kk[0];
                const double u1 = fma(dts, lu, u0);
                const double r = um_div(dts * lin, o.abstol + fmax(fabs(u0), fabs(u1)) * o.reltol);
                pe[jj] = valid ? r * r : 0.;
                un[jj] = u1;
                bad[jj] = valid && (u1 != u1);
            }
#pragma unroll
            for (int off = 1; off < 32; off <<= 1)
#pragma unroll
                for (int jj = 0; jj < JBT; jj++) pe[jj] += __shfl_xor_sync(0xffffffffu, pe[jj], off);
#pragma unroll
            for (int jj = 0; jj < JBT; jj++) {
                const bool anybad = __any_sync(0xffffffffu, bad[jj]);
                if (lane == 0) {
                    sm_red0[wq * JB + j0 + jj] = pe[jj];
                    sm_red1[wq * JB + j0 + jj] = anybad ? 1. : 0.;
                }
            }
        }
        __syncthreads();
    }
#undef KP
    if (TIMING && tid == 0) {
        for (int k = 0; k < 8; k++) atomicAdd(a.timing + k, (unsigned long long)tph[k]);
    }
#undef UM_TICK
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) {
        if (SPLIT) {
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base));
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_a));
        } else {
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(UmCols<JB>::value));
        }
    }
}

}  // namespace mcbb
#include "batched_umma_lw.cuh"
namespace mcbb {

// KL-hist measures of the runs of one chunk from their saved samples: one CTA per run, thread r = oscillator r
__global__ void __launch_bounds__(128) k_umma_kl_post(const UmmaArgs a, const long long n_runs) {
    extern __shared__ __align__(16) unsigned char klp_smem[];  // [kl_bins][128] histogram counters
    const int tid = threadIdx.x, N = a.N, n_t = a.io.n_t;
    const size_t rl = blockIdx.x;
    if ((long long)rl >= n_runs || tid >= N) return;
    const int fpos = a.filter_pos[tid];
    if (fpos < 0) return;
    const double2 ms = *reinterpret_cast<const double2*>(a.stat + (rl * N + tid) * 2);
    double kl = nan("");
    if (a.okflag[rl])
        kl = um_kl_hist(a.smp + rl * (size_t)n_t * N + tid, N, n_t, ms.x, ms.y, a.kl_bins, a.kl_nstd, a.kl_tol,
                        reinterpret_cast<unsigned short*>(klp_smem) + tid);
    for (int m = 0; m < a.n_meas; m++)
        if (a.meas[m] == MCBB_MEAS_KL_HIST) a.io.feat[((long long)m * a.n_filter + fpos) * a.io.n_mc + a.run_begin + (long long)rl] = kl;
}

template <int JB, int MINB, bool TIMING = false, bool KL = false, int NW = 1>
static int launch_umma_one(const UmmaArgs& a, size_t smem, int grid, cudaStream_t st) {
    MCBB_CUDA_OK(cudaFuncSetAttribute(k_solve_batched_umma<JB, MINB, TIMING, KL, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_solve_batched_umma<JB, MINB, TIMING, KL, NW><<<grid, 128 * NW, smem, st>>>(a);
    MCBB_LAUNCHED();
    return 0;
}
template <int JB, int MINB, bool TIMING = false, int NW = 1>
static int launch_umma_spec(const UmmaArgs& a, size_t smem, int grid, cudaStream_t st) {
    if (a.need_kl) return launch_umma_one<JB, MINB, false, true, 1>(a, smem, grid, st);  // KL: one thread per row
    return launch_umma_one<JB, MINB, TIMING, false, NW>(a, smem, grid, st);
}

static size_t umma_smem_bytes(int JB, int N, int n_t) {
    (void)n_t;  // the save times live in the constant bank
    return (size_t)JB * 2048 + sizeof(double) * ((size_t)(N + 1) * (UM_NSLOT * JB + 2) + 8 * (size_t)JB) +
           sizeof(UmSlotCtl) * (size_t)JB + 24 + 128;
}

// Returns handled = true when the problem was run on the tcgen05 path.
int launch_batched_umma(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F,
                        const SolveIO& io, cudaStream_t st, bool* handled) {
    *handled = false;
    const int N = S.n_osc;
    if (S.system != MCBB_SYS_KURAMOTO_NETWORK || hs == nullptr) return 0;
    if (o.alg != MCBB_ALG_TSIT5 || F.n_global > 0 || F.n_comp != 1 || F.matrix_meas) return 0;
    if (F.need_kl && (F.kl_bins > UM_MAX_KL_BINS || F.kl_bins < 3)) return 0;
    if (N < 33 || N > 128) return 0;  // one 128-row MMA tile; small networks stay on the mma.sync kernel
    if (io.n_t > UM_MAX_NT) return 0;  // save times are kept in the constant bank
    if (tune_get("no_imma", 0) || tune_get("no_umma", 0)) return 0;  // test hook (mcbb_test_hook_set), never set in production
    for (int m = 0; m < F.n_meas; m++)
        if (F.meas[m] != MCBB_MEAS_MEAN && F.meas[m] != MCBB_MEAS_STD && F.meas[m] != MCBB_MEAS_KL_HIST) return 0;
    double uw = 0.;  // unit-weight adjacency?
    for (size_t q = 0; q < (size_t)N * N; q++) {
        const double v = hs->mat_a[q];
        if (v != 0.) {
            if (uw == 0.) uw = v;
            if (v != uw) return 0;  // weighted couplings: the FP64 kernel (batched.cu) handles them
        }
    }
    if (uw == 0.) uw = 1.;

    UmmaArgs a = {};
    a.N = N;
    a.KS = (N + 31) / 32;
    a.unit_weight = uw;
    a.inv_n = 1. / (double)N;
    a.ln_qoldinit = std::log(o.qoldinit);
    std::vector<unsigned char> img(UM_A_BYTES, 0);
    std::vector<int> deg(128, 0);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            if (hs->mat_a[i + (size_t)N * j] != 0.) {
                img[(size_t)i * 128 + j] = 1;
                deg[i]++;
            }
    std::vector<double> bias(128, 0.);
    for (int i = 0; i < N; i++) bias[i] = hs->vec_a[i];
    std::vector<int> fpos(128, -1);
    if (F.filter) {
        std::vector<int> fh(F.n_filter);
        MCBB_CUDA_OK(cudaMemcpyAsync(fh.data(), F.filter, sizeof(int) * F.n_filter, cudaMemcpyDeviceToHost, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        for (int ii = 0; ii < F.n_filter; ii++) fpos[fh[ii] - 1] = ii;
    } else {
        for (int i = 0; i < N; i++) fpos[i] = i;
    }
    unsigned char* d_img = (unsigned char*)scratch_get(59, img.size());
    int* d_deg = (int*)scratch_get(38, sizeof(int) * deg.size());
    double* d_bias = (double*)scratch_get(61, sizeof(double) * bias.size());
    int* d_fpos = (int*)scratch_get(62, sizeof(int) * fpos.size());
    if (!d_img || !d_deg || !d_bias || !d_fpos) MCBB_FAIL("solve: device allocation failed");
    MCBB_CUDA_OK(cudaMemcpyAsync(d_img, img.data(), img.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_deg, deg.data(), sizeof(int) * deg.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_bias, bias.data(), sizeof(double) * bias.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_fpos, fpos.data(), sizeof(int) * fpos.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    a.a_img = d_img;
    a.deg = d_deg;
    a.bias = d_bias;
    a.filter_pos = d_fpos;
    for (int q = 0; q < MCBB_N_SCALARS; q++) a.scalars[q] = S.scalars[q];
    a.n_par = S.n_par;
    for (int p = 0; p < MCBB_MAX_PAR; p++) a.par_slot[p] = S.par_slot[p];
    a.o = o;
    a.n_meas = F.n_meas;
    for (int m = 0; m < F.n_meas; m++) a.meas[m] = F.meas[m];
    a.n_filter = F.n_filter;
    a.cyclic = F.cyclic;
    a.io = io;

    // configurations (trajectories per CTA, CTAs per SM): TMEM columns and shared memory both have to fit
    struct Cfg { int jb, per_sm, nw; };
    static const Cfg cfgs[] = {{6, 4, 1}, {4, 4, 1}, {6, 3, 1}, {8, 2, 1}, {4, 3, 1}, {4, 2, 1}, {2, 2, 1}, {2, 1, 1},
                               // five CTAs of four trajectories (split tensor-memory allocation, 96 registers): on request
                               {4, 5, 1},
                               // two threads per oscillator row (experiment switch "umma_nw" = 2)
                               {6, 3, 2}, {6, 4, 2}, {4, 4, 2}, {6, 2, 2}};
    int want_jb = (int)tune_get("umma_jb", 0);
    int want_per = (int)tune_get("umma_per_sm", 0);
    const int want_nw = F.need_kl ? 1 : (int)tune_get("umma_nw", 1);
    const Cfg* pick = nullptr;
    for (const Cfg& c : cfgs) {
        if (c.nw != want_nw) continue;
        if (want_jb && c.jb != want_jb) continue;
        if (want_per && c.per_sm != want_per) continue;
        const size_t need = umma_smem_bytes(c.jb, N, io.n_t);
        if ((need + 1024) * c.per_sm <= (size_t)227 * 1024 && need <= (size_t)226 * 1024) {
            pick = &c;
            break;
        }
    }
    if (!pick) return 0;
    const size_t smem = umma_smem_bytes(pick->jb, N, io.n_t);
    int dev = 0, sms = 148;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    a.n_sms = sms;
    const long long want = (io.n_mc + pick->jb - 1) / pick->jb;
    const int grid = (int)std::min<long long>(want, (long long)sms * pick->per_sm);
    a.need_kl = F.need_kl;
    a.kl_bins = F.kl_bins;
    a.kl_nstd = F.kl_nstd;
    a.kl_tol = F.kl_tol;
    a.smp = nullptr;
    MCBB_CUDA_OK(cudaMemcpyToSymbolAsync(c_um_ts, io.t_save, sizeof(double) * io.n_t, 0, cudaMemcpyDeviceToDevice, st));
    if (a.need_kl) {
        // the ensemble in chunks whose saved samples fit the free device memory (<= 64 GB; test hook "umma_post_mb"), each
        // followed by the binning kernel
        const size_t run_bytes = sizeof(double) * (size_t)io.n_t * (size_t)N;
        // (cudaMemGetInfo costs milliseconds: only ensembles whose samples exceed 2 GB ask how much memory is free)
        size_t cap = (size_t)2 << 30;
        if (run_bytes * (size_t)io.n_mc > cap) {
            size_t free_b = 0, total_b = 0;
            MCBB_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
            cap = std::max<size_t>(cap, std::min<size_t>((size_t)64 << 30, free_b / 2 + scratch_size(39)));
        }
        if (const long long mb = tune_get("umma_post_mb", 0)) cap = (size_t)mb << 20;
        const long long chunk = std::max<long long>(1, std::min<long long>(io.n_mc, (long long)(cap / run_bytes)));
        a.smp = (double*)scratch_get(39, run_bytes * (size_t)chunk);
        a.stat = (double*)scratch_get(106, sizeof(double) * 2 * (size_t)N * (size_t)chunk);
        a.okflag = (int*)scratch_get(107, sizeof(int) * (size_t)chunk);
        if (!a.smp || !a.stat || !a.okflag) MCBB_FAIL("solve: device allocation failed");
        const size_t post_smem = (size_t)a.kl_bins * 256;
        MCBB_CUDA_OK(cudaFuncSetAttribute(k_umma_kl_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem));
        for (long long r0 = 0; r0 < io.n_mc; r0 += chunk) {
            const long long r1 = std::min<long long>(io.n_mc, r0 + chunk);
            a.run_begin = r0;
            a.run_end = r1;
            const unsigned long long c0 = (unsigned long long)r0;
            MCBB_CUDA_OK(cudaMemcpyAsync(io.work_counter, &c0, sizeof(c0), cudaMemcpyHostToDevice, st));
            MCBB_CUDA_OK(cudaStreamSynchronize(st));  // c0 is stack-owned
            const int g = (int)std::min<long long>((r1 - r0 + pick->jb - 1) / pick->jb, (long long)sms * pick->per_sm);
            int rk = 1;
            if (pick->jb == 6 && pick->per_sm == 4) rk = launch_umma_spec<6, 4>(a, smem, g, st);
            else if (pick->jb == 6 && pick->per_sm == 3) rk = launch_umma_spec<6, 3>(a, smem, g, st);
            else if (pick->jb == 4 && pick->per_sm == 4) rk = launch_umma_spec<4, 4>(a, smem, g, st);
            else if (pick->jb == 4 && pick->per_sm == 5) rk = launch_umma_spec<4, 5>(a, smem, g, st);
            else if (pick->jb == 8 && pick->per_sm == 2) rk = launch_umma_spec<8, 2>(a, smem, g, st);
            else if (pick->jb == 4 && pick->per_sm == 3) rk = launch_umma_spec<4, 3>(a, smem, g, st);
            else if (pick->jb == 4 && pick->per_sm == 2) rk = launch_umma_spec<4, 2>(a, smem, g, st);
            else if (pick->jb == 2 && pick->per_sm == 2) rk = launch_umma_spec<2, 2>(a, smem, g, st);
            else if (pick->jb == 2 && pick->per_sm == 1) rk = launch_umma_spec<2, 1>(a, smem, g, st);
            if (rk) return rk;
            k_umma_kl_post<<<(unsigned)(r1 - r0), 128, post_smem, st>>>(a, r1 - r0);
            MCBB_LAUNCHED();
        }
        *handled = true;
        return 0;
    }
    MCBB_CUDA_OK(cudaMemsetAsync(io.work_counter, 0, sizeof(unsigned long long), st));
    int rc = 1;
    if (tune_get("umma_timing", 0) && pick->jb == 4 && (pick->per_sm == 4 || pick->per_sm == 2)) {  // development aid
        a.timing = (unsigned long long*)scratch_get(63, 64);
        if (!a.timing) MCBB_FAIL("solve: device allocation failed");
        MCBB_CUDA_OK(cudaMemsetAsync(a.timing, 0, 64, st));
        rc = pick->per_sm == 4 ? launch_umma_spec<4, 4, true>(a, smem, grid, st) : launch_umma_spec<4, 2, true>(a, smem, grid, st);
        if (rc) return rc;
        unsigned long long h[8];
        MCBB_CUDA_OK(cudaMemcpyAsync(h, a.timing, 64, cudaMemcpyDeviceToHost, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        static const char* nm[8] = {"error+control+barrier", "samples/update", "stage input+sincos", "publish/issue", "MMA wait", "TMEM->k",
                                    "-", "-"};
        double tot = 0;
        for (int k = 0; k < 8; k++) tot += (double)h[k];
        for (int k = 0; k < 8; k++)
            fprintf(stderr, "[umma timing] %-24s %6.2f %%  %10.0f clk per CTA\n", nm[k], 100. * h[k] / tot, (double)h[k] / grid);
        *handled = true;
        return 0;
    }
    // light-warp variant (batched_umma_lw.cuh): the short last row block carried one element per lane
    if (pick->nw == 1 && pick->jb == 6 && pick->per_sm == 4 && !a.need_kl && N > 96 && (N - 96) * 6 <= 32 && tune_get("umma_lw", 0)) {
        MCBB_CUDA_OK(cudaFuncSetAttribute(k_solve_batched_umma_lw<6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a.lw_mode = (int)tune_get("umma_lw", 0);
        k_solve_batched_umma_lw<6, 4><<<grid, 128, smem, st>>>(a);
        MCBB_LAUNCHED();
        *handled = true;
        return 0;
    }
    if (pick->nw == 2) {
        if (pick->jb == 6 && pick->per_sm == 3) rc = launch_umma_spec<6, 3, false, 2>(a, smem, grid, st);
        else if (pick->jb == 6 && pick->per_sm == 4) rc = launch_umma_spec<6, 4, false, 2>(a, smem, grid, st);
        else if (pick->jb == 4 && pick->per_sm == 4) rc = launch_umma_spec<4, 4, false, 2>(a, smem, grid, st);
        else if (pick->jb == 6 && pick->per_sm == 2) rc = launch_umma_spec<6, 2, false, 2>(a, smem, grid, st);
    } else if (pick->jb == 6 && pick->per_sm == 4) rc = launch_umma_spec<6, 4>(a, smem, grid, st);
    else if (pick->jb == 6 && pick->per_sm == 3) rc = launch_umma_spec<6, 3>(a, smem, grid, st);
    else if (pick->jb == 4 && pick->per_sm == 4) rc = launch_umma_spec<4, 4>(a, smem, grid, st);
    else if (pick->jb == 4 && pick->per_sm == 5) rc = launch_umma_spec<4, 5>(a, smem, grid, st);
    else if (pick->jb == 8 && pick->per_sm == 2) rc = launch_umma_spec<8, 2>(a, smem, grid, st);
    else if (pick->jb == 4 && pick->per_sm == 3) rc = launch_umma_spec<4, 3>(a, smem, grid, st);
    else if (pick->jb == 4 && pick->per_sm == 2) rc = launch_umma_spec<4, 2>(a, smem, grid, st);
    else if (pick->jb == 2 && pick->per_sm == 2) rc = launch_umma_spec<2, 2>(a, smem, grid, st);
    else if (pick->jb == 2 && pick->per_sm == 1) rc = launch_umma_spec<2, 1>(a, smem, grid, st);
    if (rc) return rc;
    *handled = true;
    return 0;
}

}  // namespace mcbb
