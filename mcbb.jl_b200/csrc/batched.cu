// batched.cu — block-cooperative ensemble Tsit5 for LARGE phase-oscillator networks (N ~ 24..160):
// kuramoto_network (src/systems.jl:119-130) and non_local_kuramoto_ring (src/systems.jl:230-239),
// with the eval_ode_run statistics (src/eval_mc_prob.jl:191-198) fused into the tail: mean / std as running moments,
// empirical_1D_KL_divergence_hist (the third default measure, eval_mc_prob.jl:312-330) from the saved samples of a chunk
// of runs, kept in a device scratch and binned after the chunk is integrated (k_batched_kl_post).
//
// Why a second kernel: one trajectory per lane needs 9 N doubles of state per lane, which no longer fits on
// chip for N = 100.  Here a CTA integrates a BATCH of B trajectories in lockstep rounds.  Every trajectory
// shares the coupling matrix M, so one right-hand-side evaluation of the whole batch is
//
//     [P | Q] = M . [sin(Y) | cos(Y)]            (N x N) . (N x 2B)   — an FP64 GEMM on the FMA pipe
//     k_i     = a1 (cos y_i P_i - sin y_i Q_i) + a2 (sin y_i P_i + cos y_i Q_i) + bias_i
//
// (sum_j M_ij sin(y_j - y_i) factored: N sincos + 2 N^2 FMA instead of N^2 sin.)  Each thread owns a
// TI x TB = 5 x 4 block of (oscillator, trajectory) elements: it keeps that block's GEMM accumulators in
// registers and ALL Runge-Kutta stage vectors of its elements in thread-private shared memory, so the only
// data shared between threads are sin(Y), cos(Y).  Zero columns of M are skipped per 5-row group (the
// reference skips A_ij == 0, systems.jl:124).  Step-size control is per trajectory (per-lane error norm,
// accept / reject, PI controller), trajectories that finish are replaced from a global queue.
#include <algorithm>
#include <vector>

#include "lane_core.cuh"
#include "umma_common.cuh"  // um_kl_hist

namespace mcbb {

constexpr int NSLOT = 8;  // u, k1..k7

struct BatchArgs {
    int N, IG, BG, B, nthreads;
    int n_cols;            // length of the per-row-group column lists (padded, +2 sentinel entries)
    const int* jlist;      // [IG][n_cols+2]            (weighted path, global)
    const double* aval;    // [IG][n_cols+2][TI]        (weighted path, global)
    const unsigned short* ent;  // [IG][n_cols+2] column | row-mask << 8   (unit-weight path, staged in smem)
    double unit_weight;    // common value of the non-zero couplings (unit-weight path)
    const double* bias;    // [IG*TI]
    const int* filter_pos;  // [IG*TI]: output position of oscillator i, or -1
    int system;
    double scalars[MCBB_N_SCALARS];
    int n_par;
    int par_slot[MCBB_MAX_PAR];
    SolverDev o;
    int n_meas, meas[MCBB_MAX_MEAS], n_filter, cyclic;
    SolveIO io;
    double* acc;  // [grid][3][NE][nthreads]  mean, M2, cyclic offset (L2-resident scratch)
    int need_kl, kl_bins;
    double kl_nstd, kl_tol;
    double* smp;  // [runs of the chunk][n_t][IG*TI] saved samples (KL-hist measures only; binned by k_batched_kl_post)
    double* stat;   // [runs of the chunk][IG*TI][2] mean, std of every run as it finished
    int* okflag;    // [runs of the chunk]
    long long run_begin, run_end;  // the runs this launch integrates (one chunk of the ensemble when samples are kept)
};

// sin and cos together, |x| < ~1e6: Cody-Waite reduction by pi/2 with two FMAs + fdlibm kernel polynomials.
// Straight-line code (no slow path) so that several independent evaluations interleave on the FP64 pipe.
__device__ __forceinline__ void sincos_fast(const double x, double* s, double* c) {
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

template <int TB>
__device__ __forceinline__ void load_panel(const double* p, double* out) {
    static_assert(TB % 2 == 0, "TB must be even (16-byte shared-memory loads)");
#pragma unroll
    for (int q = 0; q < TB; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(p + q);
        out[q] = v.x;
        out[q + 1] = v.y;
    }
}

struct SlotCtl {  // per-trajectory control block (shared memory, one per batch slot)
    long long run, iter, nrej, natt;
    double t, dt, dts, qold, al1, al2, d0, d1, dt0, t_old, dts_old;
    int mode, isave, ret, clipped, accept, fresh, finished, s0, s1;
};

template <int TI, int TB, bool UNIT>
__global__ void __launch_bounds__(256, 1) k_solve_batched(const BatchArgs a) {
    constexpr int NE = TI * TB;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, nt = a.nthreads;
    const int ig = tid / a.BG, bg = tid - ig * a.BG;
    const int B = a.B, N = a.N;
    const int NP = a.IG * TI;
    const int ncp = a.n_cols + 2;
    double* sm_k = (double*)smem;                                    // [NSLOT][NE][nt]
    double* sm_S = sm_k + (size_t)NSLOT * NE * nt;                   // [NP][B]
    double* sm_C = sm_S + (size_t)NP * B;                            // [NP][B]
    double* sm_red0 = sm_C + (size_t)NP * B;                         // [IG][B]
    double* sm_red1 = sm_red0 + (size_t)a.IG * B;                    // [IG][B]
    double* sm_ts = sm_red1 + (size_t)a.IG * B;                      // [n_t]
    SlotCtl* ctl = (SlotCtl*)(sm_ts + ((a.io.n_t + 1) & ~1));        // [B]
    unsigned short* sm_ent = (unsigned short*)(ctl + B);             // [IG][ncp]  (unit-weight path)
    const SolverDev& o = a.o;
    const int n_t = a.io.n_t;
    const long long n_mc = a.io.n_mc;
    double* acc = a.acc + (size_t)blockIdx.x * 3 * NE * nt + tid;

#define KS(slot, e) sm_k[((size_t)(slot)*NE + (e)) * nt + tid]
#define ACCS(k, e) acc[((size_t)(k)*NE + (e)) * nt]

    for (int i = tid; i < n_t; i += nt) sm_ts[i] = a.io.t_save[i];
    if (UNIT)
        for (int i = tid; i < a.IG * ncp; i += nt) sm_ent[i] = a.ent[i];
    if (tid < B) {
        ctl[tid].mode = MODE_FETCH;
        ctl[tid].fresh = 0;
        ctl[tid].finished = 0;
        ctl[tid].accept = 0;
    }
    double bias_r[TI];
#pragma unroll
    for (int r = 0; r < TI; r++) bias_r[r] = a.bias[ig * TI + r];
    __syncthreads();

    // begin a step attempt for slot c (maxiters / tstop clipping / dtmin), as in lane_core.cuh
    auto begin_attempt = [&](SlotCtl& c) -> bool {
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
        return true;
    };

    for (;;) {
        // ---- refill finished slots from the global queue ------------------------------------------------
        if (tid < B) {
            SlotCtl& c = ctl[tid];
            c.fresh = 0;
            if (c.mode == MODE_FETCH) {
                const long long run = (long long)atomicAdd(a.io.work_counter, 1ULL);
                if (run >= a.run_end) {
                    c.mode = MODE_DONE;
                } else {
                    double sc[MCBB_N_SCALARS];
                    for (int q = 0; q < MCBB_N_SCALARS; q++) sc[q] = a.scalars[q];
                    for (int p = 0; p < a.n_par; p++) sc[a.par_slot[p]] = a.io.par[(long long)p * n_mc + run];
                    if (a.system == MCBB_SYS_KURAMOTO_NETWORK) {
                        c.al1 = sc[0] / (double)N;  // K/N, systems.jl:128
                        c.al2 = 0.;
                    } else {  // non_local_kuramoto_ring: omega_0 is carried by the bias table
                        double sa, ca;
                        sincos(sc[1], &sa, &ca);
                        c.al1 = ca / sc[2];
                        c.al2 = -sa / sc[2];
                    }
                    if (UNIT) {
                        c.al1 *= a.unit_weight;
                        c.al2 *= a.unit_weight;
                    }
                    c.run = run;
                    c.t = o.t0;
                    c.iter = 0;
                    c.nrej = 0;
                    c.natt = 0;
                    c.qold = o.qoldinit;
                    c.isave = 0;
                    c.ret = 0;
                    c.mode = MODE_INIT;
                    c.fresh = 1;
                }
            }
        }
        const bool all_done = __syncthreads_and(tid >= B || ctl[tid].mode == MODE_DONE);
        if (all_done) break;
#pragma unroll 1
        for (int q = 0; q < TB; q++) {
            const int b = bg * TB + q;
            if (ctl[b].fresh) {
                const long long run = ctl[b].run;
#pragma unroll
                for (int r = 0; r < TI; r++) {
                    const int i = ig * TI + r, e = r * TB + q;
                    KS(0, e) = (i < N) ? a.io.ic[(long long)i * n_mc + run] : 0.;
                    ACCS(0, e) = 0.;
                    ACCS(1, e) = 0.;
                    ACCS(2, e) = 0.;
                }
            }
        }

        // ---- one round: 6 right-hand-side evaluations of the whole batch -----------------------------------
#pragma unroll 1
        for (int s = 0; s < 6; s++) {
            // (a) stage input Y -> sin, cos into the shared S / C panels
#pragma unroll 1
            for (int q = 0; q < TB; q++) {
                const int b = bg * TB + q;
                const int mode = ctl[b].mode;
                const double h = (mode == MODE_RUN) ? ctl[b].dts : ctl[b].dt0;
                double y[TI];
#pragma unroll
                for (int r = 0; r < TI; r++) y[r] = KS(0, r * TB + q);
                if (mode == MODE_RUN) {
                    double lin[TI];
                    const double a0 = TS_A(s, 0);
#pragma unroll
                    for (int r = 0; r < TI; r++) lin[r] = a0 * KS(1, r * TB + q);
                    for (int j = 1; j <= s; j++) {
                        const double aj = TS_A(s, j);
#pragma unroll
                        for (int r = 0; r < TI; r++) lin[r] = fma(aj, KS(1 + j, r * TB + q), lin[r]);
                    }
#pragma unroll
                    for (int r = 0; r < TI; r++) y[r] = fma(h, lin[r], y[r]);
                } else if (mode == MODE_INIT && s == 1) {
#pragma unroll
                    for (int r = 0; r < TI; r++) y[r] = fma(h, KS(1, r * TB + q), y[r]);
                }
#pragma unroll
                for (int r = 0; r < TI; r++) {
                    double sv, cv;
                    sincos_fast(y[r], &sv, &cv);
                    sm_S[(size_t)(ig * TI + r) * B + b] = sv;
                    sm_C[(size_t)(ig * TI + r) * B + b] = cv;
                }
            }
            __syncthreads();
            // (b) [P | Q] = M . [S | C] for this thread's TI x TB block
            double P[TI][TB], Q[TI][TB];
#pragma unroll
            for (int r = 0; r < TI; r++)
#pragma unroll
                for (int q = 0; q < TB; q++) P[r][q] = Q[r][q] = 0.;
            {
                const double* Sb = sm_S + bg * TB;
                const double* Cb = sm_C + bg * TB;
                if (UNIT) {
                    // adjacency as (column, row-mask) entries in shared memory; S / C rows one column ahead
                    const unsigned short* en = sm_ent + ig * ncp;
                    unsigned e1 = en[1];
                    int j = en[0] & 0xff;
                    unsigned mask = en[0] >> 8;
                    double sq[TB], cq[TB];
                    load_panel<TB>(Sb + (size_t)j * B, sq);
                    load_panel<TB>(Cb + (size_t)j * B, cq);
#pragma unroll 2
                    for (int c = 0; c < a.n_cols; c++) {
                        const unsigned e2 = en[c + 2];
                        const int jn = e1 & 0xff;
                        double sn[TB], cn[TB];
                        load_panel<TB>(Sb + (size_t)jn * B, sn);
                        load_panel<TB>(Cb + (size_t)jn * B, cn);
#pragma unroll
                        for (int r = 0; r < TI; r++) {
                            const double mr = ((mask >> r) & 1u) ? 1.0 : 0.0;
#pragma unroll
                            for (int q = 0; q < TB; q++) {
                                P[r][q] = fma(mr, sq[q], P[r][q]);
                                Q[r][q] = fma(mr, cq[q], Q[r][q]);
                            }
                        }
                        mask = e1 >> 8;
                        e1 = e2;
#pragma unroll
                        for (int q = 0; q < TB; q++) {
                            sq[q] = sn[q];
                            cq[q] = cn[q];
                        }
                    }
                } else {
                    const int* jl = a.jlist + (size_t)ig * ncp;
                    const double* av = a.aval + (size_t)ig * ncp * TI;
                    int jn = __ldg(jl);
                    double an[TI];
#pragma unroll
                    for (int r = 0; r < TI; r++) an[r] = __ldg(av + r);
#pragma unroll 2
                    for (int c = 0; c < a.n_cols; c++) {
                        const int j = jn;
                        double ar[TI];
#pragma unroll
                        for (int r = 0; r < TI; r++) ar[r] = an[r];
                        jn = __ldg(jl + c + 1);  // sentinel entries make c+1 always valid
#pragma unroll
                        for (int r = 0; r < TI; r++) an[r] = __ldg(av + (size_t)(c + 1) * TI + r);
                        double sq[TB], cq[TB];
                        load_panel<TB>(Sb + (size_t)j * B, sq);
                        load_panel<TB>(Cb + (size_t)j * B, cq);
#pragma unroll
                        for (int r = 0; r < TI; r++)
#pragma unroll
                            for (int q = 0; q < TB; q++) {
                                P[r][q] = fma(ar[r], sq[q], P[r][q]);
                                Q[r][q] = fma(ar[r], cq[q], Q[r][q]);
                            }
                    }
                }
            }
            // (c) k = a1 (c P - s Q) + a2 (s P + c Q) + bias  -> thread-private stage slot
#pragma unroll
            for (int q = 0; q < TB; q++) {
                const int b = bg * TB + q;
                const int mode = ctl[b].mode;
                const double al1 = ctl[b].al1, al2 = ctl[b].al2;
                double p0 = 0., p1 = 0.;
#pragma unroll
                for (int r = 0; r < TI; r++) {
                    const int e = r * TB + q, i = ig * TI + r;
                    const double sv = sm_S[(size_t)i * B + b], cv = sm_C[(size_t)i * B + b];
                    double kv = fma(al1, cv * P[r][q] - sv * Q[r][q], bias_r[r]);
                    kv = fma(al2, sv * P[r][q] + cv * Q[r][q], kv);
                    if (i >= N) kv = 0.;
                    if (mode == MODE_RUN) {
                        KS(2 + s, e) = kv;
                    } else if (mode == MODE_INIT && s < 2) {  // OrdinaryDiffEq initdt.jl norms
                        const double u0 = KS(0, e);
                        const double sk = o.abstol + fabs(u0) * o.reltol;
                        if (s == 0) {
                            KS(1, e) = kv;
                            const double v0 = u0 / sk, v1 = kv / sk;
                            p0 = fma(v0, v0, p0);
                            p1 = fma(v1, v1, p1);
                        } else {
                            const double v = (kv - KS(1, e)) / sk;
                            p0 = fma(v, v, p0);
                        }
                    }
                }
                if (mode == MODE_INIT && s < 2) {
                    sm_red0[ig * B + b] = p0;
                    sm_red1[ig * B + b] = p1;
                }
            }
            __syncthreads();
            if (s < 2 && tid < B && ctl[tid].mode == MODE_INIT) {
                SlotCtl& c = ctl[tid];
                double r0 = 0., r1 = 0.;
                for (int g = 0; g < a.IG; g++) {  // fixed order: deterministic
                    r0 += sm_red0[g * B + tid];
                    r1 += sm_red1[g * B + tid];
                }
                if (o.dt_user > 0) {
                    c.dt = o.dt_user;
                    c.dt0 = 0.;
                } else if (s == 0) {
                    c.d0 = sqrt(r0 / N);
                    c.d1 = sqrt(r1 / N);
                    double dt0 = (c.d0 < 1e-5 || c.d1 < 1e-5) ? 1e-6 : 0.01 * (c.d0 / c.d1);
                    c.dt0 = fmin(dt0, o.dtmax);
                } else {
                    const double d2 = sqrt(r0 / N) / c.dt0;
                    const double dm = fmax(c.d1, d2);
                    const double dt1 = (dm <= 1e-15) ? fmax(1e-6, c.dt0 * 1e-3) : pow(10., -(2. + log10(dm)) / 5.);
                    c.dt = fmin(fmin(100. * c.dt0, dt1), o.dtmax);
                }
            }
            if (s < 2) __syncthreads();
        }

        // ---- error estimate partials (running slots) -----------------------------------------------------------
#pragma unroll 1
        for (int q = 0; q < TB; q++) {
            const int b = bg * TB + q;
            if (ctl[b].mode != MODE_RUN) continue;
            const double dts = ctl[b].dts;
            double p = 0.;
#pragma unroll
            for (int r = 0; r < TI; r++) {
                const int e = r * TB + q;
                double lin = TS_BT(0) * KS(1, e);
                double lu = TS_A(5, 0) * KS(1, e);
#pragma unroll
                for (int j = 1; j < 7; j++) {
                    lin = fma(TS_BT(j), KS(1 + j, e), lin);
                    if (j < 6) lu = fma(TS_A(5, j), KS(1 + j, e), lu);
                }
                const double u0 = KS(0, e);
                const double un = fma(dts, lu, u0);
                const double rr = (dts * lin) / (o.abstol + fmax(fabs(u0), fabs(un)) * o.reltol);
                p = fma(rr, rr, p);
            }
            sm_red0[ig * B + b] = p;
        }
        __syncthreads();

        // ---- per-trajectory control: PI controller, accept / reject, save window ---------------------------------
        if (tid < B) {
            SlotCtl& c = ctl[tid];
            c.accept = 0;
            c.finished = 0;
            if (c.mode == MODE_INIT) {
                c.mode = MODE_RUN;
                if (!begin_attempt(c)) c.finished = 1;
            } else if (c.mode == MODE_RUN) {
                c.natt++;
                double e2 = 0.;
                for (int g = 0; g < a.IG; g++) e2 += sm_red0[g * B + tid];
                const double EEst = sqrt(e2 / N);
                double q, q11 = 0.;
                if (EEst == 0.) {
                    q = 1. / o.qmax;
                } else {
                    q11 = pow(EEst, o.beta1);
                    q = q11 / pow(c.qold, o.beta2);
                    q = fmax(1. / o.qmax, fmin(1. / o.qmin, q / o.gamma));
                }
                if (!(EEst > 1.)) {
                    c.qold = fmax(EEst, o.qoldinit);
                    const double tnew = c.clipped ? o.t1 : c.t + c.dts;
                    int k1 = c.isave;
                    while (k1 < n_t && sm_ts[k1] <= tnew) k1++;
                    c.s0 = c.isave;
                    c.s1 = k1;
                    c.isave = k1;
                    c.t_old = c.t;
                    c.dts_old = c.dts;
                    c.t = tnew;
                    c.dt = fmin(o.dtmax, fmax(o.dtmin, c.dts / q));
                    c.accept = 1;
                } else {
                    c.nrej++;
                    c.dt = c.dts / fmin(1. / o.qmin, q11 / o.gamma);
                }
            }
        }
        __syncthreads();

        // ---- accepted steps: fused featurizer on the saved samples, then u <- u_new, k1 <- k7 ----------------
#pragma unroll 1
        for (int q = 0; q < TB; q++) {
            const int b = bg * TB + q;
            if (!ctl[b].accept) continue;
            const double dts = ctl[b].dts_old, t_old = ctl[b].t_old;
            const int ks0 = ctl[b].s0 < 1 ? 1 : ctl[b].s0, ks1 = ctl[b].s1;  // sample 0 is dropped (eval_mc_prob.jl:152)
            if (ks0 < ks1) {
                double* const smp_b = a.need_kl ? a.smp + (size_t)(ctl[b].run - a.run_begin) * n_t * NP : nullptr;
                // Welford state in registers across the samples of this step: one L2 round trip per accepted step
                double am[TI], a2[TI], as[TI];
#pragma unroll
                for (int r = 0; r < TI; r++) {
                    const int e = r * TB + q;
                    am[r] = ACCS(0, e);
                    a2[r] = ACCS(1, e);
                    as[r] = ACCS(2, e);
                }
                for (int k = ks0; k < ks1; k++) {
                    const double ts = sm_ts[k];
                    double bth[7];
                    tsit5_btheta((ts - t_old) / dts, bth);
                    const double rk = 1. / (double)k;
#pragma unroll
                    for (int r = 0; r < TI; r++) {
                        const int e = r * TB + q;
                        double lin = KS(1, e) * bth[0];
#pragma unroll
                        for (int j = 1; j < 7; j++) lin = fma(KS(1 + j, e), bth[j], lin);
                        double x = fma(dts, lin, KS(0, e));
                        if (a.cyclic) {  // _cyclic_setback!, eval_mc_prob.jl:286-296
                            if (k == 1) {
                                const double twopi = 6.283185307179586, pi = 3.141592653589793;
                                double v = julia_mod_pos(x, twopi);
                                if (v > pi) v -= twopi;
                                as[r] = x - v;
                                x = v;
                            } else {
                                x = x - as[r];
                            }
                        }
                        if (a.need_kl) smp_b[(size_t)k * NP + ig * TI + r] = x;
                        const double delta = x - am[r];
                        if (fabs(delta) <= 1.7976931348623157e308) {
                            const double mnew = fma(delta, rk, am[r]);
                            a2[r] = fma(delta, x - mnew, a2[r]);
                            am[r] = mnew;
                        } else {
                            am[r] = nonfinite_mean(am[r], x);
                            a2[r] = nan("");
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < TI; r++) {
                    const int e = r * TB + q;
                    ACCS(0, e) = am[r];
                    ACCS(1, e) = a2[r];
                    ACCS(2, e) = as[r];
                }
            }
            bool bad = false;
#pragma unroll
            for (int r = 0; r < TI; r++) {
                const int e = r * TB + q;
                double lu = TS_A(5, 0) * KS(1, e);
#pragma unroll
                for (int j = 1; j < 6; j++) lu = fma(TS_A(5, j), KS(1 + j, e), lu);
                const double un = fma(dts, lu, KS(0, e));
                bad |= (un != un);
                KS(0, e) = un;
                KS(1, e) = KS(7, e);  // FSAL
            }
            if (bad) ctl[b].ret = MCBB_RET_UNSTABLE;  // benign race: every writer stores the same value
        }
        __syncthreads();
        if (tid < B) {
            SlotCtl& c = ctl[tid];
            if (c.mode == MODE_RUN && !c.finished) {
                if (c.ret != 0 || (c.accept && !(c.t < o.t1))) {
                    c.finished = 1;
                } else if (!begin_attempt(c)) {
                    c.finished = 1;
                }
            }
        }
        __syncthreads();
        // ---- finished trajectories: write mean / std of the owned oscillators -----------------------------------
#pragma unroll 1
        for (int q = 0; q < TB; q++) {
            const int b = bg * TB + q;
            if (!ctl[b].finished) continue;
            const long long run = ctl[b].run;
            const bool ok = (ctl[b].ret == 0) && ctl[b].isave >= n_t;
#pragma unroll
            for (int r = 0; r < TI; r++) {
                const int i = ig * TI + r, e = r * TB + q;
                if (i >= N) continue;
                if (a.need_kl && i == 0) a.okflag[(size_t)(run - a.run_begin)] = ok ? 1 : 0;
                const int pos = a.filter_pos[i];
                if (pos < 0) continue;
                const double mean = ok ? ACCS(0, e) : nan("");
                const double m2 = ACCS(1, e);
                const double sd = ok ? sqrt((m2 < 0. ? 0. : m2) / (double)(n_t - 2)) : nan("");
                if (a.need_kl) {  // the KL measures are binned after the solve: the whole block would wait here otherwise
                    const size_t rl = (size_t)(run - a.run_begin);
                    *reinterpret_cast<double2*>(a.stat + (rl * NP + i) * 2) = make_double2(mean, sd);
                }
                for (int m = 0; m < a.n_meas; m++)
                    if (a.meas[m] != MCBB_MEAS_KL_HIST)
                        a.io.feat[((long long)m * a.n_filter + pos) * n_mc + run] = (a.meas[m] == MCBB_MEAS_MEAN) ? mean : sd;
            }
        }
        __syncthreads();
        if (tid < B && ctl[tid].finished) {
            SlotCtl& c = ctl[tid];
            if (a.io.retcode) a.io.retcode[c.run] = c.ret;
            if (a.io.nsteps) {
                a.io.nsteps[c.run] = c.natt;
                a.io.nsteps[n_mc + c.run] = c.nrej;
            }
            c.mode = MODE_FETCH;
            c.finished = 0;
        }
        // the next iteration starts with a barrier-protected refill
    }
#undef KS
#undef ACCS
}

// KL-hist measures of the runs of one chunk from their saved samples: one CTA per run, thread i = oscillator i
__global__ void __launch_bounds__(192) k_batched_kl_post(const BatchArgs a, const int NP, const long long n_runs) {
    const int i = threadIdx.x, n_t = a.io.n_t;
    const size_t rl = blockIdx.x;
    if ((long long)rl >= n_runs || i >= a.N) return;
    const int pos = a.filter_pos[i];
    if (pos < 0) return;
    const double2 ms = *reinterpret_cast<const double2*>(a.stat + (rl * NP + i) * 2);
    double kl = nan("");
    if (a.okflag[rl]) kl = um_kl_hist(a.smp + rl * (size_t)n_t * NP + i, NP, n_t, ms.x, ms.y, a.kl_bins, a.kl_nstd, a.kl_tol, nullptr);
    for (int m = 0; m < a.n_meas; m++)
        if (a.meas[m] == MCBB_MEAS_KL_HIST) a.io.feat[((long long)m * a.n_filter + pos) * a.io.n_mc + a.run_begin + (long long)rl] = kl;
}

// ---- host side ---------------------------------------------------------------------------------------------
constexpr int TI_ = 5;

template <int TB, bool UNIT>
static int launch_variant(BatchArgs& a, size_t smem, int grid, cudaStream_t st) {
    MCBB_CUDA_OK(cudaFuncSetAttribute(k_solve_batched<TI_, TB, UNIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    k_solve_batched<TI_, TB, UNIT><<<grid, a.nthreads, smem, st>>>(a);
    MCBB_LAUNCHED();
    return 0;
}

int launch_batched_kuramoto(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F,
                            const SolveIO& io, cudaStream_t st, bool* handled) {
    *handled = false;
    const int N = S.n_osc;
    constexpr int TI = TI_;
    if (S.system != MCBB_SYS_KURAMOTO_NETWORK && S.system != MCBB_SYS_NONLOCAL_KURAMOTO_RING) return 0;
    if (o.alg != MCBB_ALG_TSIT5 || F.n_global > 0 || F.n_comp != 1) return 0;
    if (F.need_kl && (F.kl_bins > UM_MAX_KL_BINS || F.kl_bins < 3 || tune_get("no_batched_kl", 0))) return 0;
    if (F.matrix_meas != MCBB_MAT_NONE) return 0;  // no matrix measure here: the lane kernels compute it
    if (N < 24 || N > 160 || hs == nullptr) return 0;
    for (int m = 0; m < F.n_meas; m++)
        if (F.meas[m] != MCBB_MEAS_MEAN && F.meas[m] != MCBB_MEAS_STD && F.meas[m] != MCBB_MEAS_KL_HIST) return 0;
    if (S.system == MCBB_SYS_NONLOCAL_KURAMOTO_RING)
        for (int p = 0; p < S.n_par; p++)
            if (S.par_slot[p] != 1) return 0;  // only phase_delay sweeps are batched (omega_0, fak enter the tables)

    BatchArgs a = {};
    a.N = N;
    a.IG = (N + TI - 1) / TI;
    const int NP = a.IG * TI;
    // dense coupling matrix M (row i, column j) on the host
    std::vector<double> M((size_t)NP * N, 0.);
    if (S.system == MCBB_SYS_KURAMOTO_NETWORK) {
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) M[(size_t)i * N + j] = hs->mat_a[i + (size_t)N * j];
    } else {
        for (int k = 0; k < N; k++)
            for (int j = 0; j < N; j++) M[(size_t)k * N + j] = hs->mat_a[(k - j) + N - 1];
    }
    // unit-weight adjacency: every non-zero coupling has the same value (the usual MCBB network)
    bool unit = true;
    double uw = 0.;
    for (size_t q = 0; q < M.size() && unit; q++)
        if (M[q] != 0.) {
            if (uw == 0.) uw = M[q];
            unit = (M[q] == uw);
        }
    if (uw == 0.) {
        unit = true;
        uw = 1.;
    }
    // per 5-row group: the columns where any row is non-zero (ascending j = the reference's summation order)
    std::vector<std::vector<int>> cols(a.IG);
    int n_cols = 1;
    for (int g = 0; g < a.IG; g++) {
        for (int j = 0; j < N; j++) {
            bool nz = false;
            for (int r = 0; r < TI; r++) nz |= (M[(size_t)(g * TI + r) * N + j] != 0.);
            if (nz) cols[g].push_back(j);
        }
        n_cols = std::max(n_cols, (int)cols[g].size());
    }
    n_cols = (n_cols + 1) & ~1;  // the column loop is unrolled by two
    a.n_cols = n_cols;
    const int ncp = n_cols + 2;  // two sentinel entries so the software pipeline never reads out of bounds
    std::vector<int> jl((size_t)a.IG * ncp, 0);
    std::vector<double> av((size_t)a.IG * ncp * TI, 0.);
    std::vector<unsigned short> ent((size_t)a.IG * ncp, 0);
    for (int g = 0; g < a.IG; g++)
        for (size_t c = 0; c < cols[g].size(); c++) {
            const int j = cols[g][c];
            jl[(size_t)g * ncp + c] = j;
            unsigned mask = 0;
            for (int r = 0; r < TI; r++) {
                const double v = M[(size_t)(g * TI + r) * N + j];
                av[((size_t)g * ncp + c) * TI + r] = v;
                if (v != 0.) mask |= 1u << r;
            }
            ent[(size_t)g * ncp + c] = (unsigned short)(j | (mask << 8));
        }
    std::vector<double> bias(NP, 0.);
    for (int i = 0; i < N; i++) bias[i] = (S.system == MCBB_SYS_KURAMOTO_NETWORK) ? hs->vec_a[i] : hs->scalars[0];
    std::vector<int> fpos(NP, -1);
    if (F.filter) {
        std::vector<int> fh(F.n_filter);
        MCBB_CUDA_OK(cudaMemcpyAsync(fh.data(), F.filter, sizeof(int) * F.n_filter, cudaMemcpyDeviceToHost, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        for (int ii = 0; ii < F.n_filter; ii++) fpos[fh[ii] - 1] = ii;
    } else {
        for (int i = 0; i < N; i++) fpos[i] = i;
    }
    int* d_jl = (int*)scratch_get(50, sizeof(int) * jl.size());
    double* d_av = (double*)scratch_get(51, sizeof(double) * av.size());
    double* d_bias = (double*)scratch_get(52, sizeof(double) * bias.size());
    int* d_fpos = (int*)scratch_get(53, sizeof(int) * fpos.size());
    unsigned short* d_ent = (unsigned short*)scratch_get(55, sizeof(unsigned short) * ent.size());
    if (!d_jl || !d_av || !d_bias || !d_fpos || !d_ent) MCBB_FAIL("solve: device allocation failed");
    MCBB_CUDA_OK(cudaMemcpyAsync(d_jl, jl.data(), sizeof(int) * jl.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_av, av.data(), sizeof(double) * av.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_bias, bias.data(), sizeof(double) * bias.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_fpos, fpos.data(), sizeof(int) * fpos.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_ent, ent.data(), sizeof(unsigned short) * ent.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    a.jlist = d_jl;
    a.aval = d_av;
    a.ent = d_ent;
    a.unit_weight = uw;
    a.bias = d_bias;
    a.filter_pos = d_fpos;
    a.system = S.system;
    for (int q = 0; q < MCBB_N_SCALARS; q++) a.scalars[q] = S.scalars[q];
    a.n_par = S.n_par;
    for (int p = 0; p < MCBB_MAX_PAR; p++) a.par_slot[p] = S.par_slot[p];
    a.o = o;
    a.n_meas = F.n_meas;
    for (int m = 0; m < F.n_meas; m++) a.meas[m] = F.meas[m];
    a.n_filter = F.n_filter;
    a.cyclic = F.cyclic;
    a.io = io;

    // register block width: 2 trajectories per thread doubles the resident warps (latency hiding) at the same
    // shared-memory footprint; MCBB_BATCH_TB=4 selects the wider block for experiments
    int TB = 2;
    if (const long long e = tune_get("batch_tb", 0)) TB = (e == 4) ? 4 : 2;
    // batch size: as many TB-trajectory groups as shared memory holds (<= 256 threads)
    const size_t cap = 227 * 1024;
    int BG = 0;
    size_t smem = 0;
    for (int cand = 256 / a.IG; cand >= 1; cand--) {
        const int nt = a.IG * cand, B = cand * TB;
        const size_t need = sizeof(double) * ((size_t)NSLOT * TI * TB * nt + 2 * (size_t)NP * B + 2 * (size_t)a.IG * B +
                                              (size_t)((io.n_t + 1) & ~1)) +
                            sizeof(SlotCtl) * (size_t)B + sizeof(unsigned short) * (size_t)a.IG * ncp + 32;
        if (need <= cap) {
            BG = cand;
            smem = need;
            break;
        }
    }
    if (BG == 0) return 0;  // not handled: fall through to the per-lane kernel's own diagnostics
    a.BG = BG;
    a.B = BG * TB;
    a.nthreads = a.IG * BG;
    int dev = 0, sms = 148;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (io.n_mc + a.B - 1) / a.B;
    const int grid = (int)std::min<long long>(want, sms);
    a.acc = (double*)scratch_get(54, sizeof(double) * (size_t)grid * 3 * TI * TB * a.nthreads);
    if (!a.acc) MCBB_FAIL("solve: device allocation failed");
    a.need_kl = F.need_kl;
    a.kl_bins = F.kl_bins;
    a.kl_nstd = F.kl_nstd;
    a.kl_tol = F.kl_tol;
    a.smp = nullptr;
    // KL-hist: the ensemble in chunks whose saved samples fit the free device memory (<= 64 GB; hook "batched_post_mb"),
    // each followed by the binning kernel.  Other measure sets: one launch.
    long long chunk = io.n_mc;
    if (F.need_kl) {
        const size_t run_bytes = sizeof(double) * (size_t)io.n_t * (size_t)NP;
        // (cudaMemGetInfo costs milliseconds: only ensembles whose samples exceed 2 GB ask how much memory is free)
        size_t cap = (size_t)2 << 30;
        if (run_bytes * (size_t)io.n_mc > cap) {
            size_t free_b = 0, total_b = 0;
            MCBB_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
            cap = std::max<size_t>(cap, std::min<size_t>((size_t)64 << 30, free_b / 2 + scratch_size(94)));
        }
        if (const long long mb = tune_get("batched_post_mb", 0)) cap = (size_t)mb << 20;
        chunk = std::max<long long>(1, std::min<long long>(io.n_mc, (long long)(cap / run_bytes)));
        a.smp = (double*)scratch_get(94, run_bytes * (size_t)chunk);
        a.stat = (double*)scratch_get(108, sizeof(double) * 2 * (size_t)NP * (size_t)chunk);
        a.okflag = (int*)scratch_get(109, sizeof(int) * (size_t)chunk);
        if (!a.smp || !a.stat || !a.okflag) MCBB_FAIL("solve: device allocation failed");
    }
    for (long long r0 = 0; r0 < io.n_mc; r0 += chunk) {
        const long long r1 = std::min<long long>(io.n_mc, r0 + chunk);
        a.run_begin = r0;
        a.run_end = r1;
        const unsigned long long c0 = (unsigned long long)r0;
        MCBB_CUDA_OK(cudaMemcpyAsync(io.work_counter, &c0, sizeof(c0), cudaMemcpyHostToDevice, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));  // c0 is stack-owned
        const int g = (int)std::min<long long>((r1 - r0 + a.B - 1) / a.B, (long long)grid);
        int rc;
        if (TB == 2)
            rc = unit ? launch_variant<2, true>(a, smem, g, st) : launch_variant<2, false>(a, smem, g, st);
        else
            rc = unit ? launch_variant<4, true>(a, smem, g, st) : launch_variant<4, false>(a, smem, g, st);
        if (rc) return rc;
        if (F.need_kl) {
            k_batched_kl_post<<<(unsigned)(r1 - r0), 192, 0, st>>>(a, NP, r1 - r0);
            MCBB_LAUNCHED();
        }
    }
    *handled = true;
    return 0;
}

}  // namespace mcbb
