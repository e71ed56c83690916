// lane_core.cuh — one-trajectory-per-lane ensemble integrator with the eval_ode_run statistics fused into
// the post-transient tail (trajectories never reach HBM).
//
// Replaces, per trajectory: prob_func (src/setup_mc_prob.jl:722-741) -> OrdinaryDiffEq solve
// (:1106, Tsit5 / RK4 / FunctionMap semantics as restated in oracle/mcbb_oracle.c) -> eval_ode_run
// (src/eval_mc_prob.jl:75-198).
//
// SIMT structure: a lane is a small state machine; one iteration of the outer loop is one ROUND of R
// right-hand-side evaluations (R = 6 for Tsit5 with FSAL, 4 for RK4, 1 for a map).  Every lane of a warp
// is always at the same stage of its round, so the RHS (the dominant cost) runs fully converged.  A lane
// whose trajectory ends refills from a global work queue and spends one round on the two initial-dt RHS
// evaluations, so warps never wait for their slowest trajectory.
//
// KL-to-normal needs (mean, std) before binning (eval_mc_prob.jl:312-330).  Instead of storing the tail,
// the lane checkpoints its integrator state just before the first saved sample and integrates the tail a
// second time (bit-identical replay) while binning.
#pragma once
#include "systems.cuh"

namespace mcbb {

// Tsit5 tableau (OrdinaryDiffEq 5.26.7 tsit_tableaus).
// Row s (stage s+2, s = 0..5) holds the coefficients of k1..k_{s+1}.  Lanes of a warp are stage-aligned,
// so the __constant__ reads are warp-uniform broadcasts.
#define MCBB_TSIT5_A                                                                                      \
    {0.161, 0, 0, 0, 0, 0, -0.008480655492356989, 0.335480655492357, 0, 0, 0, 0, 2.8971530571054935,       \
     -6.359448489975075, 4.3622954328695815, 0, 0, 0, 5.325864828439257, -11.748883564062828,              \
     7.4955393428898365, -0.09249506636175525, 0, 0, 5.86145544294642, -12.92096931784711,                 \
     8.159367898576159, -0.071584973281401, -0.028269050394068383, 0, 0.09646076681806523, 0.01,           \
     0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}
#define MCBB_TSIT5_BT                                                                                     \
    {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,           \
     0.5823571654525552, -0.45808210592918697, 0.015151515151515152}
static __constant__ double c_ts_a[36] = MCBB_TSIT5_A;
static __constant__ double c_ts_bt[7] = MCBB_TSIT5_BT;
static const double h_ts_a[36] = MCBB_TSIT5_A;
static const double h_ts_bt[7] = MCBB_TSIT5_BT;
#ifdef __CUDA_ARCH__
#define TS_A(s, j) c_ts_a[(s)*6 + (j)]
#define TS_BT(j) c_ts_bt[(j)]
#else
#define TS_A(s, j) h_ts_a[(s)*6 + (j)]
#define TS_BT(j) h_ts_bt[(j)]
#endif

MCBB_HD void tsit5_btheta(const double th, double* b) {
    const double th2 = th * th;
    b[0] = th * (1. + th * (-2.763706197274826 + th * (2.9132554618219126 + th * -1.0530884977290216)));
    b[1] = th2 * (0.13169999999999998 + th * (-0.2234 + th * 0.1017));
    b[2] = th2 * (3.9302962368947516 + th * (-5.941033872131505 + th * 2.490627285651253));
    b[3] = th2 * (-12.411077166933676 + th * (30.33818863028232 + th * -16.548102889244902));
    b[4] = th2 * (37.50931341651104 + th * (-88.1789048947664 + th * 47.37952196281928));
    b[5] = th2 * (-27.896526289197286 + th * (65.09189467479366 + th * -34.87065786149661));
    b[6] = th2 * (1.5 + th * (-4. + th * 2.5));
}

// Julia Base mod(x, y), y > 0 (floored modulo on the exact remainder)
MCBB_HD double julia_mod_pos(const double x, const double y) {
    double r = fmod(x, y);
    if (r == 0.) return fabs(r);
    if (r < 0.) return r + y;
    return r;
}

// a * b + c with both roundings (Julia does not contract): the KL bin layout below is 1-ulp sensitive (the range length
// flips between 2k and 2k + 1 bins), so every kernel family must take it through the SAME arithmetic — left to the
// compiler, one inlined copy was contracted to an fma and another was not
MCBB_HD double mul_add_2r(const double a, const double b, const double c) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(__dmul_rn(a, b), c);
#else
    volatile double p = a * b;
    return p + c;
#endif
}

// Julia Base (:)(start, step, stop) length, Float64 fallback path (twiceprecision.jl)
MCBB_HD int julia_range_len(const double start, const double step, const double stop) {
    const double lf = (stop - start) / step;
    if (lf < 0) return 0;
    if (lf == 0) return 1;
    int len = (int)rint(lf) + 1;
    const double stop2 = mul_add_2r((double)(len - 1), step, start);
    if ((start < stop && stop < stop2) || (start > stop && stop > stop2)) len -= 1;
    return len;
}

// running "sum" state once a sample or the mean is Inf / NaN (finite+inf=inf, inf+inf=inf, inf-inf=NaN)
MCBB_HD double nonfinite_mean(const double mean, const double x) {
    const double big = 1.7976931348623157e308;
    if (fabs(mean) <= big) return x;
    if (fabs(x) <= big || x == mean) return mean;
    return nan("");
}

// accumulator slots per channel
enum { ACC_MEAN = 0, ACC_M2 = 1, ACC_SUBTR = 2, ACC_START = 3, ACC_BW = 4, ACC_NC = 5, ACC_SLOTS_KL = 6, ACC_SLOTS = 3,
       ACC_DELTA = 6, ACC_RESID = 7, ACC_SLOTS_CORR = 8 };
enum { SLOT_U = 0, SLOT_K1 = 1, SLOT_TMP = 8, SLOT_SCR = 9 };
enum { MODE_FETCH = 0, MODE_INIT = 1, MODE_RUN = 2, MODE_DONE = 3 };

struct LaneShared {  // CTA-wide read-only data staged in shared memory
    const double* t_save;
    SysTabs tabs;
};

// SOLVAR: one solver option (io.solver_field) is taken per run from io.solver_par — SolverParameterVar /
// SolverVarEnsembleProblem, src/solver_parametervar.jl:75-90 (`solve(prob_i, alg; name => par[i], kwargs...)`).
template <int SYS, class WorkQ, bool SOLVAR = false>
MCBB_HD void lane_main(const SysDev& S, const SolverDev& o_in, const FeatDev& F, const SolveIO& io,
                       const LaneShared& L, double* st, double* acc, unsigned short* bins, double* como,
                       const int bs, WorkQ& wq) {
    SolverDev o_lane;  // per-lane copy, only live in the SOLVAR instantiations
    if constexpr (SOLVAR) o_lane = o_in;
    const SolverDev& o = SOLVAR ? o_lane : o_in;
    const int n = S.n_dim;
    const int n_ch = F.n_ch;
    const int nb = F.kl_bins;
    const int n_t = io.n_t;
    const int ns_used = n_t - 1;  // sol[i_dim, 2:end]
    const double* tsave = L.t_save;
    const int alg = o.alg;
    const int R = alg == MCBB_ALG_TSIT5 ? 6 : (alg == MCBB_ALG_RK4 ? 4 : 1);
    const double NaN = nan("");

#define X(slot, d) st[((size_t)(slot)*n + (d)) * bs]
#define A(slot, ch) acc[((size_t)(slot)*n_ch + (ch)) * bs]
#define B(ch, b) bins[((size_t)(ch)*nb + (b)) * bs]
#define CM(p) como[(size_t)(p)*bs]  /* co-moments of channel pairs (i > j), p = i(i-1)/2 + j */
    const bool corr = F.matrix_meas != MCBB_MAT_NONE;  // correlation_ecdf or correlation_hist
    // DiffEqBase.ODE_DEFAULT_NORM works on the ELEMENTS of the state array: for the Complex{Float64} state of
    // Stuart-Landau an element is one complex number — its error scale uses the complex modulus (shared by both
    // components) and the RMS divides by N, not 2N (calculate_residuals / ODE_DEFAULT_NORM, DiffEqBase 6.10.1)
    constexpr bool CPLX = (SYS == MCBB_SYS_STUART_LANDAU);
    const int n_elem = CPLX ? n / 2 : n;
    auto elem_abs = [&](const int slot, const int d) -> double {
        if (CPLX) return hypot(X(slot, d & ~1), X(slot, d | 1));
        return fabs(X(slot, d));
    };

    int mode = MODE_FETCH;
    long long run = -1;
    int pass = 0, isave = 0, ret = 0;
    bool clipped = false, ckpt_done = false;
    double sc[MCBB_N_SCALARS];
    double t = 0., dt = 0., dts = 0., qold = 0., gsum = 0.;
    double d0 = 0., d1 = 0., dt0 = 0.;
    long long iter = 0, nrej = 0, natt = 0;

    // channel -> state dimension
    auto ch_dim = [&](int ch) -> int {
        const int ii = ch / F.n_comp, c = ch - ii * F.n_comp;
        const int sidx = F.filter ? (F.filter[ii] - 1) : ii;
        return F.n_comp == 2 ? 2 * sidx + c : sidx;
    };

    // begin a step attempt: maxiters / tstop clipping / dtmin / checkpoint peek.  Returns false when the
    // pass must end (ret set).
    auto begin_attempt = [&]() -> bool {
        if (++iter > o.maxiters) {
            ret = MCBB_RET_MAXITERS;
            return false;
        }
        if (alg == MCBB_ALG_FUNCTIONMAP) {
            dts = 1.;
            clipped = false;
        } else {
            dts = dt;
            clipped = false;
            if (alg == MCBB_ALG_RK4) {
                if (t + dts >= o.t1 || fabs((t + dts) - o.t1) < 100. * 2.220446049250313e-16 * fabs(o.t1)) {
                    dts = o.t1 - t;
                    clipped = true;
                }
            } else {
                if (dts >= o.t1 - t) {
                    dts = o.t1 - t;
                    clipped = true;
                }
                if (!clipped && dts <= o.dtmin) {
                    ret = MCBB_RET_DTLESSTHANMIN;
                    return false;
                }
            }
        }
        if (F.need_kl && pass == 0 && !ckpt_done && io.ckpt != nullptr) {
            const double tn = clipped ? o.t1 : t + dts;
            if (n_t > 0 && tn >= tsave[0]) {  // this attempt may produce the first saved sample
                double* ck = io.ckpt + run;
                const long long str = io.n_mc;
                for (int d = 0; d < n; d++) {
                    ck[(long long)d * str] = X(SLOT_U, d);
                    ck[(long long)(n + d) * str] = X(SLOT_K1, d);
                }
                ck[(long long)(2 * n) * str] = t;
                ck[(long long)(2 * n + 1) * str] = dt;
                ck[(long long)(2 * n + 2) * str] = qold;
                ck[(long long)(2 * n + 3) * str] = (double)(iter - 1);
            }
        }
        return true;
    };

    auto start_pass = [&]() {
        isave = 0;
        ret = 0;
        gsum = 0.;
        if (pass == 0) {
            t = o.t0;
            iter = 0;
            nrej = 0;
            natt = 0;
            qold = o.qoldinit;
            ckpt_done = false;
            for (int d = 0; d < n; d++) X(SLOT_U, d) = io.ic[(long long)d * io.n_mc + run];
            for (int ch = 0; ch < n_ch; ch++) {
                A(ACC_MEAN, ch) = 0.;
                A(ACC_M2, ch) = 0.;
                A(ACC_SUBTR, ch) = 0.;
            }
            if (corr)
                for (int p = 0; p < n_ch * (n_ch - 1) / 2; p++) CM(p) = 0.;
            if (alg == MCBB_ALG_FUNCTIONMAP) {
                mode = MODE_RUN;
                if (!begin_attempt()) mode = MODE_FETCH;  // only maxiters < 1
            } else {
                mode = MODE_INIT;
            }
        } else {  // KL replay from the checkpoint
            const double* ck = io.ckpt + run;
            const long long str = io.n_mc;
            for (int d = 0; d < n; d++) {
                X(SLOT_U, d) = ck[(long long)d * str];
                X(SLOT_K1, d) = ck[(long long)(n + d) * str];
            }
            t = ck[(long long)(2 * n) * str];
            dt = ck[(long long)(2 * n + 1) * str];
            qold = ck[(long long)(2 * n + 2) * str];
            iter = (long long)ck[(long long)(2 * n + 3) * str];
            for (int ch = 0; ch < n_ch; ch++) {
                // bin layout of empirical_1D_KL_divergence_hist, eval_mc_prob.jl:319-328
                const double mu = A(ACC_MEAN, ch), sig = A(ACC_M2, ch);
                const double k = (nb - 1) / 2.;
                const double bw = (F.kl_nstd * sig) / k;
                const double start = mul_add_2r(-k, bw, mu), stop = mul_add_2r(k, bw, mu);
                int nc = 0;
                if (sig >= F.kl_tol) nc = julia_range_len(start, bw, stop);
                if (nc > nb) nc = nb;
                if (sig != sig) nc = -1;  // NaN std (non-finite samples): the reference's KL is NaN
                A(ACC_START, ch) = start;
                A(ACC_BW, ch) = bw;
                A(ACC_NC, ch) = (double)nc;
                for (int b = 0; b < nb; b++) B(ch, b) = 0;
            }
            mode = MODE_RUN;
            begin_attempt();  // cannot fail: the same attempt succeeded in pass 0
        }
    };

    // one saved sample (index isave in t_save; value of state dim d given by val(d))
    auto consume = [&](auto&& val) {
        if (io.saved && pass == 0) {  // sample export: every sample, or only sol[end]
            if (io.save_mode == 1) {
                double* dst = io.saved + ((size_t)run * n_t + isave) * n;
                for (int d = 0; d < n; d++) dst[d] = val(d);
            } else if (isave == n_t - 1) {
                for (int d = 0; d < n; d++) io.saved[(size_t)d * io.n_mc + run] = val(d);
            }
        }
        if (isave >= 1) {  // the first saved sample is dropped, eval_mc_prob.jl:152
            ckpt_done = true;
            const double rk = 1. / (double)isave;
            for (int ch = 0; ch < n_ch; ch++) {
                double x = val(ch_dim(ch));
                if (pass == 0 && F.n_global > 0) gsum += x;
                if (F.cyclic) {  // _cyclic_setback!, eval_mc_prob.jl:286-296
                    if (isave == 1) {
                        const double twopi = 6.283185307179586, pi = 3.141592653589793;
                        double v = julia_mod_pos(x, twopi);
                        if (v > pi) v -= twopi;
                        A(ACC_SUBTR, ch) = x - v;
                        x = v;
                    } else {
                        x = x - A(ACC_SUBTR, ch);
                    }
                }
                if (pass == 0) {  // online mean / M2
                    const double mean = A(ACC_MEAN, ch);
                    const double delta = x - mean;
                    if (fabs(delta) <= 1.7976931348623157e308) {
                        const double mnew = fma(delta, rk, mean);
                        A(ACC_MEAN, ch) = mnew;
                        A(ACC_M2, ch) = fma(delta, x - mnew, A(ACC_M2, ch));
                        if (corr) {
                            A(ACC_DELTA, ch) = delta;
                            A(ACC_RESID, ch) = x - mnew;
                        }
                    } else {  // Inf / NaN samples: keep sum(x)/n and sum((x-m)^2) semantics of the reference
                        A(ACC_MEAN, ch) = nonfinite_mean(mean, x);
                        A(ACC_M2, ch) = NaN;
                        if (corr) {
                            A(ACC_DELTA, ch) = NaN;
                            A(ACC_RESID, ch) = NaN;
                        }
                    }
                } else {  // fit(Histogram, u, edges; closed=:left): searchsortedlast on the edge VECTOR
                    const int nc = (int)A(ACC_NC, ch);
                    if (nc > 0) {
                        const double start = A(ACC_START, ch), bw = A(ACC_BW, ch);
                        const double hb = bw / 2.;
                        const double last_edge = mul_add_2r((nb - 1) / 2., bw, hb);  // `mu` missing: eval_mc_prob.jl:328
                        int lo = 0, hi = nc + 2;
                        while (lo < hi - 1) {
                            const int m = (lo + hi) >> 1;
                            const double e = (m <= nc) ? fma((double)(m - 1), bw, start) - hb : last_edge;
                            if (x < e)
                                hi = m;
                            else
                                lo = m;
                        }
                        if (lo >= 1 && lo <= nc) B(ch, lo - 1) += 1;
                    }
                }
            }
            if (corr && pass == 0) {  // online co-moments: C_ij += (x_i - mean_i_old)(x_j - mean_j_new)
                int p = 0;
                for (int i = 1; i < n_ch; i++) {
                    const double di = A(ACC_DELTA, i);
                    for (int j = 0; j < i; j++, p++) CM(p) = fma(di, A(ACC_RESID, j), CM(p));
                }
            }
        }
        isave++;
    };

    auto write_nan_features = [&]() {
        for (int m = 0; m < F.n_meas; m++)
            for (int ii = 0; ii < F.n_filter; ii++) io.feat[((long long)m * F.n_filter + ii) * io.n_mc + run] = NaN;
        for (int g = 0; g < F.n_global; g++)
            if (io.glob) io.glob[(long long)g * io.n_mc + run] = NaN;
        if (corr && io.matrix)
            for (int b = 0; b < F.corr_nbins; b++) io.matrix[(long long)b * io.n_mc + run] = NaN;
    };

    auto end_pass = [&]() {
        const bool ok = (ret == MCBB_RET_SUCCESS) && isave >= n_t;
        if (pass == 0) {
            if (!ok) {
                write_nan_features();
            } else {
                if (corr && io.matrix) {  // correlation_ecdf, eval_mc_prob.jl:462-488
                    const int nbn = F.corr_nbins;
                    int h[64];
                    for (int b = 0; b < nbn; b++) h[b] = 0;
                    const double step = 1. / (double)nbn;
                    const int n_st = F.n_filter;  // states: real, or complex as (re, im) channel pairs
                    for (int i = 1; i < n_st; i++)
                        for (int j = 0; j < i; j++) {
                            double c;
                            if (F.n_comp == 1) {
                                c = CM(i * (i - 1) / 2 + j) / (sqrt(A(ACC_M2, i)) * sqrt(A(ACC_M2, j)));
                                c = fabs(fmin(1., fmax(-1., c)));  // clampcor, then abs
                            } else {
                                // cor of complex series: sum (z_i - m_i) conj(z_j - m_j) over the samples, from the
                                // real co-moments of the (re, im) channels; clampcor leaves complex values alone
                                const int ai = 2 * i, bi = 2 * i + 1, aj = 2 * j, bj = 2 * j + 1;
                                const double re = CM(ai * (ai - 1) / 2 + aj) + CM(bi * (bi - 1) / 2 + bj);
                                const double im = CM(bi * (bi - 1) / 2 + aj) - CM(ai * (ai - 1) / 2 + bj);
                                c = hypot(re, im) / (sqrt(A(ACC_M2, ai) + A(ACC_M2, bi)) * sqrt(A(ACC_M2, aj) + A(ACC_M2, bj)));
                            }
                            if (c != c) continue;  // zero-variance dimension: dropped
                            double f = c / step + 1.;          // Base.searchsortedlast on the range 0:1/nbins:1
                            f = fmin(fmax(f, 1.), (double)(nbn + 1));
                            const int n0 = (int)rint(f);
                            const int idx = (c < (double)(n0 - 1) / (double)nbn) ? n0 - 1 : n0;
                            if (idx >= 1 && idx <= nbn) h[idx - 1] += 2;  // (i,j) and (j,i); the diagonal (1.0) is dropped
                        }
                    double cum = 0., tot = 0.;
                    for (int b = 0; b < nbn; b++) tot += (double)h[b];
                    if (F.matrix_meas == MCBB_MAT_CORR_HIST)  // the weights themselves, eval_mc_prob.jl:472-481
                        for (int b = 0; b < nbn; b++) io.matrix[(long long)b * io.n_mc + run] = (double)h[b];
                    else
                        for (int b = 0; b < nbn; b++) {  // ecdf_hist, eval_clustering.jl:759-762
                            cum += (double)h[b];
                            io.matrix[(long long)b * io.n_mc + run] = cum / tot;
                        }
                }
                for (int ch = 0; ch < n_ch; ch++)  // std(u; mean, corrected=true)
                    {
                    const double m2 = A(ACC_M2, ch);  // fmax would swallow a NaN
                    A(ACC_M2, ch) = sqrt((m2 < 0. ? 0. : m2) / (double)(ns_used - 1));
                }
                for (int m = 0; m < F.n_meas; m++) {
                    if (F.meas[m] == MCBB_MEAS_KL_HIST) continue;
                    const int slot = F.meas[m] == MCBB_MEAS_MEAN ? ACC_MEAN : ACC_M2;
                    for (int ii = 0; ii < F.n_filter; ii++)
                        io.feat[((long long)m * F.n_filter + ii) * io.n_mc + run] = A(slot, ii * F.n_comp + F.comp[m]);
                }
                for (int g = 0; g < F.n_global; g++)
                    if (io.glob) io.glob[(long long)g * io.n_mc + run] = gsum;
            }
            if (io.retcode) io.retcode[run] = ret;
            if (io.nsteps) {
                io.nsteps[run] = natt;
                io.nsteps[io.n_mc + run] = nrej;
            }
            if (ok && F.need_kl) {
                pass = 1;
                start_pass();
            } else {
                mode = MODE_FETCH;
            }
        } else {
            for (int m = 0; m < F.n_meas; m++) {
                if (F.meas[m] != MCBB_MEAS_KL_HIST) continue;
                for (int ii = 0; ii < F.n_filter; ii++) {
                    const int ch = ii * F.n_comp + F.comp[m];
                    const int nc = (int)A(ACC_NC, ch);
                    double kl = nc < 0 ? NaN : 0.;
                    if (nc > 0) {  // sig >= sig_tol
                        const double mu = A(ACC_MEAN, ch), sig = A(ACC_M2, ch);
                        const double start = A(ACC_START, ch), bw = A(ACC_BW, ch);
                        double tot = 0., rs = 0.;
                        for (int b = 0; b < nc; b++) {
                            tot += (double)B(ch, b);
                            const double z = (fma((double)b, bw, start) - mu) / sig;
                            rs += exp(-(z * z) / 2.);
                        }
                        for (int b = 0; b < nc; b++) {
                            const double p = (double)B(ch, b) / tot;
                            if (p > 0.) {
                                const double z = (fma((double)b, bw, start) - mu) / sig;
                                const double q = exp(-(z * z) / 2.) / rs;
                                kl += p * log(p / q);
                            }
                        }
                        if (tot == 0.) kl = NaN;  // 0/0 weights -> NaN, as StatsBase would give
                    }
                    io.feat[((long long)m * F.n_filter + ii) * io.n_mc + run] = kl;
                }
            }
            mode = MODE_FETCH;
        }
    };

    // ------------------------------------------------------------------------------------------------
    for (;;) {
        if (mode == MODE_FETCH) {
            run = wq.next();
            if (run >= io.n_mc) {
                mode = MODE_DONE;
            } else {
                for (int q = 0; q < MCBB_N_SCALARS; q++) sc[q] = S.scalars[q];
                for (int p = 0; p < S.n_par; p++) sc[S.par_slot[p]] = io.par[(long long)p * io.n_mc + run];
                if constexpr (SOLVAR) {
                    const double v = io.solver_par[run];
                    switch (io.solver_field) {
                    case MCBB_SOLVER_ABSTOL: o_lane.abstol = v; break;
                    case MCBB_SOLVER_RELTOL: o_lane.reltol = v; break;
                    case MCBB_SOLVER_DT: o_lane.dt_user = v; break;
                    case MCBB_SOLVER_DTMAX: o_lane.dtmax = v; break;
                    case MCBB_SOLVER_DTMIN: o_lane.dtmin = v; break;
                    default: break;
                    }
                }
                pass = 0;
                start_pass();
            }
        }
        // Maps (FunctionMap): iterations of the transient produce no sample and cannot fail except by a non-finite
        // state or maxiters — they run in a tight loop here instead of one full round each (C2: 800 of 1000 iterations).
        // Exactly the generic sequence (rhs -> count the attempt -> commit -> begin the next attempt); an iteration
        // that would produce a sample, reach the end of the span or a NaN is left to the round below, untouched.
        if (alg == MCBB_ALG_FUNCTIONMAP) {
            while (mode == MODE_RUN && ret == 0) {
                const double tnew = t + 1.;
                if (!(tnew < o.t1)) break;
                if (isave < n_t && tsave[isave] <= tnew) break;
                rhs_lane<SYS>(S, L.tabs, sc, &X(SLOT_U, 0), &X(SLOT_TMP, 0), &X(SLOT_SCR, 0), bs);
                bool bad = false;
                for (int d = 0; d < n; d++) bad |= (X(SLOT_TMP, d) != X(SLOT_TMP, d));
                if (bad) break;
                natt += (pass == 0);
                t = tnew;
                for (int d = 0; d < n; d++) X(SLOT_U, d) = X(SLOT_TMP, d);
                if (!begin_attempt()) {
                    end_pass();
                    break;
                }
            }
        }
        if (wq.all_done(mode == MODE_DONE)) break;

        // ---- one round of R right-hand-side evaluations, stage-aligned across the warp -----------------
        for (int s = 0; s < R; s++) {
            int in_slot = SLOT_TMP, out_slot = SLOT_K1 + 1 + s;
            if (mode == MODE_INIT) {
                in_slot = (s == 0) ? SLOT_U : SLOT_TMP;
                out_slot = (s == 0) ? SLOT_K1 : SLOT_K1 + 1;
            } else if (mode == MODE_RUN) {
                if (alg == MCBB_ALG_TSIT5) {
                    for (int d = 0; d < n; d++) {  // stage s+2 input: u + dt * sum_j a[s][j] k_j
                        double lin = TS_A(s, 0) * X(SLOT_K1, d);
                        for (int j = 1; j <= s; j++) lin = fma(TS_A(s, j), X(SLOT_K1 + j, d), lin);
                        X(SLOT_TMP, d) = fma(dts, lin, X(SLOT_U, d));
                    }
                } else if (alg == MCBB_ALG_RK4) {
                    if (s < 3) {
                        const double h = (s < 2) ? dts / 2. : dts;
                        for (int d = 0; d < n; d++) X(SLOT_TMP, d) = fma(h, X(SLOT_K1 + s, d), X(SLOT_U, d));
                    } else {
                        for (int d = 0; d < n; d++)
                            X(SLOT_TMP, d) = X(SLOT_U, d) + (dts / 6.) * (2. * (X(SLOT_K1 + 1, d) + X(SLOT_K1 + 2, d)) +
                                                                          (X(SLOT_K1, d) + X(SLOT_K1 + 3, d)));
                    }
                } else {
                    in_slot = SLOT_U;
                    out_slot = SLOT_TMP;
                }
            }
            const int init_evals = (alg == MCBB_ALG_TSIT5 && !(o.dt_user > 0)) ? 2 : 1;
            const bool idle = (mode == MODE_DONE) || (mode == MODE_INIT && s >= init_evals) || (mode == MODE_FETCH);
            if (!idle)
                rhs_lane<SYS>(S, L.tabs, sc, &X(in_slot, 0), &X(out_slot, 0), &X(SLOT_SCR, 0), bs);

            if (mode == MODE_INIT) {  // OrdinaryDiffEq initdt.jl (Hairer-Wanner), order 5
                if (alg == MCBB_ALG_RK4 || o.dt_user > 0) {
                    if (s == 0) dt = o.dt_user;
                } else if (s == 0) {
                    double a0 = 0., a1 = 0.;
                    for (int d = 0; d < n; d++) {
                        const double u0 = X(SLOT_U, d);
                        const double sk = o.abstol + elem_abs(SLOT_U, d) * o.reltol;
                        const double v0 = u0 / sk, v1 = X(SLOT_K1, d) / sk;
                        a0 = fma(v0, v0, a0);
                        a1 = fma(v1, v1, a1);
                    }
                    d0 = sqrt(a0 / n_elem);
                    d1 = sqrt(a1 / n_elem);
                    dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
                    dt0 = fmin(dt0, o.dtmax);
                    for (int d = 0; d < n; d++) X(SLOT_TMP, d) = fma(dt0, X(SLOT_K1, d), X(SLOT_U, d));
                } else if (s == 1) {
                    double a2 = 0.;
                    for (int d = 0; d < n; d++) {
                        const double sk = o.abstol + elem_abs(SLOT_U, d) * o.reltol;
                        const double v = (X(SLOT_K1 + 1, d) - X(SLOT_K1, d)) / sk;
                        a2 = fma(v, v, a2);
                    }
                    const double d2 = sqrt(a2 / n_elem) / dt0;
                    const double dm = fmax(d1, d2);
                    const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10., -(2. + log10(dm)) / 5.);
                    dt = fmin(fmin(100. * dt0, dt1), o.dtmax);
                }
            }
        }

        // ---- end of round -------------------------------------------------------------------------------
        if (mode == MODE_INIT) {
            mode = MODE_RUN;
            if (!begin_attempt()) end_pass();
        } else if (mode == MODE_RUN) {
            natt += (pass == 0);
            bool accept = true;
            double q = 1., q11 = 0.;
            if (alg == MCBB_ALG_TSIT5) {
                double e2 = 0.;
                for (int d = 0; d < n; d++) {
                    double lin = TS_BT(0) * X(SLOT_K1, d);
                    for (int j = 1; j < 7; j++) lin = fma(TS_BT(j), X(SLOT_K1 + j, d), lin);
                    const double ut = dts * lin;
                    const double r = ut / (o.abstol + fmax(elem_abs(SLOT_U, d), elem_abs(SLOT_TMP, d)) * o.reltol);
                    e2 = fma(r, r, e2);
                }
                const double EEst = sqrt(e2 / n_elem);
                if (EEst == 0.) {
                    q = 1. / o.qmax;
                } else {
                    q11 = pow(EEst, o.beta1);
                    q = q11 / pow(qold, o.beta2);
                    q = fmax(1. / o.qmax, fmin(1. / o.qmin, q / o.gamma));
                }
                accept = !(EEst > 1.);
                if (accept) qold = fmax(EEst, o.qoldinit);
            }
            if (accept) {
                const double tnew = (alg == MCBB_ALG_FUNCTIONMAP) ? t + 1. : (clipped ? o.t1 : t + dts);
                if (alg == MCBB_ALG_TSIT5) {
                    while (isave < n_t && tsave[isave] <= tnew) {
                        const double ts = tsave[isave];
                        if (ts == tnew) {
                            consume([&](int d) { return X(SLOT_TMP, d); });
                        } else {
                            double b[7];
                            tsit5_btheta((ts - t) / dts, b);
                            consume([&](int d) {
                                double lin = X(SLOT_K1, d) * b[0];
                                for (int j = 1; j < 7; j++) lin = fma(X(SLOT_K1 + j, d), b[j], lin);
                                return fma(dts, lin, X(SLOT_U, d));
                            });
                        }
                    }
                } else if (alg == MCBB_ALG_RK4) {  // cubic Hermite between step ends (u, f at both ends)
                    while (isave < n_t && tsave[isave] <= tnew) {
                        const double ts = tsave[isave];
                        if (ts == tnew) {
                            consume([&](int d) { return X(SLOT_TMP, d); });
                        } else {
                            const double th = (ts - t) / dts;
                            consume([&](int d) {
                                const double u0 = X(SLOT_U, d), u1 = X(SLOT_TMP, d);
                                return (1. - th) * u0 + th * u1 +
                                       th * (th - 1.) * ((1. - 2. * th) * (u1 - u0) + (th - 1.) * dts * X(SLOT_K1, d) +
                                                         th * dts * X(SLOT_K1 + 4, d));
                            });
                        }
                    }
                } else {
                    while (isave < n_t && tsave[isave] <= tnew) consume([&](int d) { return X(SLOT_TMP, d); });
                }
                t = tnew;
                bool bad = false;
                const int fsal = (alg == MCBB_ALG_TSIT5) ? SLOT_K1 + 6 : SLOT_K1 + 4;
                for (int d = 0; d < n; d++) {
                    const double un = X(SLOT_TMP, d);
                    bad |= (un != un);
                    X(SLOT_U, d) = un;
                    if (alg != MCBB_ALG_FUNCTIONMAP) X(SLOT_K1, d) = X(fsal, d);
                }
                if (alg == MCBB_ALG_TSIT5) dt = fmin(o.dtmax, fmax(o.dtmin, dts / q));
                if (bad) ret = MCBB_RET_UNSTABLE;
            } else {
                nrej += (pass == 0);
                dt = dts / fmin(1. / o.qmin, q11 / o.gamma);
            }
            if (ret != 0 || !(t < o.t1)) {
                end_pass();
            } else if (!begin_attempt()) {
                end_pass();
            }
        }
    }
#undef X
#undef A
#undef B
#undef CM
}

}  // namespace mcbb
