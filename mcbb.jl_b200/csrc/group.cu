// group.cu — ensemble Tsit5 for SMALL networks: a group of G = 8 / 16 / 32 lanes integrates one trajectory, one
// oscillator per lane.
//
// Why: the lane kernel (lane_core.cuh) gives every lane a whole trajectory.  For the small networks of BASELINE configs
// C1 (kuramoto_network, N = 6) and C3 (second_order_kuramoto, N = 10, 15 edges) that leaves two problems: the per-lane
// stage vectors (9 x n_dim doubles in shared memory) cap the SM at 4 warps for C3, and every right-hand side is a serial
// chain of N sincos (C1) or N_edges sines (C3) per lane — C1's 2 500 runs are one warp per SM running 31 steps of pure
// dependent latency.  Here the sines of one right-hand side run in parallel lanes, every stage vector is a register
// (one or two components per lane, static indexing: the six stages are unrolled), and a warp carries 32 / G trajectories.
//
// Arithmetic is the lane kernel's, operation by operation: the sums the reference takes over the state (error norm,
// initial-dt norms, coupling sums, incidence products) are taken in the SAME ORDER, every lane gathering the terms by
// shuffle and adding them sequentially — so measures, return codes and step counts are bitwise those of the lane kernel
// (tests/test_gpu_parity.py::test_group_kernel_equals_lane_kernel_bitwise).  Control (PI controller, accept / reject,
// save window, OrdinaryDiffEq initdt) is replicated in all lanes of a group; groups of one warp advance in lock-step
// rounds of six right-hand-side evaluations like the lanes of the lane kernel, a finished group refills from the queue.
//
//   kuramoto_network (systems.jl:119-130):  du_i = w_i + K/N (cos u_i (A sin u)_i - sin u_i (A cos u)_i), lane i holds
//       row i of A in shared memory and gathers sin u_j, cos u_j for j = 0 .. N-1 in order.
//   kuramoto (mean field, systems.jl:85-94): the same with the two plain sums over all oscillators.
//   second_order_kuramoto (systems.jl:198-202):  lane e < N_edges evaluates sin((E' theta)_e) from the two ends of edge e;
//       lane i < N owns (theta_i, omega_i) and adds its incident flows in ascending edge order.
// Features: mean / std (Welford, per lane), cyclic_setback, state_filter, KL-hist from the saved samples of the run in a
// device scratch (um_kl_hist: the lane kernel's binning arithmetic).
#include <algorithm>
#include <vector>

#include "lane_core.cuh"
#include "umma_common.cuh"  // um_kl_hist

namespace mcbb {

constexpr int GR_DEGMAX = 8;   // incident edges per node (second_order_kuramoto)
constexpr int GR_THREADS = 128;

struct GroupArgs {
    int N, E, n_dim;              // oscillators, edges (second order), state dimensions
    const double* mat;            // kuramoto_network: A [N][N] column-major (device)
    const double* vec_a;          // [N] natural frequencies / drive
    const int* edge_node;         // [2][EPL * G] ends of edge e (second order); -1 = no entry
    const double* edge_val;       // [2][EPL * G]
    const int* node_deg;          // [G]
    const int* node_edge;         // [GR_DEGMAX][G] incident edges of node i, ascending
    const double* node_val;       // [GR_DEGMAX][G]
    int degmax;
    const int* fpos;              // [2][G] output position of (component, lane) or -1
    double scalars[MCBB_N_SCALARS];
    int n_par;
    int par_slot[MCBB_MAX_PAR];
    SolverDev o;
    int n_meas, meas[MCBB_MAX_MEAS], n_filter, cyclic;
    int need_kl, kl_bins;
    double kl_nstd, kl_tol;
    SolveIO io;
    double* smp;                  // [runs of the chunk][n_t][n_dim] saved samples (KL only; binned by k_group_kl_post)
    double* stat;                 // [runs of the chunk][n_dim][2] mean, std of every reported dimension
    int* okflag;                  // [runs of the chunk]
    long long run_begin, run_end; // the runs this launch integrates (one chunk of the ensemble when samples are kept)
};

template <int G>
__device__ __forceinline__ double gr_shfl(const double v, const int src) {
    return __shfl_sync(0xffffffffu, v, src, G);
}

// SYS: MCBB_SYS_KURAMOTO_NETWORK (one component per lane) or MCBB_SYS_SECOND_ORDER_KURAMOTO (theta_i, omega_i per lane)
// EPL: edges per lane (second order): 2 lets a 32-lane group carry up to 64 edges — the paper's N = 30, 3-regular power
// grid has 45 (paper/kuramoto/1-simulate.jl:22-26); lane l evaluates the sines of edges l and l + G
// KL: the instantiation that keeps the saved samples for the KL-hist measures
template <int SYS, int G, int EPL = 1, bool KL = false>
__global__ void __launch_bounds__(GR_THREADS) k_solve_group(const GroupArgs a) {
    constexpr int NC = (SYS == MCBB_SYS_SECOND_ORDER_KURAMOTO) ? 2 : 1;
    extern __shared__ __align__(16) unsigned char gr_smem[];
    double* s_ts = reinterpret_cast<double*>(gr_smem);                    // [n_t]
    double* s_mat = s_ts + ((a.io.n_t + 1) & ~1);                         // kuramoto: [N][N]; second order: incident-edge tables
    double* s_nv = s_mat;                                                 // [GR_DEGMAX][G] incidence values
    int* s_ne = reinterpret_cast<int*>(s_nv + GR_DEGMAX * G);             // [GR_DEGMAX][G] incident edges, ascending
    const int tid = threadIdx.x;
    const int gl = tid & (G - 1);                                         // lane within the group = oscillator / edge index
    const int N = a.N, n_t = a.io.n_t;
    const long long n_mc = a.io.n_mc;
    const SolverDev& o = a.o;
    for (int i = tid; i < n_t; i += GR_THREADS) s_ts[i] = a.io.t_save[i];
    if (SYS == MCBB_SYS_KURAMOTO_NETWORK) {
        for (int i = tid; i < N * N; i += GR_THREADS) s_mat[i] = a.mat[i];
    } else if (SYS == MCBB_SYS_SECOND_ORDER_KURAMOTO) {
        for (int i = tid; i < GR_DEGMAX * G; i += GR_THREADS) {
            s_nv[i] = a.node_val[i];
            s_ne[i] = a.node_edge[i];
        }
    }
    __syncthreads();
    const bool own = gl < N;                       // this lane owns an oscillator
    const int gi = own ? gl : 0;                   // safe index for the table reads of idle lanes
    const double w_i = a.vec_a[gi];
    // second order: the two ends of this lane's edge, the incident edges of this lane's node
    int en0[EPL], en1[EPL], deg = 0;
    double ev0[EPL], ev1[EPL];
#pragma unroll
    for (int p = 0; p < EPL; p++) {
        en0[p] = en1[p] = 0;
        ev0[p] = ev1[p] = 0.;
    }
    if (SYS == MCBB_SYS_SECOND_ORDER_KURAMOTO) {
#pragma unroll
        for (int p = 0; p < EPL; p++) {  // tables are [2][EPL * G]
            en0[p] = a.edge_node[p * G + gl];
            en1[p] = a.edge_node[EPL * G + p * G + gl];
            ev0[p] = a.edge_val[p * G + gl];
            ev1[p] = a.edge_val[EPL * G + p * G + gl];
        }
        deg = a.node_deg[gl];
    }
    int fpos[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) fpos[c] = own ? a.fpos[c * G + gl] : -1;
    double* smp = nullptr;  // this run's rows of the sample scratch (KL only)
    const double NaN = nan("");

    // ---- trajectory state: registers -------------------------------------------------------------------------------
    double u[NC], k[7][NC], tmp[NC], am[NC], a2[NC], as_[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        u[c] = tmp[c] = am[c] = a2[c] = as_[c] = 0.;
#pragma unroll
        for (int q = 0; q < 7; q++) k[q][c] = 0.;
    }
    int mode = MODE_FETCH, isave = 0, ret = 0;
    long long run = -1, iter = 0, nrej = 0, natt = 0;
    bool clipped = false;
    double sc[MCBB_N_SCALARS];
#pragma unroll
    for (int q = 0; q < MCBB_N_SCALARS; q++) sc[q] = 0.;
    double t = 0., dt = 0., dts = 0., qold = 0., d0 = 0., d1 = 0., dt0 = 0.;

    // sum over the state in the reference's order (d = 0 .. n_dim - 1): every lane gathers the terms and adds them
    // itself, so all lanes of a group hold the same bits.  term(c) = this lane's contribution of component c.
    auto state_sum_sq = [&](const double (&v)[NC]) -> double {
        double acc = 0.;
#pragma unroll
        for (int c = 0; c < NC; c++)
            for (int j = 0; j < N; j++) {
                const double vj = gr_shfl<G>(v[c], j);
                acc = fma(vj, vj, acc);
            }
        return acc;
    };

    // right-hand side of the whole group: y (this lane's components) -> out.  All lanes of the warp execute it together.
    auto rhs = [&](const double (&y)[NC], double (&out)[NC]) {
        if constexpr (SYS == MCBB_SYS_KURAMOTO) {  // mean field, systems.jl:85-94: the two sums in index order
            double s, c;
            sincos(y[0], &s, &c);
            double sum_s = 0., sum_c = 0.;
            for (int j = 0; j < N; j++) {
                sum_s += gr_shfl<G>(s, j);
                sum_c += gr_shfl<G>(c, j);
            }
            const double KN = sc[0] / (double)N;
            out[0] = fma(KN, c * sum_s - s * sum_c, w_i);
        } else if constexpr (SYS == MCBB_SYS_KURAMOTO_NETWORK) {
            double s, c;
            sincos(y[0], &s, &c);
            double as = 0., ac = 0.;
            for (int j = 0; j < N; j++) {
                const double sj = gr_shfl<G>(s, j), cj = gr_shfl<G>(c, j);
                const double aij = s_mat[gi + N * j];
                if (aij != 0.) {
                    as = fma(aij, sj, as);
                    ac = fma(aij, cj, ac);
                }
            }
            const double KN = sc[0] / (double)N;
            out[0] = fma(KN, c * as - s * ac, w_i);
        } else {
            // lane e: sin of (incidence' theta)_e; entries in ascending node order, as the compressed table lists them
            double sv[EPL];
#pragma unroll
            for (int p = 0; p < EPL; p++) {
                double arg = 0.;
                const double t0 = gr_shfl<G>(y[0], en0[p] < 0 ? 0 : en0[p]);
                const double t1 = gr_shfl<G>(y[0], en1[p] < 0 ? 0 : en1[p]);
                if (en0[p] >= 0) arg += ev0[p] * t0;
                if (en1[p] >= 0) arg += ev1[p] * t1;
                sv[p] = sin(arg);
            }
            double acc = 0.;
#pragma unroll
            for (int q = 0; q < GR_DEGMAX; q++) {
                if (q < a.degmax) {  // uniform
                    const int e = s_ne[q * G + gl];  // edge index: lane e mod G, slot e / G
                    double f = gr_shfl<G>(sv[0], e & (G - 1));
                    if (EPL == 2) {
                        const double f1 = gr_shfl<G>(sv[EPL - 1], e & (G - 1));
                        f = (e >= G) ? f1 : f;
                    }
                    if (q < deg) acc += s_nv[q * G + gl] * f;
                }
            }
            const double om = y[1];
            out[0] = om;
            out[1] = (w_i - sc[0] * om) - sc[1] * acc;
        }
    };

    auto begin_attempt = [&]() -> bool {
        if (++iter > o.maxiters) {
            ret = MCBB_RET_MAXITERS;
            return false;
        }
        dts = dt;
        clipped = false;
        if (dts >= o.t1 - t) {
            dts = o.t1 - t;
            clipped = true;
        }
        if (!clipped && dts <= o.dtmin) {
            ret = MCBB_RET_DTLESSTHANMIN;
            return false;
        }
        return true;
    };

    // results of the finished run; the group goes back to the queue
    auto end_run = [&]() {
        const bool ok = (ret == MCBB_RET_SUCCESS) && isave >= n_t;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            if (fpos[c] < 0) continue;
            const double m2 = a2[c];
            const double mean = ok ? am[c] : NaN;
            const double sd = ok ? sqrt((m2 < 0. ? 0. : m2) / (double)(n_t - 2)) : NaN;
            if (KL)  // the KL measures are binned after the solve (k_group_kl_post): the other groups of the warp would wait
                *reinterpret_cast<double2*>(a.stat + ((size_t)(run - a.run_begin) * a.n_dim +
                                                      (SYS == MCBB_SYS_SECOND_ORDER_KURAMOTO ? c * N + gl : gl)) * 2) = make_double2(mean, sd);
            for (int m = 0; m < a.n_meas; m++)
                if (!KL || a.meas[m] != MCBB_MEAS_KL_HIST)
                    a.io.feat[((long long)m * a.n_filter + fpos[c]) * n_mc + run] = (a.meas[m] == MCBB_MEAS_MEAN) ? mean : sd;
        }
        if (gl == 0) {
            if (KL) a.okflag[run - a.run_begin] = ok ? 1 : 0;
            if (a.io.retcode) a.io.retcode[run] = ret;
            if (a.io.nsteps) {
                a.io.nsteps[run] = natt;
                a.io.nsteps[n_mc + run] = nrej;
            }
        }
        mode = MODE_FETCH;
    };

    // one saved sample of this lane's components (x[c]), index isave
    auto consume = [&](const double (&x)[NC]) {
        if (isave >= 1) {  // the first saved sample is dropped, eval_mc_prob.jl:152
            const double rk = 1. / (double)isave;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                double xv = x[c];
                if (a.cyclic) {  // _cyclic_setback!, eval_mc_prob.jl:286-296
                    if (isave == 1) {
                        const double twopi = 6.283185307179586, pi = 3.141592653589793;
                        double v = julia_mod_pos(xv, twopi);
                        if (v > pi) v -= twopi;
                        as_[c] = xv - v;
                        xv = v;
                    } else {
                        xv = xv - as_[c];
                    }
                }
                if (KL && own) smp[(long long)isave * a.n_dim + (SYS == MCBB_SYS_SECOND_ORDER_KURAMOTO ? c * N + gl : gl)] = xv;
                const double mean = am[c];
                const double delta = xv - mean;
                if (fabs(delta) <= 1.7976931348623157e308) {
                    const double mnew = fma(delta, rk, mean);
                    am[c] = mnew;
                    a2[c] = fma(delta, xv - mnew, a2[c]);
                } else {
                    am[c] = nonfinite_mean(mean, xv);
                    a2[c] = NaN;
                }
            }
        }
        isave++;
    };

    const int init_evals = (o.dt_user > 0) ? 1 : 2;
    for (;;) {
        // refill: the first lane of a group in FETCH mode takes the next run; the shuffle runs warp-wide (convergent)
        long long r = 0;
        if (mode == MODE_FETCH && gl == 0) r = (long long)atomicAdd(a.io.work_counter, 1ULL);
        r = __shfl_sync(0xffffffffu, r, 0, G);
        if (mode == MODE_FETCH) {
            run = r;
            if (run >= a.run_end) {
                mode = MODE_DONE;
            } else {
                if (KL) smp = a.smp + (size_t)(run - a.run_begin) * n_t * a.n_dim;
#pragma unroll
                for (int q = 0; q < MCBB_N_SCALARS; q++) sc[q] = a.scalars[q];
                for (int p = 0; p < a.n_par; p++) sc[a.par_slot[p]] = a.io.par[(long long)p * n_mc + run];
                isave = 0;
                ret = 0;
                t = o.t0;
                iter = 0;
                nrej = 0;
                natt = 0;
                qold = o.qoldinit;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    u[c] = own ? a.io.ic[(long long)(c * N + gl) * n_mc + run] : 0.;
                    am[c] = a2[c] = as_[c] = 0.;
                }
                mode = MODE_INIT;
            }
        }
        if (__all_sync(0xffffffffu, mode == MODE_DONE)) break;

        // ---- one round of six right-hand-side evaluations, all groups of the warp stage-aligned ----------------------
        const bool any_init = __any_sync(0xffffffffu, mode == MODE_INIT);
#pragma unroll
        for (int s = 0; s < 6; s++) {
            double y[NC], out[NC];
            if (mode == MODE_RUN) {
#pragma unroll
                for (int c = 0; c < NC; c++) {  // stage s + 2 input: u + dt * sum_j a[s][j] k_j
                    double lin = TS_A(s, 0) * k[0][c];
#pragma unroll
                    for (int j = 1; j <= s; j++) lin = fma(TS_A(s, j), k[j][c], lin);
                    tmp[c] = fma(dts, lin, u[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < NC; c++) y[c] = (mode == MODE_INIT && s == 0) ? u[c] : tmp[c];
            rhs(y, out);
            const bool idle = (mode != MODE_RUN) && !(mode == MODE_INIT && s < init_evals);
            if (!idle) {
                if (mode == MODE_INIT) {
#pragma unroll
                    for (int c = 0; c < NC; c++) k[s == 0 ? 0 : 1][c] = out[c];
                } else {
#pragma unroll
                    for (int c = 0; c < NC; c++) k[1 + s][c] = out[c];
                }
            }
            // OrdinaryDiffEq initdt.jl (Hairer-Wanner), order 5 — executed by every lane of the warp (the sums shuffle)
            // whenever one of its groups starts a run
            if (s < 2 && any_init) {
                double v0[NC], v1[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const double sk = o.abstol + fabs(u[c]) * o.reltol;
                    if (s == 0) {
                        v0[c] = own ? u[c] / sk : 0.;
                        v1[c] = own ? k[0][c] / sk : 0.;
                    } else {
                        v0[c] = own ? (k[1][c] - k[0][c]) / sk : 0.;
                        v1[c] = 0.;
                    }
                }
                const double r0 = state_sum_sq(v0);
                const double r1 = (s == 0) ? state_sum_sq(v1) : 0.;
                if (mode == MODE_INIT) {
                    if (o.dt_user > 0) {
                        if (s == 0) dt = o.dt_user;
                    } else if (s == 0) {
                        d0 = sqrt(r0 / a.n_dim);
                        d1 = sqrt(r1 / a.n_dim);
                        dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
                        dt0 = fmin(dt0, o.dtmax);
#pragma unroll
                        for (int c = 0; c < NC; c++) tmp[c] = fma(dt0, k[0][c], u[c]);
                    } else {
                        const double d2 = sqrt(r0 / a.n_dim) / dt0;
                        const double dm = fmax(d1, d2);
                        const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10., -(2. + log10(dm)) / 5.);
                        dt = fmin(fmin(100. * dt0, dt1), o.dtmax);
                    }
                }
            }
        }

        // ---- end of round: error estimate (every lane: the sum shuffles), then per-group control ---------------------
        double rr[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            double lin = TS_BT(0) * k[0][c];
#pragma unroll
            for (int j = 1; j < 7; j++) lin = fma(TS_BT(j), k[j][c], lin);
            const double ut = dts * lin;
            rr[c] = own ? ut / (o.abstol + fmax(fabs(u[c]), fabs(tmp[c])) * o.reltol) : 0.;
        }
        const double e2 = state_sum_sq(rr);
        // NaN anywhere in the new state of this group (tmp = the state at the end of the step after stage six)
        bool nan_lane = false;
#pragma unroll
        for (int c = 0; c < NC; c++) nan_lane |= own && (tmp[c] != tmp[c]);
        const unsigned nan_ballot = __ballot_sync(0xffffffffu, nan_lane);
        const bool nan_group = ((nan_ballot >> (tid & 31 & ~(G - 1))) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u))) != 0u;
        // the two powers of the PI controller, EEst^beta1 and qold^beta2, in ONE call: even lanes take the first, odd
        // lanes the second, a shuffle hands both to every lane (same inputs, same function: the same bits as two calls)
        const double EEst = sqrt(e2 / a.n_dim);
        const bool odd = (gl & 1) != 0;
        const double pw = pow(odd ? qold : EEst, odd ? o.beta2 : o.beta1);
        const double pw_e = gr_shfl<G>(pw, 0), pw_q = gr_shfl<G>(pw, 1);
        if (mode == MODE_INIT) {
            mode = MODE_RUN;
            if (!begin_attempt()) end_run();
        } else if (mode == MODE_RUN) {
            natt++;
            double q, q11 = 0.;
            if (EEst == 0.) {
                q = 1. / o.qmax;
            } else {
                q11 = pw_e;
                q = q11 / pw_q;
                q = fmax(1. / o.qmax, fmin(1. / o.qmin, q / o.gamma));
            }
            const bool accept = !(EEst > 1.);
            if (accept) {
                qold = fmax(EEst, o.qoldinit);
                const double tnew = clipped ? o.t1 : t + dts;
                while (isave < n_t && s_ts[isave] <= tnew) {
                    const double ts = s_ts[isave];
                    if (ts == tnew) {
                        consume(tmp);
                    } else {
                        double b[7], x[NC];
                        tsit5_btheta((ts - t) / dts, b);
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            double lin = k[0][c] * b[0];
#pragma unroll
                            for (int j = 1; j < 7; j++) lin = fma(k[j][c], b[j], lin);
                            x[c] = fma(dts, lin, u[c]);
                        }
                        consume(x);
                    }
                }
                t = tnew;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    u[c] = tmp[c];
                    k[0][c] = k[6][c];  // FSAL
                }
                dt = fmin(o.dtmax, fmax(o.dtmin, dts / q));
                if (nan_group) ret = MCBB_RET_UNSTABLE;
            } else {
                nrej++;
                dt = dts / fmin(1. / o.qmin, q11 / o.gamma);
            }
            if (ret != 0 || !(t < o.t1)) {
                end_run();
            } else if (!begin_attempt()) {
                end_run();
            }
        }
    }
}

// KL-hist measures of the runs of one chunk from their saved samples: one thread per (run, state dimension)
__global__ void __launch_bounds__(128) k_group_kl_post(const GroupArgs a, const int* __restrict__ fpos_dim, const long long n_runs) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_runs * a.n_dim) return;
    const long long rl = id / a.n_dim;
    const int d = (int)(id - rl * a.n_dim);
    const int pos = fpos_dim[d];
    if (pos < 0) return;
    const double2 ms = *reinterpret_cast<const double2*>(a.stat + ((size_t)rl * a.n_dim + d) * 2);
    double kl = nan("");
    if (a.okflag[rl])
        kl = um_kl_hist(a.smp + (size_t)rl * a.io.n_t * a.n_dim + d, a.n_dim, a.io.n_t, ms.x, ms.y, a.kl_bins, a.kl_nstd, a.kl_tol, nullptr);
    for (int m = 0; m < a.n_meas; m++)
        if (a.meas[m] == MCBB_MEAS_KL_HIST) a.io.feat[((long long)m * a.n_filter + pos) * a.io.n_mc + a.run_begin + rl] = kl;
}

template <int SYS, int G, int EPL = 1>
static int launch_group_one(const GroupArgs& a, size_t smem, int grid, cudaStream_t st) {
    if (a.need_kl) {
        MCBB_CUDA_OK(cudaFuncSetAttribute(k_solve_group<SYS, G, EPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_solve_group<SYS, G, EPL, true><<<grid, GR_THREADS, smem, st>>>(a);
    } else {
        MCBB_CUDA_OK(cudaFuncSetAttribute(k_solve_group<SYS, G, EPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_solve_group<SYS, G, EPL, false><<<grid, GR_THREADS, smem, st>>>(a);
    }
    MCBB_LAUNCHED();
    return 0;
}

// Returns handled = true when the problem was run by the group kernel.
int launch_group(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F, const SolveIO& io,
                 cudaStream_t st, bool* handled) {
    *handled = false;
    const int N = S.n_osc;
    if (hs == nullptr || tune_get("no_group", 0)) return 0;
    if (S.system != MCBB_SYS_KURAMOTO_NETWORK && S.system != MCBB_SYS_SECOND_ORDER_KURAMOTO && S.system != MCBB_SYS_KURAMOTO)
        return 0;
    if (o.alg != MCBB_ALG_TSIT5 || F.n_global > 0 || F.n_comp != 1 || F.matrix_meas != MCBB_MAT_NONE) return 0;
    if (F.need_kl && (F.kl_bins > UM_MAX_KL_BINS || F.kl_bins < 3)) return 0;
    for (int m = 0; m < F.n_meas; m++)
        if (F.meas[m] != MCBB_MEAS_MEAN && F.meas[m] != MCBB_MEAS_STD && F.meas[m] != MCBB_MEAS_KL_HIST) return 0;
    const bool second = S.system == MCBB_SYS_SECOND_ORDER_KURAMOTO;
    const int E = second ? S.n_edges : 0;
    if (N < 2 || N > 32 || E > 64) return 0;
    const int EPL = E > 32 ? 2 : 1;  // edges per lane
    const int width = std::max(N, EPL == 2 ? 32 : E);
    const int G = width <= 8 ? 8 : (width <= 16 ? 16 : 32);

    GroupArgs a = {};
    a.N = N;
    a.E = E;
    a.n_dim = S.n_dim;
    a.mat = S.mat_a;
    a.vec_a = S.vec_a;
    const int EW = EPL * G;  // edge slots
    std::vector<int> edge_node(2 * EW, -1), node_deg(G, 0), node_edge(GR_DEGMAX * G, 0), fpos(2 * G, -1);
    std::vector<double> edge_val(2 * EW, 0.), node_val(GR_DEGMAX * G, 0.);
    int degmax = 0;
    if (second) {
        for (int e = 0; e < E; e++) {
            int cnt = 0;
            for (int i = 0; i < N; i++) {
                const double w = hs->mat_a[i + (size_t)N * e];
                if (w == 0.) continue;
                if (cnt == 2 || node_deg[i] == GR_DEGMAX) return 0;  // not an edge / a hub: the lane kernel takes it
                edge_node[cnt * EW + e] = i;
                edge_val[cnt * EW + e] = w;
                cnt++;
                node_edge[node_deg[i] * G + i] = e;
                node_val[node_deg[i] * G + i] = w;
                node_deg[i]++;
                degmax = std::max(degmax, node_deg[i]);
            }
        }
    }
    a.degmax = degmax;
    {
        std::vector<int> fh;
        if (F.filter) {
            fh.resize(F.n_filter);
            MCBB_CUDA_OK(cudaMemcpyAsync(fh.data(), F.filter, sizeof(int) * F.n_filter, cudaMemcpyDeviceToHost, st));
            MCBB_CUDA_OK(cudaStreamSynchronize(st));
        }
        for (int ii = 0; ii < F.n_filter; ii++) {
            const int d = F.filter ? fh[ii] - 1 : ii;  // state dimension: theta_0..theta_{N-1}, then omega_0..
            if (d < 0 || d >= S.n_dim) return 0;
            if (fpos[(d / N) * G + (d % N)] >= 0) return 0;  // a dimension listed twice: the lane kernel reports it twice
            fpos[(d / N) * G + (d % N)] = ii;
        }
    }
    int* d_int = (int*)scratch_get(103, sizeof(int) * (size_t)(2 * EW + G + GR_DEGMAX * G + 2 * G + S.n_dim));
    double* d_dbl = (double*)scratch_get(104, sizeof(double) * (size_t)(2 * EW + GR_DEGMAX * G));
    if (!d_int || !d_dbl) MCBB_FAIL("solve: device allocation failed");
    std::vector<int> hi;
    hi.insert(hi.end(), edge_node.begin(), edge_node.end());
    hi.insert(hi.end(), node_deg.begin(), node_deg.end());
    hi.insert(hi.end(), node_edge.begin(), node_edge.end());
    hi.insert(hi.end(), fpos.begin(), fpos.end());
    for (int d = 0; d < S.n_dim; d++) hi.push_back(fpos[(d / N) * G + (d % N)]);  // by state dimension, for the KL post kernel
    std::vector<double> hd;
    hd.insert(hd.end(), edge_val.begin(), edge_val.end());
    hd.insert(hd.end(), node_val.begin(), node_val.end());
    MCBB_CUDA_OK(cudaMemcpyAsync(d_int, hi.data(), sizeof(int) * hi.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_dbl, hd.data(), sizeof(double) * hd.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));  // the staging vectors are stack-owned
    a.edge_node = d_int;
    a.node_deg = d_int + 2 * EW;
    a.node_edge = d_int + 2 * EW + G;
    a.fpos = d_int + 2 * EW + G + GR_DEGMAX * G;
    a.edge_val = d_dbl;
    a.node_val = d_dbl + 2 * EW;
    for (int q = 0; q < MCBB_N_SCALARS; q++) a.scalars[q] = S.scalars[q];
    a.n_par = S.n_par;
    for (int p = 0; p < MCBB_MAX_PAR; p++) a.par_slot[p] = S.par_slot[p];
    a.o = o;
    a.n_meas = F.n_meas;
    for (int m = 0; m < F.n_meas; m++) a.meas[m] = F.meas[m];
    a.n_filter = F.n_filter;
    a.cyclic = F.cyclic;
    a.need_kl = F.need_kl;
    a.kl_bins = F.kl_bins;
    a.kl_nstd = F.kl_nstd;
    a.kl_tol = F.kl_tol;
    a.io = io;
    const bool net = S.system == MCBB_SYS_KURAMOTO_NETWORK;
    const size_t smem = sizeof(double) * (((size_t)io.n_t + 1) / 2 * 2 + (second ? (size_t)GR_DEGMAX * G * 2 : (net ? (size_t)N * N : 0)));
    if (smem > 200 * 1024) return 0;
    int dev = 0, sms = 148;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int per_block = GR_THREADS / G;
    const long long want = (io.n_mc + per_block - 1) / per_block;
    const int per_sm = (int)std::max<long long>(1, std::min<long long>(tune_get("group_per_sm", 4), (long long)(220 * 1024) / (long long)(smem + 1024)));
    const int grid = (int)std::min<long long>(want, (long long)sms * per_sm);
    a.smp = nullptr;
    // KL-hist: the ensemble in chunks whose saved samples fit the free device memory (<= 64 GB; hook "group_post_mb"), each
    // followed by the binning kernel.  Other measure sets: one launch.
    long long chunk = io.n_mc;
    const int* d_fpos_dim = d_int + 2 * EW + G + GR_DEGMAX * G + 2 * G;
    if (F.need_kl) {
        const size_t run_bytes = sizeof(double) * (size_t)io.n_t * (size_t)S.n_dim;
        // (cudaMemGetInfo costs milliseconds: only ensembles whose samples exceed 2 GB ask how much memory is free)
        size_t cap = (size_t)2 << 30;
        if (run_bytes * (size_t)io.n_mc > cap) {
            size_t free_b = 0, total_b = 0;
            MCBB_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
            cap = std::max<size_t>(cap, std::min<size_t>((size_t)64 << 30, free_b / 2 + scratch_size(105)));
        }
        if (const long long mb = tune_get("group_post_mb", 0)) cap = (size_t)mb << 20;
        chunk = std::max<long long>(1, std::min<long long>(io.n_mc, (long long)(cap / run_bytes)));
        a.smp = (double*)scratch_get(105, run_bytes * (size_t)chunk);
        a.stat = (double*)scratch_get(110, sizeof(double) * 2 * (size_t)S.n_dim * (size_t)chunk);
        a.okflag = (int*)scratch_get(111, sizeof(int) * (size_t)chunk);
        if (!a.smp || !a.stat || !a.okflag) MCBB_FAIL("solve: device allocation failed");
    }
    for (long long r0 = 0; r0 < io.n_mc; r0 += chunk) {
        const long long r1 = std::min<long long>(io.n_mc, r0 + chunk);
        a.run_begin = r0;
        a.run_end = r1;
        const unsigned long long c0 = (unsigned long long)r0;
        MCBB_CUDA_OK(cudaMemcpyAsync(io.work_counter, &c0, sizeof(c0), cudaMemcpyHostToDevice, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));  // c0 is stack-owned
        const int g = (int)std::min<long long>((r1 - r0 + per_block - 1) / per_block, (long long)grid);
        int rc;
        if (second && EPL == 2)
            rc = launch_group_one<MCBB_SYS_SECOND_ORDER_KURAMOTO, 32, 2>(a, smem, g, st);
        else if (second)
            rc = G == 8 ? launch_group_one<MCBB_SYS_SECOND_ORDER_KURAMOTO, 8>(a, smem, g, st)
                        : (G == 16 ? launch_group_one<MCBB_SYS_SECOND_ORDER_KURAMOTO, 16>(a, smem, g, st)
                                   : launch_group_one<MCBB_SYS_SECOND_ORDER_KURAMOTO, 32>(a, smem, g, st));
        else if (!net)
            rc = G == 8 ? launch_group_one<MCBB_SYS_KURAMOTO, 8>(a, smem, g, st)
                        : (G == 16 ? launch_group_one<MCBB_SYS_KURAMOTO, 16>(a, smem, g, st)
                                   : launch_group_one<MCBB_SYS_KURAMOTO, 32>(a, smem, g, st));
        else
            rc = G == 8 ? launch_group_one<MCBB_SYS_KURAMOTO_NETWORK, 8>(a, smem, g, st)
                        : (G == 16 ? launch_group_one<MCBB_SYS_KURAMOTO_NETWORK, 16>(a, smem, g, st)
                                   : launch_group_one<MCBB_SYS_KURAMOTO_NETWORK, 32>(a, smem, g, st));
        if (rc) return rc;
        if (F.need_kl) {
            const long long nthr = (r1 - r0) * S.n_dim;
            k_group_kl_post<<<(unsigned)((nthr + 127) / 128), 128, 0, st>>>(a, d_fpos_dim, r1 - r0);
            MCBB_LAUNCHED();
        }
    }
    *handled = true;
    return 0;
}

}  // namespace mcbb
