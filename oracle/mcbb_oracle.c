/*
 * mcbb_oracle.c — CPU ORACLE for the MCBB.jl hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call this file.  The product (libmcbb_b200.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference (pure Julia) cannot run in this image (no julia binary) and its own tests
 * hold no golden vectors or known-answer checks for this path (all seven test scripts end in a literal
 * `true`).  The arithmetic below lives mostly in third-party Julia packages that are NOT under
 * /root/reference; their published algorithms are restated from the pinned versions in Manifest.toml:
 *   OrdinaryDiffEq 5.26.7 (Tsit5 tableau/controller/initial dt/saveat interpolant, RK4, FunctionMap),
 *   DiffEqBase 6.10.1 (error norm, default tolerances), StatsBase 0.32.0 (mean, std, Histogram fit,
 *   kldivergence), Distributions 0.22.0 (Normal pdf), Statistics (cor), Clustering 0.13.3 (dbscan; also
 *   vendored in the reference at src/sparse_clustering.jl:138-209), Julia Base (float ranges, mod).
 * What pins this file instead: analytic known answers, scipy cross-checks at tight tolerance, Tsit5 order
 * conditions, sklearn DBSCAN on tie-free inputs, and committed golden vectors (tests/golden/).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference tree).
 * Formulas are kept LITERAL (no factorisation, no FMA contraction: build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mcbb_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------------ */
/* per-run system context                                                                                   */
/* ------------------------------------------------------------------------------------------------------ */
typedef struct {
    const mcbb_system_desc* d;
    double sc[MCBB_N_SCALARS]; /* scalars after the per-run parameter override (setup_mc_prob.jl:722-762) */
    double* tmp;               /* scratch of length >= n_edges (2nd-order Kuramoto)                        */
} sysctx;

/* src/systems.jl — right-hand sides, literal. */
static void rhs(const sysctx* S, const double* u, double* du, double t) {
    const mcbb_system_desc* d = S->d;
    const int N = d->n_osc;
    (void)t;
    switch (d->system) {
    case MCBB_SYS_LOGISTIC: /* systems.jl:34-36 */
        du[0] = S->sc[0] * u[0] * (1. - u[0]);
        break;
    case MCBB_SYS_KURAMOTO: { /* systems.jl:85-94 */
        const double KN = S->sc[0] / (double)N;
        for (int i = 0; i < N; i++) {
            double acc = 0.;
            for (int j = 0; j < N; j++) acc += sin(u[j] - u[i]);
            acc *= KN;
            acc += d->vec_a[i];
            du[i] = acc;
        }
    } break;
    case MCBB_SYS_KURAMOTO_NETWORK: { /* systems.jl:119-130 */
        const double KN = S->sc[0] / (double)N;
        const double* A = d->mat_a;
        for (int i = 0; i < N; i++) {
            double acc = 0.;
            for (int j = 0; j < N; j++) {
                const double a = A[i + (size_t)N * j];
                if (a != 0) acc += a * sin(u[j] - u[i]);
            }
            acc *= KN;
            acc += d->vec_a[i];
            du[i] = acc;
        }
    } break;
    case MCBB_SYS_SECOND_ORDER_KURAMOTO: { /* systems.jl:198-202 */
        const int E = d->n_edges;
        const double* Inc = d->mat_a; /* N x E */
        double* s = S->tmp;
        for (int e = 0; e < E; e++) { /* incidence' * theta */
            double acc = 0.;
            for (int i = 0; i < N; i++) acc += Inc[i + (size_t)N * e] * u[i];
            s[e] = sin(acc);
        }
        for (int i = 0; i < N; i++) du[i] = u[N + i];
        for (int i = 0; i < N; i++) { /* incidence * sin(.) */
            double acc = 0.;
            for (int e = 0; e < E; e++) acc += Inc[i + (size_t)N * e] * s[e];
            du[N + i] = d->vec_a[i] - S->sc[0] * u[N + i] - S->sc[1] * acc;
        }
    } break;
    case MCBB_SYS_NONLOCAL_KURAMOTO_RING: { /* systems.jl:230-239 */
        const double omega0 = S->sc[0], alpha = S->sc[1], fak = S->sc[2];
        const double* G = d->mat_a; /* G[(k-j)+N-1] */
        for (int k = 0; k < N; k++) {
            double acc = 0.;
            for (int j = 0; j < N; j++) acc -= G[(k - j) + N - 1] * sin(u[k] - u[j] + alpha);
            acc /= fak;
            acc += omega0;
            du[k] = acc;
        }
    } break;
    case MCBB_SYS_STUART_LANDAU: { /* paper/stuartlandau/...standard-config.ipynb cell 1 */
        const double eps = S->sc[0], lambda = S->sc[1], omega = S->sc[2];
        const double epsP1 = eps / (2. * S->sc[3]);
        const double epsP2 = eps / (2. * S->sc[4]);
        const double *A1 = d->mat_a, *A2 = d->mat_b;
        for (int i = 0; i < N; i++) {
            const double x = u[2 * i], y = u[2 * i + 1];
            double attr = 0., rep = 0.;
            for (int j = 0; j < N; j++)
                if (A1[i + (size_t)N * j] != 0) attr += u[2 * j] - x;
            attr *= epsP1;
            for (int j = 0; j < N; j++)
                if (A2[i + (size_t)N * j] != 0) rep += u[2 * j + 1] - y;
            rep *= epsP2;
            const double a = hypot(x, y);
            const double cr = lambda - a * a; /* complex(lambda,omega) - abs(z)^2 */
            du[2 * i] = (cr * x - omega * y) + attr;
            du[2 * i + 1] = (cr * y + omega * x) - rep;
        }
    } break;
    case MCBB_SYS_ROESSLER_NETWORK: { /* systems.jl:286-298, including the du-coupling quirk (:292) */
        const double K = S->sc[0];
        const double* L = d->mat_a;
        for (int i = 0; i < N; i++) {
            const int ii = 3 * i;
            double dx = -u[ii + 1] - u[ii + 2];
            for (int j = 0; j < N; j++) {
                const double l = L[i + (size_t)N * j];
                if (l != 0) dx -= K * l * dx;
            }
            du[ii] = dx;
            du[ii + 1] = u[ii] + d->vec_a[i] * u[ii + 1];
            du[ii + 2] = d->vec_b[i] + u[ii + 2] * (dx - d->vec_c[i]);
        }
    } break;
    default:
        for (int i = 0; i < d->n_dim; i++) du[i] = NAN;
    }
}

/* ------------------------------------------------------------------------------------------------------ */
/* Tsit5 — OrdinaryDiffEq 5.26.7 (third-party; restated)                           */
/* ------------------------------------------------------------------------------------------------------ */
static const double c_a21 = 0.161;
static const double c_a31 = -0.008480655492356989, c_a32 = 0.335480655492357;
static const double c_a41 = 2.8971530571054935, c_a42 = -6.359448489975075, c_a43 = 4.3622954328695815;
static const double c_a51 = 5.325864828439257, c_a52 = -11.748883564062828, c_a53 = 7.4955393428898365,
                    c_a54 = -0.09249506636175525;
static const double c_a61 = 5.86145544294642, c_a62 = -12.92096931784711, c_a63 = 8.159367898576159,
                    c_a64 = -0.071584973281401, c_a65 = -0.028269050394068383;
static const double c_a71 = 0.09646076681806523, c_a72 = 0.01, c_a73 = 0.4798896504144996,
                    c_a74 = 1.379008574103742, c_a75 = -3.290069515436081, c_a76 = 2.324710524099774;
static const double c_bt1 = -0.00178001105222577714, c_bt2 = -0.0008164344596567469,
                    c_bt3 = 0.007880878010261995, c_bt4 = -0.1447110071732629, c_bt5 = 0.5823571654525552,
                    c_bt6 = -0.45808210592918697, c_bt7 = 0.015151515151515152;
static const double c_c[7] = {0., 0.161, 0.327, 0.9, 0.9800255409045097, 1., 1.};

static void tsit5_btheta(double th, double* b) { /* Tsit5 dense-output weights (4th order) */
    const double th2 = th * th;
    b[0] = th * (1. + th * (-2.763706197274826 + th * (2.9132554618219126 + th * -1.0530884977290216)));
    b[1] = th2 * (0.13169999999999998 + th * (-0.2234 + th * 0.1017));
    b[2] = th2 * (3.9302962368947516 + th * (-5.941033872131505 + th * 2.490627285651253));
    b[3] = th2 * (-12.411077166933676 + th * (30.33818863028232 + th * -16.548102889244902));
    b[4] = th2 * (37.50931341651104 + th * (-88.1789048947664 + th * 47.37952196281928));
    b[5] = th2 * (-27.896526289197286 + th * (65.09189467479366 + th * -34.87065786149661));
    b[6] = th2 * (1.5 + th * (-4. + th * 2.5));
}

typedef struct {
    double abstol, reltol, dt_user, dtmin, dtmax, qmin, qmax, gamma, beta1, beta2, qoldinit, t0, t1;
    int64_t maxiters;
    int alg;
} sopts;

static sopts resolve_opts(const mcbb_solver_opts* o) { /* DiffEq defaults, SURVEY §5 / Appendix A.3-A.4 */
    sopts s;
    s.alg = o->alg;
    s.t0 = o->t0;
    s.t1 = o->t1;
    s.abstol = o->abstol > 0 ? o->abstol : 1e-6;
    s.reltol = o->reltol > 0 ? o->reltol : 1e-3;
    s.dt_user = o->dt;
    s.dtmin = o->dtmin > 0 ? o->dtmin : 2.220446049250313e-16;
    s.dtmax = o->dtmax > 0 ? o->dtmax : (o->t1 - o->t0);
    s.maxiters = o->maxiters > 0 ? o->maxiters : 100000;
    s.qmin = o->qmin > 0 ? o->qmin : 0.2;
    s.qmax = o->qmax > 0 ? o->qmax : 10.;
    s.gamma = o->gamma > 0 ? o->gamma : 0.9;
    s.beta1 = o->beta1 > 0 ? o->beta1 : 7. / 50.;
    s.beta2 = o->beta2 > 0 ? o->beta2 : 2. / 25.;
    s.qoldinit = o->qoldinit > 0 ? o->qoldinit : 1e-4;
    return s;
}

/* DiffEqBase.ODE_DEFAULT_NORM: sqrt(sum(abs2, x) / length(x)) over the ELEMENTS of the state array.  For a
 * Complex{Float64} state (Stuart-Landau) an element is one complex number: the scale of element i is
 * abstol + abs(z_i) * reltol with the complex modulus (ODE_DEFAULT_NORM(z::Complex) = abs(z)), both components are
 * divided by it, and length(x) = N — not 2N.  x and sk are indexed by real component; sk[2i] == sk[2i+1] then. */
static int is_complex_state(const sysctx* S) { return S->d->system == MCBB_SYS_STUART_LANDAU; }

static double rms_scaled(const double* x, const double* sk, int n, int cplx) {
    double acc = 0.;
    for (int i = 0; i < n; i++) {
        const double v = x[i] / sk[i];
        acc += v * v;
    }
    return sqrt(acc / (cplx ? n / 2 : n));
}

/* element-wise magnitude used in the error scale: |x| for reals, the modulus of the complex element otherwise */
static double elem_abs(const double* u, int i, int cplx) {
    return cplx ? hypot(u[i & ~1], u[i | 1]) : fabs(u[i]);
}

/* OrdinaryDiffEq initdt.jl (Hairer-Wanner), order 5.  f0 = f(u0,t0) is passed in. */
static double tsit5_initdt(const sysctx* S, const sopts* o, const double* u0, const double* f0, int n,
                           double* w1, double* w2, double* w3) {
    const int cplx = is_complex_state(S);
    const int ne = cplx ? n / 2 : n;
    double* sk = w1;
    for (int i = 0; i < n; i++) sk[i] = o->abstol + elem_abs(u0, i, cplx) * o->reltol;
    const double d0 = rms_scaled(u0, sk, n, cplx);
    const double d1 = rms_scaled(f0, sk, n, cplx);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    dt0 = fmin(dt0, o->dtmax);
    double* u1 = w2;
    for (int i = 0; i < n; i++) u1[i] = u0[i] + dt0 * f0[i];
    double* f1 = w3;
    rhs(S, u1, f1, o->t0 + dt0);
    double acc = 0.;
    for (int i = 0; i < n; i++) {
        const double v = (f1[i] - f0[i]) / sk[i];
        acc += v * v;
    }
    const double d2 = sqrt(acc / ne) / dt0;
    const double dm = fmax(d1, d2);
    double dt1;
    if (dm <= 1e-15)
        dt1 = fmax(1e-6, dt0 * 1e-3);
    else
        dt1 = pow(10., -(2. + log10(dm)) / 5.);
    return fmin(fmin(100. * dt0, dt1), o->dtmax);
}

/* Integrate one trajectory and fill saved[n x nt] (column it holds the state at t_save[it]).
 * Returns the retcode; *n_saved = number of columns actually written (all nt on success). */
static int integrate(const sysctx* S, const sopts* o, const double* u0, int n, const double* tsave, int nt,
                     double* saved, int64_t* nsteps, int64_t* nrej, int* n_saved, double* work /* 12n */) {
    const int cplx = is_complex_state(S);
    double* u = work;
    double* k[7];
    for (int s = 0; s < 7; s++) k[s] = work + (size_t)(1 + s) * n;
    double* tmp = work + (size_t)8 * n;
    double* unew = work + (size_t)9 * n;
    double* w3 = work + (size_t)10 * n;
    double* w4 = work + (size_t)11 * n;
    int isave = 0;
    int ret = MCBB_RET_SUCCESS;
    *nsteps = 0;
    *nrej = 0;
    memcpy(u, u0, sizeof(double) * n);
    double t = o->t0;
    const double t1 = o->t1;

    if (o->alg == MCBB_ALG_FUNCTIONMAP) { /* OrdinaryDiffEq FunctionMap: u_{n+1} = f(u_n, p, t_n), dt = 1 */
        /* savestart=false: t0 itself is never saved */
        while (t < t1) {
            rhs(S, u, unew, t);
            t += 1.;
            memcpy(u, unew, sizeof(double) * n);
            (*nsteps)++;
            while (isave < nt && tsave[isave] <= t) {
                memcpy(saved + (size_t)isave * n, u, sizeof(double) * n); /* integer grid: ts == t */
                isave++;
            }
            int bad = 0;
            for (int i = 0; i < n; i++) bad |= isnan(u[i]);
            if (bad) {
                ret = MCBB_RET_UNSTABLE;
                break;
            }
        }
        *n_saved = isave;
        return ret;
    }

    if (o->alg == MCBB_ALG_RK4) { /* classical RK4, fixed dt; saveat via cubic Hermite (u,f at both ends) */
        const double h = o->dt_user;
        rhs(S, u, k[0], t);
        int64_t it = 0;
        while (t < t1) {
            if (++it > o->maxiters) {
                ret = MCBB_RET_MAXITERS;
                break;
            }
            double dt = h;
            int clipped = 0;
            if (t + dt >= t1 || fabs((t + dt) - t1) < 100. * 2.220446049250313e-16 * fabs(t1)) {
                dt = t1 - t;
                clipped = 1;
            }
            for (int i = 0; i < n; i++) tmp[i] = u[i] + (dt / 2.) * k[0][i];
            rhs(S, tmp, k[1], t + dt / 2.);
            for (int i = 0; i < n; i++) tmp[i] = u[i] + (dt / 2.) * k[1][i];
            rhs(S, tmp, k[2], t + dt / 2.);
            for (int i = 0; i < n; i++) tmp[i] = u[i] + dt * k[2][i];
            rhs(S, tmp, k[3], t + dt);
            for (int i = 0; i < n; i++)
                unew[i] = u[i] + (dt / 6.) * (2. * (k[1][i] + k[2][i]) + (k[0][i] + k[3][i]));
            rhs(S, unew, k[4], t + dt); /* f at the new point (FSAL for the next step, Hermite end slope) */
            const double tnew = clipped ? t1 : t + dt;
            (*nsteps)++;
            while (isave < nt && tsave[isave] <= tnew) {
                const double ts = tsave[isave];
                double* out = saved + (size_t)isave * n;
                if (ts == tnew) {
                    memcpy(out, unew, sizeof(double) * n);
                } else {
                    const double th = (ts - t) / dt;
                    for (int i = 0; i < n; i++) /* hermite_interpolant, OrdinaryDiffEq dense/generic_dense.jl */
                        out[i] = (1. - th) * u[i] + th * unew[i] +
                                 th * (th - 1.) * ((1. - 2. * th) * (unew[i] - u[i]) + (th - 1.) * dt * k[0][i] +
                                                   th * dt * k[4][i]);
                }
                isave++;
            }
            t = tnew;
            memcpy(u, unew, sizeof(double) * n);
            memcpy(k[0], k[4], sizeof(double) * n);
            int bad = 0;
            for (int i = 0; i < n; i++) bad |= isnan(u[i]);
            if (bad) {
                ret = MCBB_RET_UNSTABLE;
                break;
            }
        }
        *n_saved = isave;
        return ret;
    }

    /* ---- Tsit5 adaptive ---- */
    rhs(S, u, k[0], t); /* fsalfirst */
    double dt = o->dt_user > 0 ? o->dt_user : tsit5_initdt(S, o, u, k[0], n, tmp, unew, w3);
    double qold = o->qoldinit;
    int64_t iter = 0;
    (void)w4;
    while (t < t1) {
        if (++iter > o->maxiters) {
            ret = MCBB_RET_MAXITERS;
            break;
        }
        /* modify_dt_for_tstops!: the only tstop is t1 */
        int clipped = 0;
        double dts = dt;
        if (dts >= t1 - t) {
            dts = t1 - t;
            clipped = 1;
        }
        if (!clipped && dts <= o->dtmin) {
            ret = MCBB_RET_DTLESSTHANMIN;
            break;
        }
        /* perform_step! (tsit_perform_step.jl) */
        for (int i = 0; i < n; i++) tmp[i] = u[i] + dts * (c_a21 * k[0][i]);
        rhs(S, tmp, k[1], t + c_c[1] * dts);
        for (int i = 0; i < n; i++) tmp[i] = u[i] + dts * (c_a31 * k[0][i] + c_a32 * k[1][i]);
        rhs(S, tmp, k[2], t + c_c[2] * dts);
        for (int i = 0; i < n; i++) tmp[i] = u[i] + dts * (c_a41 * k[0][i] + c_a42 * k[1][i] + c_a43 * k[2][i]);
        rhs(S, tmp, k[3], t + c_c[3] * dts);
        for (int i = 0; i < n; i++)
            tmp[i] = u[i] + dts * (c_a51 * k[0][i] + c_a52 * k[1][i] + c_a53 * k[2][i] + c_a54 * k[3][i]);
        rhs(S, tmp, k[4], t + c_c[4] * dts);
        for (int i = 0; i < n; i++)
            tmp[i] = u[i] + dts * (c_a61 * k[0][i] + c_a62 * k[1][i] + c_a63 * k[2][i] + c_a64 * k[3][i] +
                                   c_a65 * k[4][i]);
        rhs(S, tmp, k[5], t + dts);
        for (int i = 0; i < n; i++)
            unew[i] = u[i] + dts * (c_a71 * k[0][i] + c_a72 * k[1][i] + c_a73 * k[2][i] + c_a74 * k[3][i] +
                                    c_a75 * k[4][i] + c_a76 * k[5][i]);
        rhs(S, unew, k[6], t + dts);
        double acc = 0.;
        for (int i = 0; i < n; i++) {
            const double ut = dts * (c_bt1 * k[0][i] + c_bt2 * k[1][i] + c_bt3 * k[2][i] + c_bt4 * k[3][i] +
                                     c_bt5 * k[4][i] + c_bt6 * k[5][i] + c_bt7 * k[6][i]);
            /* calculate_residuals: ut / (abstol + max(norm(u_i), norm(unew_i)) * reltol), element-wise norm */
            const double r = ut / (o->abstol + fmax(elem_abs(u, i, cplx), elem_abs(unew, i, cplx)) * o->reltol);
            acc += r * r;
        }
        const double EEst = sqrt(acc / (cplx ? n / 2 : n));
        (*nsteps)++;
        /* stepsize_controller! (standard PI) */
        double q, q11 = 0.;
        if (EEst == 0.) {
            q = 1. / o->qmax;
        } else {
            q11 = pow(EEst, o->beta1);
            q = q11 / pow(qold, o->beta2);
            q = fmax(1. / o->qmax, fmin(1. / o->qmin, q / o->gamma));
            if (q <= 1. && q >= 1.) q = 1.; /* qsteady_min = qsteady_max = 1 */
        }
        if (!(EEst > 1.)) { /* accept (a NaN EEst is accepted, then caught by the unstable check) */
            qold = fmax(EEst, o->qoldinit);
            const double tnew = clipped ? t1 : t + dts;
            double b[7];
            while (isave < nt && tsave[isave] <= tnew) { /* savevalues!: saveat through the interpolant */
                const double ts = tsave[isave];
                double* out = saved + (size_t)isave * n;
                if (ts == tnew) {
                    memcpy(out, unew, sizeof(double) * n);
                } else {
                    const double th = (ts - t) / dts;
                    tsit5_btheta(th, b);
                    for (int i = 0; i < n; i++)
                        out[i] = u[i] + dts * (k[0][i] * b[0] + k[1][i] * b[1] + k[2][i] * b[2] + k[3][i] * b[3] +
                                               k[4][i] * b[4] + k[5][i] * b[5] + k[6][i] * b[6]);
                }
                isave++;
            }
            double dtnew = dts / q;
            dtnew = fmin(o->dtmax, fmax(o->dtmin, dtnew));
            t = tnew;
            memcpy(u, unew, sizeof(double) * n);
            memcpy(k[0], k[6], sizeof(double) * n); /* FSAL */
            dt = dtnew;
            int bad = 0;
            for (int i = 0; i < n; i++) bad |= isnan(u[i]);
            if (bad) {
                ret = MCBB_RET_UNSTABLE;
                break;
            }
        } else { /* step_reject_controller! */
            (*nrej)++;
            dt = dts / fmin(1. / o->qmin, q11 / o->gamma);
        }
    }
    *n_saved = isave;
    return ret;
}

/* ------------------------------------------------------------------------------------------------------ */
/* featurizer — src/eval_mc_prob.jl                                                                        */
/* ------------------------------------------------------------------------------------------------------ */

/* Julia Base mod(x, y) for floats (floored modulo built on the exact rem). */
static double julia_mod(double x, double y) {
    double r = fmod(x, y);
    if (r == 0.) return copysign(r, y);
    if ((r > 0.) != (y > 0.)) return r + y;
    return r;
}

/* _cyclic_setback!, eval_mc_prob.jl:286-296 */
static void cyclic_setback(double* a, int n) {
    const double twopi = 6.283185307179586; /* 2pi = 2*Float64(pi) */
    const double pi = 3.141592653589793;
    const double a1 = a[0];
    double v = julia_mod(a[0], twopi);
    if (v > pi) v -= twopi;
    a[0] = v;
    const double subtr = a1 - v;
    for (int i = 1; i < n; i++) a[i] = a[i] - subtr;
}


/* ---- Julia Base (:)(start, step, stop) for Float64 (base/twiceprecision.jl): length and elements.  First the
 * "rational lift" (exact arithmetic when start, step, stop are ratios of small integers, e.g. 0.0:0.1:0.7),
 * then the literal fallback. ---- */
static void jl_rat(double x, long long* num, long long* den) {
    double y = x;
    long long a = 1, d = 1, b = 0, c = 0;
    const double m = 16777216.0; /* maxintfloat(Float32) */
    while (fabs(y) <= m) {
        const long long f = (long long)y; /* trunc */
        y -= (double)f;
        const long long a2 = f * a + c, b2 = f * b + d;
        c = a;
        d = b;
        a = a2;
        b = b2;
        const long long aa = a < 0 ? -a : a, bb = b < 0 ? -b : b;
        if ((aa > bb ? aa : bb) > 16777216LL) {
            *num = c;
            *den = d;
            return;
        }
        if ((double)a / (double)b == x) break;
        y = 1. / y;
    }
    *num = a;
    *den = b;
}
static long long jl_gcd(long long a, long long b) {
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    while (b) {
        const long long t = a % b;
        a = b;
        b = t;
    }
    return a;
}
static int jl_between(double a, double x, double b) { return (a <= x && x <= b) || (b <= x && x <= a); }
/* returns the length; when out != NULL writes the elements */
static int julia_range(double start, double step, double stop, double* out, int cap) {
    long long sn, sd, an, ad, bn, bd;
    jl_rat(step, &sn, &sd);
    if (sd != 0 && (double)sn / (double)sd == step) {
        jl_rat(start, &an, &ad);
        jl_rat(stop, &bn, &bd);
        if (ad != 0 && bd != 0 && (double)an / (double)ad == start && (double)bn / (double)bd == stop) {
            const long long g = jl_gcd(ad, sd);
            const long long den = g ? (ad / g) * sd : 0;
            const double mx = 9007199254740992.0;
            if (den != 0) {
                if (fabs(start * (double)den) <= mx && fabs(step * (double)den) <= mx && den % ad == 0 && den % sd == 0) {
                    const long long start_n = llround(start * (double)den), step_n = llround(step * (double)den);
                    long long len = (den * bn - bd * start_n + step_n * bd) / (step_n * bd);
                    if (len < 0) len = 0;
                    if (jl_between(start, start + (double)(len - 1) * step, stop + step / 2) &&
                        !jl_between(start, start + (double)len * step, stop)) {
                        if (out)
                            for (long long i = 0; i < len && i < cap; i++)
                                out[i] = (double)((long double)(start_n + i * step_n) / (long double)den);
                        return (int)len;
                    }
                }
            }
        }
    }
    const double lf = (stop - start) / step;
    int len;
    if (lf < 0)
        len = 0;
    else if (lf == 0)
        len = 1;
    else {
        len = (int)nearbyint(lf) + 1;
        const double stop2 = start + (len - 1) * step;
        if ((start < stop && stop < stop2) || (start > stop && stop > stop2)) len -= 1;
    }
    if (out)
        for (int i = 0; i < len && i < cap; i++) out[i] = (double)((long double)start + (long double)i * (long double)step);
    return len;
}

/* Length only (used for the KL bin centres, whose start/step are generic floats: the lift never applies). */
static int julia_range_len(double start, double step, double stop) {
    return julia_range(start, step, stop, NULL, 0);
}

/* Base.searchsortedlast on a Vector (binary search as in sort.jl; edges may be unsorted — quirk B.2) */
static int searchsortedlast_vec(const double* v, int n, double x) {
    int lo = 0, hi = n + 1;
    while (lo < hi - 1) {
        const int m = (int)(((unsigned)lo + (unsigned)hi) >> 1);
        if (x < v[m - 1])
            hi = m;
        else
            lo = m;
    }
    return lo;
}

/* empirical_1D_KL_divergence_hist, eval_mc_prob.jl:310-337.
 * force_len: 0 = Julia's own range length; >0 = force that many bin centres (test hook for the
 * 1-ulp-sensitive length decision, see DESIGN.md "KL range-length quirk"). */
double orc_kl_hist(const double* u, int n, double mu, double sig, int hist_bins, double n_stds, double sig_tol,
                   int force_len) {
    if (sig < sig_tol) return 0.;
    if (hist_bins % 2 == 0) hist_bins += 1;
    const double k = (hist_bins - 1) / 2.;
    const double bw = (n_stds * sig) / k;
    const double start = mu - k * bw, stop = mu + k * bw;
    int nc = force_len > 0 ? force_len : julia_range_len(start, bw, stop);
    if (nc < 1) return NAN;
    double centers[128], edges[130], wts[130], ref[130];
    if (nc > 127) nc = 127;
    for (int i = 0; i < nc; i++) {
        /* StepRangeLen in twice precision: element i is the correctly rounded start + i*step */
        centers[i] = (double)((long double)start + (long double)i * (long double)bw);
        edges[i] = (double)(((long double)start + (long double)i * (long double)bw) - (long double)(bw / 2.));
    }
    edges[nc] = k * bw + bw / 2.; /* push!(bin_edges, k*bin_width + bin_width/2.) — `mu` missing, :328 */
    const int nb = nc;            /* bins = edges - 1 */
    for (int b = 0; b < nb; b++) wts[b] = 0.;
    for (int i = 0; i < n; i++) { /* fit(Histogram, u, weights, edges; closed=:left) */
        const int idx = searchsortedlast_vec(edges, nb + 1, u[i]);
        if (idx >= 1 && idx <= nb) wts[idx - 1] += 1.;
    }
    double tot = 0.;
    for (int b = 0; b < nb; b++) tot += wts[b];
    for (int b = 0; b < nb; b++) wts[b] /= tot; /* normalize!(mode=:probability) */
    double rs = 0.;
    const double invsqrt2pi = 0.3989422804014327;
    for (int b = 0; b < nb; b++) { /* pdf.(Normal(mu,sig), centers): StatsFuns normpdf */
        const double z = (centers[b] - mu) / sig;
        ref[b] = exp(-(z * z) / 2.) * invsqrt2pi / sig;
        rs += ref[b];
    }
    for (int b = 0; b < nb; b++) ref[b] /= rs;
    double kl = 0.; /* StatsBase.kldivergence */
    for (int b = 0; b < nb; b++)
        if (wts[b] > 0) kl += wts[b] * log(wts[b] / ref[b]);
    return kl;
}

static sopts g_dummy;

/* Base.searchsortedlast on a float range (used for the 0:1/nbins:1 edges of correlation_hist) */
static int searchsortedlast_range(double first, double step, int len, double x) {
    double f = (x - first) / step + 1.;
    if (f < 1.) f = 1.;
    if (f > len) f = len;
    const int n = (int)nearbyint(f);
    const double an = (double)((long double)first + (long double)(n - 1) * (long double)step);
    return x < an ? n - 1 : n;
}

/* correlation_ecdf, eval_mc_prob.jl:462-488 (real states).  X is nf x ns (sample index slowest). */
static void corr_ecdf(const double* X, int nf, int ns, int nbins, int raw_weights, double* out) {
    double* mean = (double*)malloc(sizeof(double) * nf);
    double* C = (double*)malloc(sizeof(double) * nf * nf);
    for (int i = 0; i < nf; i++) {
        double s = 0.;
        for (int t = 0; t < ns; t++) s += X[i + (size_t)nf * t];
        mean[i] = s / ns;
    }
    for (int i = 0; i < nf; i++)
        for (int j = 0; j <= i; j++) {
            double s = 0.;
            for (int t = 0; t < ns; t++) s += (X[i + (size_t)nf * t] - mean[i]) * (X[j + (size_t)nf * t] - mean[j]);
            C[i + nf * j] = C[j + nf * i] = s;
        }
    int64_t* h = (int64_t*)calloc(nbins, sizeof(int64_t));
    const double step = 1. / nbins;
    for (int j = 0; j < nf; j++)
        for (int i = 0; i < nf; i++) {
            double c;
            if (i == j)
                c = 1.;
            else {
                c = C[i + nf * j] / (sqrt(C[i + nf * i]) * sqrt(C[j + nf * j]));
                if (c > 1.) c = 1.;
                if (c < -1.) c = -1.;
            }
            c = fabs(c);
            if (isnan(c)) continue; /* Julia would throw on NaN here; the oracle drops it */
            const int idx = searchsortedlast_range(0., step, nbins + 1, c);
            if (idx >= 1 && idx <= nbins) h[idx - 1]++;
        }
    if (raw_weights) { /* correlation_hist returns hist.weights, eval_mc_prob.jl:472-481 */
        for (int b = 0; b < nbins; b++) out[b] = (double)h[b];
        free(h);
        free(C);
        free(mean);
        return;
    }
    double cum = 0.;
    for (int b = 0; b < nbins; b++) { /* ecdf_hist, eval_clustering.jl:759-762 */
        cum += (double)h[b];
        out[b] = cum;
    }
    const double last = out[nbins - 1];
    for (int b = 0; b < nbins; b++) out[b] = out[b] / last;
    free(h);
    free(C);
    free(mean);
}

/* correlation_hist / correlation_ecdf of COMPLEX states (the Stuart-Landau configuration): Statistics.cor(x, dims=2)
 * centres each row, forms x*x' (second factor conjugated) and divides by the root of the diagonal; clampcor is the
 * identity for complex values; then abs, the same histogram and ecdf_hist.  Z holds interleaved (re, im), nf x ns. */
static void corr_ecdf_complex(const double* Z, int nf, int ns, int nbins, int raw_weights, double* out) {
    double* mre = (double*)calloc(nf, sizeof(double));
    double* mim = (double*)calloc(nf, sizeof(double));
    double* var = (double*)calloc(nf, sizeof(double));
    for (int i = 0; i < nf; i++) {
        double sr = 0., si = 0.;
        for (int t = 0; t < ns; t++) {
            sr += Z[2 * i + (size_t)2 * nf * t];
            si += Z[2 * i + 1 + (size_t)2 * nf * t];
        }
        mre[i] = sr / ns;
        mim[i] = si / ns;
        double v = 0.;
        for (int t = 0; t < ns; t++) {
            const double a = Z[2 * i + (size_t)2 * nf * t] - mre[i], b = Z[2 * i + 1 + (size_t)2 * nf * t] - mim[i];
            v += a * a + b * b;
        }
        var[i] = v;
    }
    int64_t* h = (int64_t*)calloc(nbins, sizeof(int64_t));
    const double step = 1. / nbins;
    for (int j = 0; j < nf; j++)
        for (int i = 0; i < nf; i++) {
            if (i == j) continue; /* the unit diagonal lies outside [0, 1) */
            double re = 0., im = 0.;
            for (int t = 0; t < ns; t++) {
                const double a = Z[2 * i + (size_t)2 * nf * t] - mre[i], b = Z[2 * i + 1 + (size_t)2 * nf * t] - mim[i];
                const double c = Z[2 * j + (size_t)2 * nf * t] - mre[j], d = Z[2 * j + 1 + (size_t)2 * nf * t] - mim[j];
                re += a * c + b * d; /* (a + ib)(c - id) */
                im += b * c - a * d;
            }
            const double den = sqrt(var[i]) * sqrt(var[j]);
            const double c = hypot(re / den, im / den);
            if (isnan(c)) continue;
            const int idx = searchsortedlast_range(0., step, nbins + 1, c);
            if (idx >= 1 && idx <= nbins) h[idx - 1]++;
        }
    double cum = 0., tot = 0.;
    for (int b = 0; b < nbins; b++) tot += (double)h[b];
    for (int b = 0; b < nbins; b++) {
        cum += (double)h[b];
        out[b] = raw_weights ? (double)h[b] : cum / tot;
    }
    free(h);
    free(var);
    free(mim);
    free(mre);
}

/* eval_ode_run, eval_mc_prob.jl:75-188, on a saved array (n_dim x nt).  Output strides: per-run values are
 * written at feat[(m*nf + d)*stride + off] so the caller can lay runs out column-major. */
static void featurize(const mcbb_system_desc* sys, const mcbb_feat_opts* fo, const double* saved, int n_dim,
                      int nt, double* feat, double* mat, double* glob, size_t stride, size_t off, double* sol_i,
                      int kl_force_len) {
    const int cplx = sys->system == MCBB_SYS_STUART_LANDAU;
    const int n_state = cplx ? n_dim / 2 : n_dim;
    const int nf = fo->n_filter > 0 ? fo->n_filter : n_state;
    const int ns = nt - 1; /* sol[i_dim, 2:end] — first saved sample dropped, :152 */
    const int kl_bins = fo->kl_bins > 0 ? fo->kl_bins : 31;
    const double kl_nstd = fo->kl_n_stds > 0 ? fo->kl_n_stds : 3.;
    const double kl_tol = fo->kl_sig_tol > 0 ? fo->kl_sig_tol : 1e-4;
    for (int ii = 0; ii < nf; ii++) {
        const int sidx = fo->n_filter > 0 ? fo->state_filter[ii] - 1 : ii;
        double last_mean[2] = {NAN, NAN}, last_std[2] = {NAN, NAN};
        int cached_comp = -1;
        for (int m = 0; m < fo->n_meas; m++) {
            const int comp = cplx ? fo->meas_comp[m] : 0;
            if (comp != cached_comp) { /* real.(arr) / imag.(arr) views share the setback per component */
                const int d = cplx ? 2 * sidx + comp : sidx;
                for (int t = 0; t < ns; t++) sol_i[t] = saved[d + (size_t)n_dim * (t + 1)];
                if (fo->cyclic_setback) cyclic_setback(sol_i, ns);
                cached_comp = comp;
            }
            double val = NAN;
            switch (fo->meas[m]) {
            case MCBB_MEAS_MEAN: {
                double s = 0.;
                for (int t = 0; t < ns; t++) s += sol_i[t];
                val = s / ns;
                last_mean[comp] = val;
            } break;
            case MCBB_MEAS_STD: { /* std(u; mean=m, corrected=true) */
                const double mu = last_mean[comp];
                double s = 0.;
                for (int t = 0; t < ns; t++) s += (sol_i[t] - mu) * (sol_i[t] - mu);
                val = sqrt(s / (ns - 1));
                last_std[comp] = val;
            } break;
            case MCBB_MEAS_KL_HIST:
                val = orc_kl_hist(sol_i, ns, last_mean[comp], last_std[comp], kl_bins, kl_nstd, kl_tol,
                                  kl_force_len);
                break;
            }
            feat[((size_t)m * nf + ii) * stride + off] = val;
        }
    }
    if (fo->n_global > 0 && glob) {
        for (int g = 0; g < fo->n_global; g++) {
            double s = 0.;
            for (int t = 1; t < nt; t++)
                for (int ii = 0; ii < nf; ii++) {
                    const int sidx = fo->n_filter > 0 ? fo->state_filter[ii] - 1 : ii;
                    s += saved[sidx + (size_t)n_dim * t];
                }
            glob[(size_t)g * stride + off] = s;
        }
    }
    if (fo->matrix_meas != MCBB_MAT_NONE && mat && cplx) {
        const int nb = fo->corr_nbins > 0 ? fo->corr_nbins : 30;
        double* Z = (double*)malloc(sizeof(double) * 2 * nf * ns);
        double* e = (double*)malloc(sizeof(double) * nb);
        for (int t = 0; t < ns; t++)
            for (int ii = 0; ii < nf; ii++) {
                const int sidx = fo->n_filter > 0 ? fo->state_filter[ii] - 1 : ii;
                Z[2 * ii + (size_t)2 * nf * t] = saved[2 * sidx + (size_t)n_dim * (t + 1)];
                Z[2 * ii + 1 + (size_t)2 * nf * t] = saved[2 * sidx + 1 + (size_t)n_dim * (t + 1)];
            }
        corr_ecdf_complex(Z, nf, ns, nb, fo->matrix_meas == MCBB_MAT_CORR_HIST, e);
        for (int b = 0; b < nb; b++) mat[(size_t)b * stride + off] = e[b];
        free(e);
        free(Z);
    }
    if (fo->matrix_meas != MCBB_MAT_NONE && mat && !cplx) {
        const int nb = fo->corr_nbins > 0 ? fo->corr_nbins : 30;
        double* X = (double*)malloc(sizeof(double) * nf * ns);
        double* e = (double*)malloc(sizeof(double) * nb);
        for (int t = 0; t < ns; t++)
            for (int ii = 0; ii < nf; ii++) {
                const int sidx = fo->n_filter > 0 ? fo->state_filter[ii] - 1 : ii;
                X[ii + (size_t)nf * t] = saved[sidx + (size_t)n_dim * (t + 1)];
            }
        corr_ecdf(X, nf, ns, nb, fo->matrix_meas == MCBB_MAT_CORR_HIST, e);
        for (int b = 0; b < nb; b++) mat[(size_t)b * stride + off] = e[b];
        free(e);
        free(X);
    }
}

/* ------------------------------------------------------------------------------------------------------ */
/* exported entry points                                                                                    */
/* ------------------------------------------------------------------------------------------------------ */

/* tsave_array, setup_mc_prob.jl:1166-1180 */
int orc_tsave_array(const mcbb_solver_opts* o, int32_t n_t, double rel_transient_time, double* t_save,
                    int32_t* n_out) {
    (void)g_dummy;
    if (o->alg == MCBB_ALG_FUNCTIONMAP) {
        const double tot = nearbyint((o->t1 - o->t0) * (1 - rel_transient_time));
        const double start = o->t1 - tot;
        *n_out = julia_range(start, 1., o->t1 - 1, t_save, 1 << 30);
        return 0;
    }
    const double tot = (o->t1 - o->t0) * (1 - rel_transient_time);
    const double start = o->t1 - tot;
    const double dt = tot / n_t;
    *n_out = julia_range(start, dt, o->t1 - dt, t_save, 1 << 30);
    return 0;
}

/* One trajectory, returning the saved samples (n_dim x n_t, column-major) — validation hook. */
int orc_trajectory(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const double* u0,
                   const double* par_row, const double* t_save, int32_t n_t, double* saved_out,
                   int64_t* nsteps_out, int32_t* n_saved_out) {
    sysctx S;
    S.d = sys;
    memcpy(S.sc, sys->scalars, sizeof(S.sc));
    for (int p = 0; p < sys->n_par; p++) S.sc[sys->par_slot[p]] = par_row[p];
    S.tmp = (double*)malloc(sizeof(double) * (sys->n_edges + 1));
    const sopts o = resolve_opts(opts);
    const int n = sys->n_dim;
    double* work = (double*)malloc(sizeof(double) * 12 * n);
    int ns = 0;
    const int ret = integrate(&S, &o, u0, n, t_save, n_t, saved_out, &nsteps_out[0], &nsteps_out[1], &ns, work);
    *n_saved_out = ns;
    free(work);
    free(S.tmp);
    return ret;
}

/* solve + eval_ode_run over the whole ensemble: setup_mc_prob.jl:1098-1116 with output_func =
 * eval_ode_run (eval_mc_prob.jl:75-198).  Same argument meaning as mcbb_solve_features. */
int orc_solve_features(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                       int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                       double* feat_out, double* matrix_out, double* global_out, int32_t* retcode_out,
                       int64_t* nsteps_out, int32_t n_threads, int32_t kl_force_len) {
    const sopts o = resolve_opts(opts);
    const int n = sys->n_dim;
    const int cplx = sys->system == MCBB_SYS_STUART_LANDAU;
    const int n_state = cplx ? n / 2 : n;
    const int nf = feat->n_filter > 0 ? feat->n_filter : n_state;
    (void)nf;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel
    {
        double* work = (double*)malloc(sizeof(double) * 12 * n);
        double* saved = (double*)malloc(sizeof(double) * (size_t)n * n_t);
        double* u0 = (double*)malloc(sizeof(double) * n);
        double* sol_i = (double*)malloc(sizeof(double) * n_t);
        sysctx S;
        S.d = sys;
        S.tmp = (double*)malloc(sizeof(double) * (sys->n_edges + 1));
#pragma omp for schedule(dynamic, 8)
        for (int64_t r = 0; r < n_mc; r++) {
            memcpy(S.sc, sys->scalars, sizeof(S.sc));
            for (int p = 0; p < sys->n_par; p++) S.sc[sys->par_slot[p]] = par[r + n_mc * p];
            for (int d = 0; d < n; d++) u0[d] = ic[r + n_mc * d];
            int64_t ns = 0, nr = 0;
            int nsaved = 0;
            const int ret = integrate(&S, &o, u0, n, t_save, n_t, saved, &ns, &nr, &nsaved, work);
            if (retcode_out) retcode_out[r] = ret;
            if (nsteps_out) {
                nsteps_out[r] = ns;
                nsteps_out[r + n_mc] = nr;
            }
            if (nsaved < n_t) /* truncated solution: mark the unsaved tail NaN */
                for (size_t q = (size_t)nsaved * n; q < (size_t)n_t * n; q++) saved[q] = NAN;
            featurize(sys, feat, saved, n, n_t, feat_out, matrix_out, global_out, (size_t)n_mc, (size_t)r,
                      sol_i, kl_force_len);
        }
        free(S.tmp);
        free(sol_i);
        free(u0);
        free(saved);
        free(work);
    }
    return 0;
}

/* eval_ode_run on caller-provided saved samples (n_mc arrays of n_dim x n_t) — featurizer-only check. */
int orc_featurize_saved(const mcbb_system_desc* sys, const mcbb_feat_opts* feat, int64_t n_mc,
                        const double* saved_all, int32_t n_t, double* feat_out, double* matrix_out,
                        double* global_out, int32_t kl_force_len) {
    const int n = sys->n_dim;
    double* sol_i = (double*)malloc(sizeof(double) * n_t);
    for (int64_t r = 0; r < n_mc; r++)
        featurize(sys, feat, saved_all + (size_t)r * n * n_t, n, n_t, feat_out, matrix_out, global_out,
                  (size_t)n_mc, (size_t)r, sol_i, kl_force_len);
    free(sol_i);
    return 0;
}

/* distance_matrix with the default L1 distance_func, eval_clustering.jl:164-254.
 * mat_elements[i,j] += weight * sum(abs.(x_i .- x_j)) per measure for i<j, then D += D'. */
int orc_distance_matrix(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                        const double* group_weight, double* D, int32_t n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; i++) {
        D[i + n * i] = 0.;
        for (int64_t j = i + 1; j < n; j++) {
            double acc = 0.;
            int64_t f = 0;
            for (int g = 0; g < n_groups; g++) {
                double s = 0.;
                for (int q = 0; q < group_len[g]; q++, f++) s += fabs(X[i + n * f] - X[j + n * f]);
                acc += group_weight[g] * s;
            }
            D[i + n * j] = acc;
            D[j + n * i] = acc;
        }
    }
    return 0;
}

/* Clustering.dbscan(D, eps, minpts) — sequential transcription of src/sparse_clustering.jl:138-209
 * (visit order 1:n, strict `<`, FIFO expansion queue).  Labels 0 = noise; seeds 1-based. */
int orc_dbscan(const double* D, int64_t n, double eps, int32_t minpts, int64_t* assignments, int64_t* seeds,
               int64_t* counts, int64_t* n_clusters) {
    if (n < 2 || !(eps > 0) || minpts < 1) return 1;
    uint8_t* visited = (uint8_t*)calloc(n, 1);
    int64_t* queue = NULL;
    size_t qcap = 0;
    for (int64_t i = 0; i < n; i++) assignments[i] = 0;
    int64_t k = 0;
    int64_t* nbs = (int64_t*)malloc(sizeof(int64_t) * n);
    for (int64_t p = 0; p < n; p++) {
        if (assignments[p] != 0 || visited[p]) continue;
        visited[p] = 1;
        int64_t nn = 0;
        for (int64_t i = 0; i < n; i++)
            if (D[i + n * p] < eps) nbs[nn++] = i; /* _dbs_region_query, :168-178 */
        if (nn < minpts) continue;
        k++;
        /* _dbs_expand_cluster!, :180-209 */
        size_t qh = 0, qt = 0;
        if (qcap < (size_t)nn) {
            qcap = (size_t)nn * 2 + 16;
            queue = (int64_t*)realloc(queue, sizeof(int64_t) * qcap);
        }
        for (int64_t i = 0; i < nn; i++) queue[qt++] = nbs[i];
        assignments[p] = k;
        int64_t cnt = 1;
        while (qh < qt) {
            const int64_t q = queue[qh++];
            if (!visited[q]) {
                visited[q] = 1;
                int64_t qn = 0;
                for (int64_t i = 0; i < n; i++)
                    if (D[i + n * q] < eps) nbs[qn++] = i;
                if (qn >= minpts) {
                    for (int64_t i = 0; i < qn; i++) {
                        if (assignments[nbs[i]] == 0) {
                            if (qt == qcap) {
                                if (qh > 0) { /* compact the consumed head first */
                                    memmove(queue, queue + qh, sizeof(int64_t) * (qt - qh));
                                    qt -= qh;
                                    qh = 0;
                                }
                                if (qt * 2 > qcap) {
                                    qcap = qcap * 2 + 16;
                                    queue = (int64_t*)realloc(queue, sizeof(int64_t) * qcap);
                                }
                            }
                            queue[qt++] = nbs[i];
                        }
                    }
                }
            }
            if (assignments[q] == 0) {
                assignments[q] = k;
                cnt++;
            }
        }
        seeds[k - 1] = p + 1;
        counts[k - 1] = cnt;
    }
    *n_clusters = k;
    free(nbs);
    free(queue);
    free(visited);
    return 0;
}

/* cluster_membership sliding-window version, eval_clustering.jl:1206-1265, one parameter.
 * windows_mins = min_par:offset:(max_par-window) (:1434-1457); member iff ip < par < ip+window (:1222).
 * out is n_windows x (k+1), column-major; returns n_windows via *n_windows (call with out=NULL first). */
int orc_cluster_membership(const double* par, const int64_t* assignments, int64_t n, int64_t k, double window,
                           double offset, int32_t normalize, double* out, double* par_mesh,
                           int32_t* n_windows) {
    double mn = par[0], mx = par[0];
    for (int64_t i = 1; i < n; i++) {
        if (par[i] < mn) mn = par[i];
        if (par[i] > mx) mx = par[i];
    }
    const int nw = julia_range(mn, offset, mx - window, NULL, 0);
    *n_windows = nw;
    if (!out) return 0;
    double* mins = (double*)malloc(sizeof(double) * (nw + 1));
    julia_range(mn, offset, mx - window, mins, nw);
    for (int w = 0; w < nw; w++) {
        const double ip = mins[w];
        if (par_mesh) par_mesh[w] = ip;
        for (int64_t c = 0; c <= k; c++) out[w + (size_t)nw * c] = 0.;
        int64_t cnt = 0;
        for (int64_t i = 0; i < n; i++)
            if (par[i] > ip && par[i] < ip + window) {
                out[w + (size_t)nw * assignments[i]] += 1.;
                cnt++;
            }
        if (normalize)
            for (int64_t c = 0; c <= k; c++) out[w + (size_t)nw * c] = out[w + (size_t)nw * c] / (double)cnt;
    }
    free(mins);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* histogram-mode distance: distance_matrix(...; histograms=true), eval_clustering.jl:187-211, 261-268,     */
/* 587-645, 746-762                                                                                         */
/* ------------------------------------------------------------------------------------------------------ */
static int cmp_double(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}
/* StatsBase.iqr: type-7 quantiles (the Julia / numpy default) */
static double quantile7(const double* sorted, int64_t n, double p) {
    const double h = (double)(n - 1) * p;
    const int64_t lo = (int64_t)floor(h);
    const int64_t hi = lo + 1 < n ? lo + 1 : lo;
    return sorted[lo] + (h - (double)lo) * (sorted[hi] - sorted[lo]);
}
/* _compute_hist_edges, eval_clustering.jl:607-642 (automatic binning branch).  Returns the number of edges. */
int orc_hist_edges(const double* data, int64_t n_total, int32_t n_dim, double k_bin, int32_t nbin_default,
                   double* edges_out /* >= nbin_default + 4 */, double* bin_width_out) {
    double* srt = (double*)malloc(sizeof(double) * n_total);
    memcpy(srt, data, sizeof(double) * n_total);
    qsort(srt, n_total, sizeof(double), cmp_double);
    const double minval = srt[0], maxval = srt[n_total - 1];
    const double iqr = quantile7(srt, n_total, 0.75) - quantile7(srt, n_total, 0.25);
    free(srt);
    double bw = 0.5 * k_bin * iqr / pow((double)n_dim, 1. / 3.); /* freedman_draconis_bin_width, :644 */
    int len = 0;
    if (bw != 0) {
        len = julia_range(minval - bw, bw, maxval + bw, NULL, 0);
        if (len <= nbin_default) {
            julia_range(minval - bw, bw, maxval + bw, edges_out, len);
            *bin_width_out = bw;
            return len;
        }
    }
    /* fallback (:620-622, 626-628): range(minval - 0.1*minval, maxval + 0.1*maxval, length=nbin_default) */
    const long double a = (long double)(minval - 0.1 * minval), b = (long double)(maxval + 0.1 * maxval);
    len = nbin_default;
    for (int i = 0; i < len; i++) edges_out[i] = (double)(a + (b - a) * (long double)i / (long double)(len - 1));
    *bin_width_out = edges_out[1] - edges_out[0];
    return len;
}
/* _compute_hist_weights for one run (eval_clustering.jl:587-598): Float32 histogram over the common edges
 * (closed = :left, out-of-range values dropped), then ecdf_hist in Float32 (:759-762). */
void orc_hist_ecdf(const double* x, int32_t n, const double* edges, int32_t n_edges, int32_t use_ecdf, float* out) {
    const int nb = n_edges - 1;
    const double first = edges[0], step = edges[1] - edges[0];
    for (int b = 0; b < nb; b++) out[b] = 0.f;
    for (int i = 0; i < n; i++) {
        double f = (x[i] - first) / step + 1.; /* Base.searchsortedlast on a range */
        if (f < 1.) f = 1.;
        if (f > n_edges) f = n_edges;
        int k = (int)nearbyint(f);
        if (x[i] < edges[k - 1]) k -= 1;
        if (x[i] != x[i]) continue;
        if (k >= 1 && k <= nb) out[k - 1] += 1.f;
    }
    if (use_ecdf) {
        float cum = 0.f;
        for (int b = 0; b < nb; b++) {
            cum += out[b];
            out[b] = cum;
        }
        const float last = out[nb - 1];
        for (int b = 0; b < nb; b++) out[b] = out[b] / last;
    }
}
/* distance_matrix(sol, prob, weights; histograms=true, k_bin, nbin_default) for per-dimension measures plus the
 * parameter term.  feat: [n_mc][n_dim][n_meas] column-major as produced by orc_solve_features. */
int orc_distance_matrix_hist(const double* feat, int64_t n_mc, int32_t n_dim, int32_t n_meas, const double* par,
                             int32_t n_par, const double* weights, double k_bin, int32_t nbin_default, double* D) {
    for (int64_t q = 0; q < n_mc * n_mc; q++) D[q] = 0.;
    double* edges = (double*)malloc(sizeof(double) * (nbin_default + 8));
    double* xi = (double*)malloc(sizeof(double) * n_dim);
    for (int m = 0; m < n_meas; m++) {
        const double* data = feat + (size_t)m * n_dim * n_mc; /* n_mc x n_dim block */
        double bw;
        const int ne = orc_hist_edges(data, n_mc * n_dim, n_dim, k_bin, nbin_default, edges, &bw);
        const int nb = ne - 1;
        float* H = (float*)malloc(sizeof(float) * (size_t)n_mc * nb);
        for (int64_t i = 0; i < n_mc; i++) {
            for (int d = 0; d < n_dim; d++) xi[d] = data[i + n_mc * d];
            orc_hist_ecdf(xi, n_dim, edges, ne, 1, H + (size_t)i * nb);
        }
        for (int64_t i = 0; i < n_mc; i++)
            for (int64_t j = i + 1; j < n_mc; j++) { /* wasserstein_histogram_distance: Float32 sum times delta */
                float s = 0.f;
                for (int b = 0; b < nb; b++) s += fabsf(H[(size_t)i * nb + b] - H[(size_t)j * nb + b]);
                D[i + n_mc * j] += weights[m] * ((double)s * bw);
            }
        free(H);
    }
    for (int p = 0; p < n_par; p++)
        for (int64_t i = 0; i < n_mc; i++)
            for (int64_t j = i + 1; j < n_mc; j++)
                D[i + n_mc * j] += weights[n_meas + p] * fabs(par[i + n_mc * p] - par[j + n_mc * p]);
    for (int64_t i = 0; i < n_mc; i++)
        for (int64_t j = i + 1; j < n_mc; j++) D[j + n_mc * i] = D[i + n_mc * j];
    free(xi);
    free(edges);
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
