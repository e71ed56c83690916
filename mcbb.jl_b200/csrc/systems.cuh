// systems.cuh — right-hand sides of the built-in systems (reference: src/systems.jl; Stuart-Landau from
// paper/stuartlandau/stuart_landau_sathiyadevi-standard-config.ipynb cell 1), one trajectory per lane.
//
// State vectors live in shared memory as [dim][lane] (element d of this lane at p[d*bs]) so that a warp's
// accesses to one dimension hit 32 consecutive banks.  Coupling tables are the same for every lane: they
// are staged once per CTA in shared memory (or read through the read-only path when too large) and every
// access is a warp-wide broadcast.
//
// Kuramoto-type couplings use  sum_j A_ij sin(u_j-u_i) = cos(u_i)(A sin u)_i - sin(u_i)(A cos u)_i :
// N sincos + 2N^2 FMA instead of N^2 sin (systems.jl:119-130 is the literal form the oracle keeps).
#pragma once
#include <math.h>

#include "common.cuh"

namespace mcbb {

struct SysTabs {  // table pointers as seen by the kernel (shared-memory copies when staged)
    const double *mat_a, *mat_b, *vec_a, *vec_b, *vec_c;
};

template <int SYS>
struct SysTraits {
    static constexpr int n_scratch =
        (SYS == MCBB_SYS_KURAMOTO_NETWORK || SYS == MCBB_SYS_KURAMOTO || SYS == MCBB_SYS_NONLOCAL_KURAMOTO_RING)
            ? 2
            : 0;
    static constexpr bool is_map = (SYS == MCBB_SYS_LOGISTIC);
};

// u, du, scr: lane base pointers; element d at [d*bs]; scr has n_scratch vectors of n_dim
template <int SYS>
MCBB_HD void rhs_lane(const SysDev& S, const SysTabs& T, const double* sc, const double* u, double* du,
                      double* scr, const int bs) {
    const int N = S.n_osc;
    if constexpr (SYS == MCBB_SYS_LOGISTIC) {  // systems.jl:34-36
        const double x = u[0];
        du[0] = sc[0] * x * (1. - x);
    } else if constexpr (SYS == MCBB_SYS_KURAMOTO) {  // systems.jl:85-94
        double* ss = scr;
        double* cs = scr + (size_t)N * bs;
        double sum_s = 0., sum_c = 0.;
        for (int j = 0; j < N; j++) {
            double s, c;
            sincos(u[j * bs], &s, &c);
            ss[j * bs] = s;
            cs[j * bs] = c;
            sum_s += s;
            sum_c += c;
        }
        const double KN = sc[0] / (double)N;
        for (int i = 0; i < N; i++) du[i * bs] = fma(KN, cs[i * bs] * sum_s - ss[i * bs] * sum_c, T.vec_a[i]);
    } else if constexpr (SYS == MCBB_SYS_KURAMOTO_NETWORK) {  // systems.jl:119-130
        double* ss = scr;
        double* cs = scr + (size_t)N * bs;
        for (int j = 0; j < N; j++) {
            double s, c;
            sincos(u[j * bs], &s, &c);
            ss[j * bs] = s;
            cs[j * bs] = c;
        }
        const double KN = sc[0] / (double)N;
        for (int i = 0; i < N; i++) {
            double as = 0., ac = 0.;
            for (int j = 0; j < N; j++) {
                const double a = T.mat_a[i + N * j];
                if (a != 0.) {  // uniform across the warp: same table for every lane
                    as = fma(a, ss[j * bs], as);
                    ac = fma(a, cs[j * bs], ac);
                }
            }
            du[i * bs] = fma(KN, cs[i * bs] * as - ss[i * bs] * ac, T.vec_a[i]);
        }
    } else if constexpr (SYS == MCBB_SYS_SECOND_ORDER_KURAMOTO) {  // systems.jl:198-202
        // mat_b holds the incidence matrix by edge in compressed form (built at upload, api.cu):
        //   [0..E]          start of each edge's entries          (as doubles)
        //   [E+1 .. E+nnz]  node index of each entry, ascending   (same order as the dense row scan)
        //   [E+1+nnz ..]    incidence value of each entry
        const int E = S.n_edges;
        const double* ptr = T.mat_b;
        const int nnz = (int)ptr[E];
        const double* node = T.mat_b + E + 1;
        const double* val = node + nnz;
        for (int i = 0; i < N; i++) du[(N + i) * bs] = 0.;
        for (int e = 0; e < E; e++) {
            const int p0 = (int)ptr[e], p1 = (int)ptr[e + 1];
            double arg = 0.;
            for (int q = p0; q < p1; q++) arg += val[q] * u[(int)node[q] * bs];  // incidence' * theta
            const double sv = sin(arg);
            for (int q = p0; q < p1; q++) du[(N + (int)node[q]) * bs] += val[q] * sv;  // incidence * sin(.)
        }
        for (int i = 0; i < N; i++) {
            const double om = u[(N + i) * bs];
            du[i * bs] = om;
            du[(N + i) * bs] = (T.vec_a[i] - sc[0] * om) - sc[1] * du[(N + i) * bs];
        }
    } else if constexpr (SYS == MCBB_SYS_NONLOCAL_KURAMOTO_RING) {  // systems.jl:230-239
        double* ss = scr;
        double* cs = scr + (size_t)N * bs;
        for (int j = 0; j < N; j++) {
            double s, c;
            sincos(u[j * bs], &s, &c);
            ss[j * bs] = s;
            cs[j * bs] = c;
        }
        double sa, ca;
        sincos(sc[1], &sa, &ca);
        const double omega0 = sc[0], fak = sc[2];
        for (int k = 0; k < N; k++) {
            double gs = 0., gc = 0.;
            for (int j = 0; j < N; j++) {
                const double g = T.mat_a[(k - j) + N - 1];
                gs = fma(g, ss[j * bs], gs);
                gc = fma(g, cs[j * bs], gc);
            }
            const double sk = ss[k * bs] * ca + cs[k * bs] * sa;  // sin(u_k + alpha)
            const double ck = cs[k * bs] * ca - ss[k * bs] * sa;  // cos(u_k + alpha)
            du[k * bs] = -(sk * gc - ck * gs) / fak + omega0;
        }
    } else if constexpr (SYS == MCBB_SYS_STUART_LANDAU) {  // SL notebook cell 1
        // mat_a / mat_b hold A_1 / A_2 as compressed rows (built at upload, api.cu): [0..N] row starts,
        // then the column indices in ascending order (= the summation order of the reference's dense scan)
        const double eps = sc[0], lambda = sc[1], omega = sc[2];
        const double epsP1 = eps / (2. * sc[3]);
        const double epsP2 = eps / (2. * sc[4]);
        const double* p1 = T.mat_a;
        const double* c1 = T.mat_a + N + 1;
        const double* p2 = T.mat_b;
        const double* c2 = T.mat_b + N + 1;
        for (int i = 0; i < N; i++) {
            const double x = u[(2 * i) * bs], y = u[(2 * i + 1) * bs];
            double attr = 0., rep = 0.;
            for (int q = (int)p1[i]; q < (int)p1[i + 1]; q++) attr += u[(2 * (int)c1[q]) * bs] - x;
            for (int q = (int)p2[i]; q < (int)p2[i + 1]; q++) rep += u[(2 * (int)c2[q] + 1) * bs] - y;
            attr *= epsP1;
            rep *= epsP2;
            const double cr = lambda - (x * x + y * y);
            du[(2 * i) * bs] = (cr * x - omega * y) + attr;
            du[(2 * i + 1) * bs] = (cr * y + omega * x) - rep;
        }
    } else if constexpr (SYS == MCBB_SYS_ROESSLER_NETWORK) {  // systems.jl:286-298 incl. the du-coupling quirk
        const double K = sc[0];
        for (int i = 0; i < N; i++) {
            const int ii = 3 * i;
            double dx = -u[(ii + 1) * bs] - u[(ii + 2) * bs];
            for (int j = 0; j < N; j++) {
                const double l = T.mat_a[i + N * j];
                if (l != 0.) dx -= K * l * dx;
            }
            du[ii * bs] = dx;
            du[(ii + 1) * bs] = u[ii * bs] + T.vec_a[i] * u[(ii + 1) * bs];
            du[(ii + 2) * bs] = T.vec_b[i] + u[(ii + 2) * bs] * (dx - T.vec_c[i]);
        }
    }
}

}  // namespace mcbb
