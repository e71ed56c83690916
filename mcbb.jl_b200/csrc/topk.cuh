// topk.cuh — fused per-row K-smallest distances for the eps-selection helpers k_dist / KNN_dist / KNN_dist_relative
// (src/eval_clustering.jl:1504-1553) without materialising or sorting rows of D.
//
// A CTA owns 64 rows of the pair space and sweeps ALL column tiles (every row needs its whole row of D; the zero
// self-distance counts, as in the reference).  The KK = max(k, K) <= TOPK_MAX smallest distances of every row live in
// shared memory as a sorted list; its largest element is the row's threshold.  The distances of a 64 x 64 tile are
// accumulated exactly like k_pair_tiles does (same order of additions, same fma at the group ends: the values are
// bit-identical to the materialised matrix), with the same monotone early exit — against the ROW's threshold instead of
// eps: once the lists are warm almost every tile is abandoned after its first feature slab.  The sweep starts at the
// diagonal tile and moves outward, because runs are sorted by parameter and index neighbours are the likely near
// neighbours.  Candidates below the threshold are appended to a per-row buffer (at most 64 per tile) and merged by one
// thread per row between tiles.
#pragma once
#include "pair_tiles.cuh"

namespace mcbb {

constexpr int TOPK_MAX = 64;

struct TopkArgs {
    const double* X;  // [F][n]
    long long n;
    long long r0, r1;  // rows handled by this launch
    int F;
    const double* fw;
    const unsigned short* fmask;
    int early_exit;
    int KK;            // list length, max(k, K)
    int k, K;
    double* kth;       // [r1 - r0] or null: sort(D[i,:])[k]
    double* cum;       // [r1 - r0] or null: sum(sort(D[i,:])[1:K])
};

__global__ void __launch_bounds__(256) k_topk_rows(const TopkArgs a) {
    extern __shared__ __align__(16) unsigned char tk_smem[];
    double (*s_slab)[2][FC][TILE] = reinterpret_cast<double (*)[2][FC][TILE]>(tk_smem);            // [2 buffers][side][f][point]
    double* s_list = reinterpret_cast<double*>(tk_smem + sizeof(double) * 2 * 2 * FC * TILE);       // [64][TOPK_MAX]
    double* s_cand = s_list + TILE * TOPK_MAX;                                                      // [64][64]
    double* s_thr = s_cand + TILE * TILE;                                                           // [64]
    int* s_ncand = reinterpret_cast<int*>(s_thr + TILE);                                            // [64]
    int* s_nlist = s_ncand + TILE;                                                                  // [64]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long tr = a.r0 + (long long)blockIdx.x * TILE;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if (tid < TILE) {
        s_thr[tid] = INF;
        s_ncand[tid] = 0;
        s_nlist[tid] = 0;
    }
    // entries a row never fills (fewer than KK finite distances) read as NaN, where a sort would have put them
    for (int q = tid; q < TILE * TOPK_MAX; q += 256) s_list[q] = nan("");
    __syncthreads();
    const long long nct = (a.n + TILE - 1) / TILE;  // column tiles
    const long long t0 = tr / TILE;                 // the diagonal tile of these rows
    for (long long step = 0; step < 2 * nct; step++) {
        // outward sweep: t0, t0+1, t0-1, t0+2, ...
        const long long off = (step + 1) >> 1;
        const long long t = (step & 1) ? t0 + off : t0 - off;
        if (step == 0 ? false : (t < 0 || t >= nct)) continue;
        if (step > 0 && off == 0) continue;
        const long long tc = t * TILE;
        double thr[TT];
#pragma unroll
        for (int i = 0; i < TT; i++) thr[i] = s_thr[tx * TT + i];
        double acc[TT][TT], accg[TT][TT];
#pragma unroll
        for (int i = 0; i < TT; i++)
#pragma unroll
            for (int j = 0; j < TT; j++) acc[i][j] = accg[i][j] = 0.;
        auto issue_slab = [&](const int f0, const int buf) {
            const int fc = min(FC, a.F - f0);
            for (int q = tid; q < FC * TILE; q += 256) {
                const int f = q / TILE, kk = q - f * TILE;
                const long long ir = tr + kk, ic = tc + kk;
                const bool fr = (f < fc) && (ir < a.r1), fcn = (f < fc) && (ic < a.n);
                const double* gr = a.X + (long long)(f0 + (f < fc ? f : 0)) * a.n + (fr ? ir : a.r0);
                const double* gc = a.X + (long long)(f0 + (f < fc ? f : 0)) * a.n + (fcn ? ic : 0);
                const unsigned dr = (unsigned)__cvta_generic_to_shared(&s_slab[buf][0][f][kk]);
                const unsigned dc = (unsigned)__cvta_generic_to_shared(&s_slab[buf][1][f][kk]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dr), "l"(gr), "r"(fr ? 8 : 0));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dc), "l"(gc), "r"(fcn ? 8 : 0));
            }
            asm volatile("cp.async.commit_group;");
        };
        issue_slab(0, 0);
        int buf = 0;
        bool warp_alive = true, tile_dead = false;
        for (int f0 = 0; f0 < a.F; f0 += FC, buf ^= 1) {
            const int fc = min(FC, a.F - f0);
            asm volatile("cp.async.wait_group 0;");
            __syncthreads();
            if (f0 + FC < a.F) issue_slab(f0 + FC, buf ^ 1);
            const unsigned emask = a.fmask[f0 / FC];
            const double (*s_r)[TILE] = s_slab[buf][0];
            const double (*s_c)[TILE] = s_slab[buf][1];
#pragma unroll 4
            for (int f = 0; f < (warp_alive ? fc : 0); f++) {
                double xr[TT], xc[TT];
#pragma unroll
                for (int i = 0; i < TT; i++) xr[i] = s_r[f][tx * TT + i];
#pragma unroll
                for (int j = 0; j < TT; j++) xc[j] = s_c[f][ty * TT + j];
#pragma unroll
                for (int i = 0; i < TT; i++)
#pragma unroll
                    for (int j = 0; j < TT; j++) accg[i][j] += fabs(xr[i] - xc[j]);
                if ((emask >> f) & 1u) {
                    const double w = __ldg(a.fw + f0 + f);
#pragma unroll
                    for (int i = 0; i < TT; i++)
#pragma unroll
                        for (int j = 0; j < TT; j++) {
                            acc[i][j] = fma(w, accg[i][j], acc[i][j]);
                            accg[i][j] = 0.;
                        }
                }
            }
            if (a.early_exit && f0 + FC < a.F) {
                bool alive = false;
                if (warp_alive) {
                    const double w = __ldg(a.fw + f0 + fc - 1);
#pragma unroll
                    for (int i = 0; i < TT; i++)
#pragma unroll
                        for (int j = 0; j < TT; j++) alive |= (fma(w, accg[i][j], acc[i][j]) < thr[i]);
                    warp_alive = __any_sync(0xffffffffu, alive);
                    if (!warp_alive) {  // freeze the bound (>= the row's threshold): it can never enter the list
#pragma unroll
                        for (int i = 0; i < TT; i++)
#pragma unroll
                            for (int j = 0; j < TT; j++) acc[i][j] = fma(w, accg[i][j], acc[i][j]);
                    }
                }
                if (!__syncthreads_or(alive)) {
                    asm volatile("cp.async.wait_group 0;");
                    tile_dead = true;
                    break;
                }
            }
        }
        if (!tile_dead) {
            // candidates: strictly below the row's threshold (equal values would not change the list's VALUES)
#pragma unroll
            for (int i = 0; i < TT; i++) {
                const long long ri = tr + tx * TT + i;
                if (ri >= a.r1) continue;
#pragma unroll
                for (int j = 0; j < TT; j++) {
                    const long long cj = tc + ty * TT + j;
                    if (cj >= a.n) continue;
                    const double d = (ri == cj) ? 0. : acc[i][j];  // D[p,p] = 0 also for Inf / NaN features
                    if (d < thr[i]) {
                        const int slot = atomicAdd(&s_ncand[tx * TT + i], 1);
                        s_cand[(tx * TT + i) * TILE + slot] = d;
                    }
                }
            }
        }
        __syncthreads();
        if (!tile_dead && tid < TILE) {  // one thread per row merges its candidates into the sorted list
            const int nc = s_ncand[tid];
            if (nc > 0) {
                double* lst = s_list + tid * TOPK_MAX;
                int nl = s_nlist[tid];
                for (int q = 0; q < nc; q++) {
                    const double d = s_cand[tid * TILE + q];
                    if (nl == a.KK && !(d < lst[nl - 1])) continue;
                    int p = (nl < a.KK) ? nl : nl - 1;  // position that is overwritten / appended
                    while (p > 0 && lst[p - 1] > d) {
                        lst[p] = lst[p - 1];
                        p--;
                    }
                    lst[p] = d;
                    if (nl < a.KK) nl++;
                }
                s_nlist[tid] = nl;
                s_ncand[tid] = 0;
                if (nl == a.KK) s_thr[tid] = lst[nl - 1];
            }
        }
        __syncthreads();
    }
    if (tid < TILE) {
        const long long ri = tr + tid;
        if (ri < a.r1) {
            const double* lst = s_list + tid * TOPK_MAX;
            if (a.kth) a.kth[ri - a.r0] = lst[a.k - 1];
            if (a.cum) {
                double s = 0.;
                for (int q = 0; q < a.K; q++) s += lst[q];
                a.cum[ri - a.r0] = s;
            }
        }
    }
}

constexpr size_t TOPK_SMEM = sizeof(double) * (2 * 2 * FC * TILE + TILE * TOPK_MAX + TILE * TILE + TILE) + sizeof(int) * 2 * TILE;

}  // namespace mcbb
