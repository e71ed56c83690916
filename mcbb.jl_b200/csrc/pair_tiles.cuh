// pair_tiles.cuh — tiled pairwise L1 distance over featurized runs; one kernel body, four consumers.
//
// Reference: distance_matrix / _compute_distance! (src/eval_clustering.jl:164-254) and the neighbourhood
// queries of dbscan (src/sparse_clustering.jl:168-178).
//
//   D[i,j] = sum_g w_g * sum_{f in g} |X[f,i] - X[f,j]|          X is [F][n], run index contiguous
//
// A CTA of 256 threads owns a 64 x 64 tile of the pair space; feature slabs of FC rows are staged in
// shared memory (coalesced 512 B rows), every thread keeps a 4 x 4 register block.  The same device
// function produces the distance for all consumers, so the materialised matrix and the on-the-fly DBSCAN
// see bit-identical values (required for label equality, SURVEY §7 hard part 6).
//
// Consumers:  TM_MATERIALIZE  write D (symmetric pair of stores per tile)
//             TM_COUNT        eps-neighbour counts (DBSCAN pass 1)
//             TM_UNION        union-find over core-core edges (pass 2)
//             TM_BORDER       smallest adjacent core root per non-core point (pass 3)
//             TM_SPARSE       compressed-column fill of the entries below a threshold, stored as d - threshold
//                             (NonzeroSparseMatrix, src/sparse_clustering.jl:37-110)
// All three DBSCAN passes stop accumulating a tile as soon as every pair in it is provably >= eps
// (terms are non-negative, floating-point addition of non-negatives is monotone, so the test is exact).
#pragma once
#include "common.cuh"

namespace mcbb {

enum TileMode { TM_MATERIALIZE = 0, TM_COUNT = 1, TM_UNION = 2, TM_BORDER = 3, TM_SPARSE = 4 };

constexpr int TILE = 64;   // tile edge
constexpr int TT = 4;      // register block edge
constexpr int FC = 16;     // features per shared-memory slab

struct TileArgs {
    const double* Xr;  // row-side feature matrix [F][ldr]
    long long ldr;
    long long r0, r1;  // row range handled by this launch
    const double* Xc;  // column-side feature matrix [F][ldc]
    long long ldc;
    long long c0, c1;
    int F;
    const double* fw;          // [F] weight of the feature's group
    const unsigned char* fend; // [F] 1 when the feature closes its group
    const unsigned short* fmask;  // [ceil(F/16)] bit f%16 set when feature f closes its group (one load per slab)
    double eps;
    int symmetric;   // Xr == Xc: only pairs col > row are visited, results applied to both sides
    int early_exit;  // all weights >= 0
    double* D;       // TM_MATERIALIZE: [n][n] column-major, ld = ldr
    int* cnt;        // TM_COUNT: indexed by row id (and column id when symmetric)
    int* parent;     // TM_UNION
    const int* col_root;  // TM_BORDER: root (original index) of each core column
    int* best;            // TM_BORDER: per non-core row, min root seen (init INT_MAX)
    // TM_COUNT, optional: the first nbr_cap neighbours found for every row are recorded (row-local index = rid[ri] or
    // ri - nbr_row0).  A point that ends below minpts has ALL its neighbours in the list, which is what the border
    // assignment needs: the third pass over the pair space disappears.
    int* nbr;
    int* nfill;
    int nbr_cap;
    const int* rid;
    long long nbr_row0;
    const long long* sp_colptr;  // TM_SPARSE: [n+1] start of every column (0-based)
    int* sp_fill;                // TM_SPARSE: [n] entries written so far per column
    int* sp_row;                 // TM_SPARSE: row index of every stored entry (unordered within a column)
    double* sp_val;              // TM_SPARSE: d - threshold
    // Saturation skips (exact: DBSCAN only needs core-ness, connectivity and the smallest adjacent seed, all of
    // which are monotone).  A tile is skipped before any feature is loaded when nothing in it can still change:
    int sat_count;        // TM_COUNT: every row (and column, if symmetric) already has >= sat_count neighbours
    int sat_seed;         // TM_BORDER: every row already holds the globally smallest seed
    int sat_union;        // TM_UNION: every row and column already shares one root
    int col_mod, col_lo, col_hi;  // col_mod > 1: only column tiles with (tile index % col_mod) in [col_lo, col_hi) are
                                  // visited: evenly spread, disjoint samples of the columns (progressive counting)
    int fold;             // symmetric full launch: the grid is nt x (nt/2 + 1); block (x, y) owns tile (y, x) if x >= y,
                          // else tile (nt - y, nt - 1 - x): every block lands on or above the diagonal, none is launched
                          // only to return
    unsigned long long* pf_ctr;  // diagnostic (mcbb_pair_feature_counter): executed pair x feature terms, or null
    int self_zero;        // non-symmetric launch whose rows and columns index the same points: the pair (p, p) is the
                          // diagonal D[p,p] = 0 of the reference (eval_clustering.jl:242-247 never computes it), also
                          // when a feature of p is Inf / NaN (|Inf - Inf| = NaN would drop the point's self count)
    int band;             // > 0: only column tiles within `band` tiles of the row tile are visited; blockIdx.x is
                          // the offset (symmetric: 0..band, else -band..band).  Runs are sorted by parameter, so
                          // index neighbours are the likely distance neighbours: a cheap first pass that saturates
                          // most points before the exact passes.
};

__device__ __forceinline__ int uf_find(int* parent, int x) {
    int p = parent[x];
    while (p != x) {  // path halving; hooks always point to a smaller index, so races are benign
        const int gp = parent[p];
        if (gp != p) parent[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            const int tmp = a;
            a = b;
            b = tmp;
        }
        // a > b: hook the larger root under the smaller one => the final root is the smallest member,
        // i.e. the seed of the sequential algorithm (sparse_clustering.jl:149-157)
        const int old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_pair_tiles(const TileArgs a) {
    __shared__ __align__(16) double s_slab[2][2][FC][TILE];  // [buffer][row side / column side][feature][point]; a
                                                             // buffer doubles as the 32 x 64 staging tile of the
                                                             // materialising store

    long long tr = a.r0 + (long long)blockIdx.y * TILE;
    long long tc = a.c0 + (long long)blockIdx.x * TILE;
    if (a.fold && blockIdx.x < blockIdx.y) {
        const long long nt = gridDim.x;
        if (2 * (long long)blockIdx.y == nt) return;  // the middle row of an even grid is already complete
        tr = a.r0 + (nt - blockIdx.y) * TILE;
        tc = a.c0 + (nt - 1 - blockIdx.x) * TILE;
    }
    if (a.col_mod > 1) {  // blockIdx.x enumerates the sampled tiles: period index * width + offset inside the residue range
        const int wdt = a.col_hi - a.col_lo;
        const long long per = blockIdx.x / wdt;
        const int off = blockIdx.x - (int)per * wdt;
        tc = a.c0 + (per * a.col_mod + a.col_lo + off) * TILE;
        if (tc >= a.c1) return;
    }
    if (a.band > 0) {  // band pass: column tile relative to the row tile (tile grids of rows and columns coincide)
        tc = tr + ((long long)blockIdx.x - (a.symmetric ? 0 : a.band)) * TILE;
        if (tc < a.c0 || tc >= a.c1) return;
    }
    if (a.symmetric && tc + TILE <= tr) return;  // strictly below the diagonal
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // tx -> rows (contiguous in memory), ty -> columns

    if (MODE == TM_COUNT && a.sat_count > 0) {
        bool sat = true;
        if (tid < TILE) {
            const long long ri = tr + tid;
            if (ri < a.r1) sat = __ldcg(a.cnt + ri) >= a.sat_count;
        } else if (tid < 2 * TILE && a.symmetric) {
            const long long cj = tc + (tid - TILE);
            if (cj < a.c1) sat = __ldcg(a.cnt + cj) >= a.sat_count;
        }
        if (__syncthreads_and(sat)) return;
    }
    if (MODE == TM_UNION && a.sat_union) {
        __shared__ int s_root0;
        int root = -1;
        if (tid < 2 * TILE) {
            const long long id = (tid < TILE) ? tr + tid : tc + (tid - TILE);
            const long long lim = (tid < TILE) ? a.r1 : a.c1;
            if (id < lim) root = uf_find(a.parent, (int)id);
        }
        if (tid == 0) s_root0 = root;  // row tr always exists
        __syncthreads();
        const bool same = (root < 0) || (root == s_root0);
        if (__syncthreads_and(same)) return;
    }
    if (MODE == TM_BORDER && a.sat_seed >= 0) {
        bool sat = true;
        if (tid < TILE) {
            const long long ri = tr + tid;
            if (ri < a.r1) sat = __ldcg(a.best + ri) == a.sat_seed;
        }
        if (__syncthreads_and(sat)) return;
    }

    double acc[TT][TT], accg[TT][TT];
#pragma unroll
    for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TT; j++) acc[i][j] = accg[i][j] = 0.;

    // Feature slabs are double buffered with cp.async (LDGSTS): slab k+1 streams into the other buffer while slab k
    // is being accumulated; out-of-range rows / columns / features are zero-filled (src-size 0).
    auto issue_slab = [&](const int f0, const int buf) {
        const int fc = min(FC, a.F - f0);
        for (int q = tid; q < FC * TILE; q += 256) {
            const int f = q / TILE, k = q - f * TILE;
            const long long ir = tr + k, ic = tc + k;
            const bool fr = (f < fc) && (ir < a.r1), fcn = (f < fc) && (ic < a.c1);
            const double* gr = a.Xr + (long long)(f0 + (f < fc ? f : 0)) * a.ldr + (fr ? ir : a.r0);
            const double* gc = a.Xc + (long long)(f0 + (f < fc ? f : 0)) * a.ldc + (fcn ? ic : a.c0);
            const unsigned dr = (unsigned)__cvta_generic_to_shared(&s_slab[buf][0][f][k]);
            const unsigned dc = (unsigned)__cvta_generic_to_shared(&s_slab[buf][1][f][k]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dr), "l"(gr), "r"(fr ? 8 : 0));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dc), "l"(gc), "r"(fcn ? 8 : 0));
        }
        asm volatile("cp.async.commit_group;");
    };
    issue_slab(0, 0);
    int buf = 0;
    bool warp_alive = true;
    int nf_done = 0;  // features this warp accumulated (x 512 pairs per warp): the executed work behind cluster.roofline
    auto flush_count = [&]() {
        if (a.pf_ctr && (tid & 31) == 0 && nf_done) atomicAdd(a.pf_ctr, (unsigned long long)nf_done * 512ull);
    };
    for (int f0 = 0; f0 < a.F; f0 += FC, buf ^= 1) {
        const int fc = min(FC, a.F - f0);
        asm volatile("cp.async.wait_group 0;");
        __syncthreads();  // slab `buf` complete for everyone; everyone is also done reading slab `buf^1`
        if (f0 + FC < a.F) issue_slab(f0 + FC, buf ^ 1);
        const unsigned emask = a.fmask[f0 / FC];
        const double (*s_r)[TILE] = s_slab[buf][0];
        const double (*s_c)[TILE] = s_slab[buf][1];
        nf_done += warp_alive ? fc : 0;
        // warp_alive: some pair of this warp's 64 x 8 block can still fall below eps.  A warp whose pairs are all
        // provably >= eps (monotone partial sums) stops accumulating — its values no longer matter to any consumer
        // but the materialising one — while the rest of the tile goes on.
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
            if ((emask >> f) & 1u) {  // group closes: mat_elements += weight * distance_func(...)
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
        if (MODE != TM_MATERIALIZE && a.early_exit && f0 + FC < a.F) {
            bool alive = false;
            if (warp_alive) {
                const double w = __ldg(a.fw + f0 + fc - 1);  // weight of the group still open (if any)
#pragma unroll
                for (int i = 0; i < TT; i++)
#pragma unroll
                    for (int j = 0; j < TT; j++) alive |= (fma(w, accg[i][j], acc[i][j]) < a.eps);
                warp_alive = __any_sync(0xffffffffu, alive);
                if (!warp_alive) {  // freeze the bound (>= eps) as the value the consumers will see
#pragma unroll
                    for (int i = 0; i < TT; i++)
#pragma unroll
                        for (int j = 0; j < TT; j++) acc[i][j] = fma(w, accg[i][j], acc[i][j]);
                }
            }
            if (!__syncthreads_or(alive)) {  // no pair of this tile can still fall below eps
                asm volatile("cp.async.wait_group 0;");  // do not leave the prefetch in flight
                flush_count();
                return;
            }
        }
    }

    flush_count();
    // ---- consume the 4 x 4 block ---------------------------------------------------------------------
    if (MODE == TM_MATERIALIZE) {
        // Stage the tile in shared memory (reusing the feature slabs: 2 x 16 x 64 doubles = one 64 x 32 half
        // at a time) so that BOTH D[tile] and its mirror D[tile]^T leave as 512-byte coalesced rows.
        double* s_t = &s_slab[0][0][0][0];  // 32 x 64 doubles = 16 KB (one buffer: both sides are contiguous)
        static_assert(2 * FC == 32, "staging buffer is 32 x 64 doubles");
        const bool diag = a.symmetric && (tc == tr);
#pragma unroll
        for (int half = 0; half < 2; half++) {  // columns [32*half, 32*half+32) of the tile
            __syncthreads();
            if ((ty >> 3) == half) {
#pragma unroll
                for (int i = 0; i < TT; i++)
#pragma unroll
                    for (int j = 0; j < TT; j++) {
                        const int cl = (ty & 7) * TT + j, rl = tx * TT + i;
                        s_t[cl * TILE + (rl ^ cl)] = acc[i][j];  // XOR swizzle: both read orientations conflict-free
                    }
            }
            __syncthreads();
            // Streaming (evict-first) stores: D is written once and is far larger than L2.  Off the diagonal and with
            // an even leading dimension two neighbouring elements leave as one 16-byte store.
            const bool vec_ok = ((a.ldr & 1) == 0) && !diag && ((reinterpret_cast<size_t>(a.D) & 15) == 0);
            // direct orientation: for each of the 32 columns, 64 contiguous rows
            if (vec_ok) {
                for (int q = tid; q < 32 * (TILE / 2); q += 256) {
                    const int cl = q >> 5, rl = (q & 31) * 2;
                    const long long ri = tr + rl, cj = tc + 32 * half + cl;
                    if (cj >= a.c1 || ri >= a.r1) continue;
                    double d0 = s_t[cl * TILE + (rl ^ cl)], d1 = s_t[cl * TILE + ((rl + 1) ^ cl)];
                    if (a.self_zero) {
                        if (ri == cj) d0 = 0.;
                        if (ri + 1 == cj) d1 = 0.;
                    }
                    if (ri + 1 < a.r1)
                        __stcs(reinterpret_cast<double2*>(a.D + ri + a.ldr * cj), make_double2(d0, d1));
                    else
                        __stcs(a.D + ri + a.ldr * cj, d0);
                }
            } else {
                for (int q = tid; q < 32 * TILE; q += 256) {
                    const int cl = q >> 6, rl = q & 63;
                    const long long ri = tr + rl, cj = tc + 32 * half + cl;
                    if (ri < a.r1 && cj < a.c1) {
                        double d = s_t[cl * TILE + (rl ^ cl)];
                        if (a.symmetric) {
                            if (cj < ri && diag) continue;      // lower part of a diagonal tile comes from the mirror
                            if (cj == ri) d = 0.;
                        }
                        if (a.self_zero && cj == ri) d = 0.;
                        __stcs(a.D + ri + a.ldr * cj, d);
                    }
                }
            }
            // mirrored orientation (symmetric only): D[cj, ri] — contiguous in cj for fixed ri
            if (a.symmetric && vec_ok) {
                for (int q = tid; q < TILE * 16; q += 256) {
                    const int rl = q >> 4, cl = (q & 15) * 2;
                    const long long ri = tr + rl, cj = tc + 32 * half + cl;
                    if (ri >= a.r1 || cj >= a.c1) continue;
                    const double d0 = s_t[cl * TILE + (rl ^ cl)], d1 = s_t[(cl + 1) * TILE + (rl ^ (cl + 1))];
                    if (cj + 1 < a.c1)
                        __stcs(reinterpret_cast<double2*>(a.D + cj + a.ldr * ri), make_double2(d0, d1));
                    else
                        __stcs(a.D + cj + a.ldr * ri, d0);
                }
            } else if (a.symmetric) {
                for (int q = tid; q < 32 * TILE; q += 256) {
                    const int rl = q >> 5, cl = q & 31;
                    const long long ri = tr + rl, cj = tc + 32 * half + cl;
                    if (ri < a.r1 && cj < a.c1 && cj > ri) __stcs(a.D + cj + a.ldr * ri, s_t[cl * TILE + (rl ^ cl)]);
                }
            }
        }
        return;
    }
    if (MODE == TM_COUNT) {
        // neighbour counts aggregated per tile in shared memory: one global atomic per row / column and tile
        __shared__ int s_cnt[2 * TILE];
        __syncthreads();
        if (tid < 2 * TILE) s_cnt[tid] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < TT; i++) {
            const long long ri = tr + tx * TT + i;
            if (ri >= a.r1) continue;
            int local = 0;
#pragma unroll
            for (int j = 0; j < TT; j++) {
                const long long cj = tc + ty * TT + j;
                if (cj >= a.c1) continue;
                if (a.symmetric && cj <= ri) continue;
                const bool self = a.self_zero && ((a.rid ? (long long)a.rid[ri] + a.nbr_row0 : ri) == cj);
                if (!((self ? 0. : acc[i][j]) < a.eps)) continue;  // strict, sparse_clustering.jl:173
                local++;
                if (a.symmetric) atomicAdd(&s_cnt[TILE + ty * TT + j], 1);
                if (a.nbr) {
                    const long long lr = a.rid ? (long long)a.rid[ri] : ri - a.nbr_row0;
                    const int slot = atomicAdd(&a.nfill[lr], 1);
                    if (slot < a.nbr_cap) a.nbr[lr * a.nbr_cap + slot] = (int)cj;
                }
            }
            if (local) atomicAdd(&s_cnt[tx * TT + i], local);
        }
        __syncthreads();
        if (tid < TILE) {
            const long long ri = tr + tid;
            if (s_cnt[tid] && ri < a.r1) atomicAdd(&a.cnt[ri], s_cnt[tid]);
        } else if (tid < 2 * TILE && a.symmetric) {
            const long long cj = tc + (tid - TILE);
            if (s_cnt[tid] && cj < a.c1) atomicAdd(&a.cnt[cj], s_cnt[tid]);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < TT; i++) {
        const long long ri = tr + tx * TT + i;
        if (ri >= a.r1) continue;
#pragma unroll
        for (int j = 0; j < TT; j++) {
            const long long cj = tc + ty * TT + j;
            if (cj >= a.c1) continue;
            const double d = acc[i][j];
            if (a.symmetric && cj <= ri) continue;
            if (!(d < a.eps)) continue;  // strict, sparse_clustering.jl:173
            if (MODE == TM_UNION) {
                uf_union(a.parent, (int)ri, (int)cj);
            } else if (MODE == TM_SPARSE) {  // symmetric launch: the pair lands in both columns
                const int s0 = atomicAdd(&a.sp_fill[cj], 1), s1 = atomicAdd(&a.sp_fill[ri], 1);
                a.sp_row[a.sp_colptr[cj] + s0] = (int)ri;
                a.sp_val[a.sp_colptr[cj] + s0] = d - a.eps;
                a.sp_row[a.sp_colptr[ri] + s1] = (int)cj;
                a.sp_val[a.sp_colptr[ri] + s1] = d - a.eps;
            } else if (MODE == TM_BORDER) {
                atomicMin(&a.best[ri], a.col_root[cj]);
            }
        }
    }
}

}  // namespace mcbb
