// cluster.cu — distance_matrix and DBSCAN entry points (C-ABI hot entries 2 and 3).
//
// DBSCAN is computed as the order-independent restatement of the sequential algorithm the reference uses
// (Clustering.dbscan, vendored at src/sparse_clustering.jl:138-209):
//   pass 1  eps-neighbour counts (self included)              -> core flags
//   pass 2  union-find over core-core edges, larger root hooked under smaller
//           -> components whose root is their smallest member = the sequential algorithm's seed
//   pass 3  every non-core point takes the smallest root among cores within eps (else noise)
//   final   clusters numbered by ascending seed (exclusive scan over "is seed"), counts by histogram
// which yields bit-identical (seeds, assignments, counts) for visit order 1:n.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include <limits.h>

#include <algorithm>
#include <cmath>

#include <vector>

#include "pair_tiles.cuh"
#include "multi.cuh"
#include "topk.cuh"

namespace mcbb {

// dist_mat.cu
int distance_matrix_tma(const double* X, long long n, int F, const double* fw, const unsigned short* fmask, double* D,
                        cudaStream_t st, bool* handled);

// ---- small helper kernels ------------------------------------------------------------------------------
__global__ void k_fill_i32(int* p, long long n, int v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_iota_i32(int* p, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}
__global__ void k_core_flags(const int* cnt, long long n, int minpts, int* core, int* noncore) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int c = cnt[i] >= minpts;
        core[i] = c;
        noncore[i] = !c;
    }
}
// scatter original indices into the compacted core / non-core index lists (stable: ascending index)
__global__ void k_compact_idx(const int* core, const int* core_pos, const int* non_pos, long long n,
                              int* core_idx, int* non_idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (core[i])
            core_idx[core_pos[i]] = (int)i;
        else
            non_idx[non_pos[i]] = (int)i;
    }
}
// Xout[f][k] = X[f][idx[k]]
__global__ void k_gather_cols(const double* X, long long ld, const int* idx, long long m, int F, double* Xout) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (k < m && f < F) Xout[(long long)f * m + k] = X[(long long)f * ld + idx[k]];
}
__global__ void k_flatten_roots(int* parent, long long m, const int* core_idx, int* col_root) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        int x = (int)i;
        while (parent[x] != x) x = parent[x];
        col_root[i] = core_idx[x];  // original index of the component's smallest core = its seed
    }
}
// root_label[orig] = seed index of the point's cluster, or -1 for noise
__global__ void k_root_labels(long long n_core, const int* core_idx, const int* col_root, long long n_non,
                              const int* non_idx, const int* best, int* root_label, int* is_seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_core) {
        const int o = core_idx[i];
        root_label[o] = col_root[i];
        is_seed[o] = (col_root[i] == o);
    } else if (i < n_core + n_non) {
        const long long k = i - n_core;
        const int o = non_idx[k];
        root_label[o] = best[k] == INT_MAX ? -1 : best[k];
        is_seed[o] = 0;
    }
}
__global__ void k_assign(const int* root_label, const int* seed_rank, const int* is_seed, long long n,
                         long long* assignments, long long* seeds, unsigned long long* counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int r = root_label[i];
        long long lab = 0;
        if (r >= 0) {
            lab = (long long)seed_rank[r] + 1;
            atomicAdd(&counts[lab - 1], 1ULL);
        }
        assignments[i] = lab;
        if (is_seed[i]) seeds[seed_rank[i]] = i + 1;  // 1-based like Julia
    }
}
__global__ void k_fill_u64(unsigned long long* p, long long n, unsigned long long v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_flag_below(const int* cnt, long long n, int thr, int* flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = cnt[i] < thr;
}
__global__ void k_compact_flagged(const int* flag, const int* pos, long long n, int* idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) idx[pos[i]] = (int)i;
}
__global__ void k_gather_i32(const int* src, const int* idx, long long m, int* dst) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) dst[k] = src[idx[k]];
}
__global__ void k_scatter_i32(const int* src, const int* idx, long long m, int* dst) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) dst[idx[k]] = src[k];
}

// ---- dense-D passes (HBM streaming over the n x n matrix, column p = neighbourhood of p) ------------------
enum DenseMode { DM_COUNT = 0, DM_UNION = 1, DM_BORDER = 2 };
template <int MODE>
__global__ void __launch_bounds__(256) k_dense_pass(const double* __restrict__ D, long long n, double eps, int* cnt,
                                                    const int* core, int* parent, const int* root, int* best) {
    const long long p = blockIdx.y;  // column
    const long long i0 = (long long)blockIdx.x * 1024;
    if (MODE == DM_UNION && !core[p]) return;
    if (MODE == DM_BORDER && !core[p]) return;
    int local = 0;
    for (int q = 0; q < 4; q++) {
        const long long i = i0 + q * 256 + threadIdx.x;
        if (i >= n) break;
        const double d = D[i + n * p];
        if (!(d < eps)) continue;
        if (MODE == DM_COUNT) {
            local++;
        } else if (MODE == DM_UNION) {
            if (i != p && core[i]) uf_union(parent, (int)i, (int)p);
        } else {
            if (!core[i]) atomicMin(&best[i], root[p]);
        }
    }
    if (MODE == DM_COUNT) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(&cnt[p], local);
    }
}
__global__ void k_flatten_dense(int* parent, const int* core, long long n, int* root, int* is_seed, int* best) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int x = (int)i;
        if (core[i]) {
            while (parent[x] != x) x = parent[x];
            root[i] = x;
            is_seed[i] = (x == i);
        } else {
            root[i] = -1;
            is_seed[i] = 0;
        }
        best[i] = INT_MAX;
    }
}
__global__ void k_labels_dense(const int* core, const int* root, const int* best, long long n, int* root_label) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) root_label[i] = core[i] ? root[i] : (best[i] == INT_MAX ? -1 : best[i]);
}

// ---- host orchestration ------------------------------------------------------------------------------------
struct PassTimer {  // MCBB_TRACE=1: per-pass device times on stderr
    cudaStream_t st;
    cudaEvent_t e0, e1;
    bool on;
    explicit PassTimer(cudaStream_t s) : st(s), on(getenv("MCBB_TRACE") != nullptr || tune_get("trace", 0) != 0) {
        if (on) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, st);
        }
    }
    void lap(const char* name) {
        if (!on) return;
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "[mcbb trace] %-28s %9.3f ms\n", name, ms);
        cudaEventRecord(e0, st);
    }
    ~PassTimer() {
        if (on) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
    }
};

static inline unsigned nblk(long long n, int b = 256) { return (unsigned)((n + b - 1) / b); }

struct GroupTables {
    double* fw;
    unsigned char* fend;
    unsigned short* fmask;
    int F;
    int early_exit;
};

// expands (group_len, group_weight) into per-feature device tables (scratch slots 10/11)
static int upload_groups(int n_groups, const int32_t* group_len, const double* group_weight, GroupTables* gt,
                         cudaStream_t st) {
    int F = 0;
    for (int g = 0; g < n_groups; g++) {
        if (group_len[g] < 1) MCBB_FAIL("distance: every group needs at least one feature");
        F += group_len[g];
    }
    std::vector<double> fw(F);
    std::vector<unsigned char> fe(F);
    int f = 0, ee = 1;
    for (int g = 0; g < n_groups; g++) {
        if (!(group_weight[g] >= 0.)) ee = 0;
        for (int q = 0; q < group_len[g]; q++, f++) {
            fw[f] = group_weight[g];
            fe[f] = (q == group_len[g] - 1);
        }
    }
    std::vector<unsigned short> fm((F + FC - 1) / FC, 0);
    for (int q = 0; q < F; q++)
        if (fe[q]) fm[q / FC] |= (unsigned short)(1u << (q % FC));
    gt->fw = (double*)scratch_get(10, sizeof(double) * F);
    gt->fend = (unsigned char*)scratch_get(11, F);
    gt->fmask = (unsigned short*)scratch_get(58, sizeof(unsigned short) * fm.size());
    if (!gt->fw || !gt->fend || !gt->fmask) MCBB_FAIL("distance: device scratch allocation failed");
    MCBB_CUDA_OK(cudaMemcpyAsync(gt->fw, fw.data(), sizeof(double) * F, cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(gt->fend, fe.data(), F, cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(gt->fmask, fm.data(), sizeof(unsigned short) * fm.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));  // fw / fe are stack-owned
    gt->F = F;
    gt->early_exit = ee;
    return 0;
}

// diagnostic counter of executed pair x feature terms (one per device), enabled by mcbb_pair_feature_counter
static unsigned long long* g_pf_ctr[16] = {};
static unsigned long long* pf_counter() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < 16) ? g_pf_ctr[dev] : nullptr;
}
int pair_feature_counter(int enable, long long* value_out, cudaStream_t st) {
    int dev = 0;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) MCBB_FAIL("pair feature counter: device index out of range");
    if (value_out) {
        *value_out = 0;
        if (g_pf_ctr[dev]) {
            MCBB_CUDA_OK(cudaStreamSynchronize(st));
            unsigned long long h = 0;
            MCBB_CUDA_OK(cudaMemcpy(&h, g_pf_ctr[dev], sizeof(h), cudaMemcpyDeviceToHost));
            *value_out = (long long)h;
        }
    }
    if (enable > 0) {
        unsigned long long* p = (unsigned long long*)scratch_get(68, 64);
        if (!p) MCBB_FAIL("pair feature counter: device allocation failed");
        MCBB_CUDA_OK(cudaMemsetAsync(p, 0, 64, st));
        g_pf_ctr[dev] = p;
    } else if (enable == 0) {
        g_pf_ctr[dev] = nullptr;
    }
    return 0;
}

template <int MODE>
static int launch_tiles(const TileArgs& a_in, cudaStream_t st) {
    TileArgs a = a_in;
    a.pf_ctr = pf_counter();
    const long long nr = a.r1 - a.r0, nc = a.c1 - a.c0;
    if (nr <= 0 || nc <= 0) return 0;
    dim3 grid((unsigned)((nc + TILE - 1) / TILE), (unsigned)((nr + TILE - 1) / TILE));
    // symmetric launch over the whole square: fold the lower triangle away instead of launching blocks that return
    a.fold = (a.symmetric && a.band == 0 && a.col_mod <= 1 && a.r0 == a.c0 && a.r1 == a.c1) ? 1 : 0;
    if (a.fold) grid.y = grid.x / 2 + 1;
    if (a.band > 0) grid.x = a.symmetric ? (unsigned)(a.band + 1) : (unsigned)(2 * a.band + 1);
    if (a.col_mod > 1) grid.x = ((grid.x + a.col_mod - 1) / a.col_mod) * (unsigned)(a.col_hi - a.col_lo);
    if (grid.y > 65535u) MCBB_FAIL("pair tiles: more than 65535 row tiles in one launch; shard the rows");
    k_pair_tiles<MODE><<<grid, 256, 0, st>>>(a);
    MCBB_LAUNCHED();
    return 0;
}

// splits the row range so that gridDim.y stays legal
template <int MODE>
static int launch_tiles_rows(TileArgs a, cudaStream_t st) {
    const long long step = 65535LL * TILE;
    const long long r0 = a.r0, r1 = a.r1;
    for (long long r = r0; r < r1; r += step) {
        a.r0 = r;
        a.r1 = r + step < r1 ? r + step : r1;
        if (int rc = launch_tiles<MODE>(a, st)) return rc;
    }
    return 0;
}


constexpr int BAND_TILES = 4;
constexpr int COL_SAMPLE = 32;

// eps-neighbour counts of rows [row0,row1) against all n columns, exact below minpts and >= minpts otherwise.
// Proving a point NON-core needs its whole row; proving it core needs only minpts hits.  The column tiles are
// therefore visited in four disjoint, evenly spread samples (residues of the tile index mod 32: [0,1) [1,5) [5,16)
// [16,32), i.e. cumulative 1/32, 5/32, 1/2, 1); after each sample the rows that already reached minpts drop out and
// the rest are compacted.  A point with m neighbours leaves after the first sample if m >~ 32 minpts, after the second
// if m >~ 6.4 minpts, ...: work ~ n^2 (1/32 + 4/32 u1 + 11/32 u2 + 16/32 u3) with u_k the still-unsaturated fractions,
// instead of n^2/2.  Counts add up exactly (disjoint column sets).
static int count_rows(const double* X, long long n, const GroupTables& gt, double eps, int minpts, long long row0,
                      long long row1, int* cnt /* row1-row0 */, cudaStream_t st, PassTimer* pt, int* nbr = nullptr,
                      int* nfill = nullptr, int nbr_cap = 0) {
    const long long m = row1 - row0;
    if (m <= 0) return 0;
    k_fill_i32<<<nblk(m), 256, 0, st>>>(cnt, m, 0);
    MCBB_LAUNCHED();
    TileArgs a = {};
    a.Xr = a.Xc = X;
    a.ldr = a.ldc = n;
    a.r0 = row0;
    a.r1 = row1;
    a.c0 = 0;
    a.c1 = n;
    a.F = gt.F;
    a.fw = gt.fw;
    a.fend = gt.fend;
    a.fmask = gt.fmask;
    a.eps = eps;
    a.symmetric = 0;  // full rows: the diagonal pair (d = 0 < eps) supplies the self count
    a.self_zero = 1;
    a.early_exit = gt.early_exit;
    a.cnt = cnt - row0;
    a.sat_count = minpts;
    a.sat_seed = -1;
    a.nbr = nbr;
    a.nfill = nfill;
    a.nbr_cap = nbr_cap;
    a.nbr_row0 = row0;
    if (nbr) {
        k_fill_i32<<<nblk(m), 256, 0, st>>>(nfill, m, 0);
        MCBB_LAUNCHED();
    }
    if (n >= (long long)COL_SAMPLE * TILE * 8) {
        static const int bounds[5] = {0, 1, 5, 16, COL_SAMPLE};
        a.col_mod = COL_SAMPLE;
        a.col_lo = bounds[0];
        a.col_hi = bounds[1];
        if (int rc = launch_tiles_rows<TM_COUNT>(a, st)) return rc;
        if (pt) pt->lap("pass 1 sample 1/32");
        int* flag = (int*)scratch_get(13, sizeof(int) * m);
        int* pos = (int*)scratch_get(15, sizeof(int) * m);
        int* uidx = (int*)scratch_get(17, sizeof(int) * m);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, pos, (int)m, st);
        void* tmp = scratch_get(21, tb);
        if (!flag || !pos || !uidx || !tmp) MCBB_FAIL("dbscan: device scratch allocation failed");
        for (int stage = 1; stage < 4; stage++) {
            k_flag_below<<<nblk(m), 256, 0, st>>>(cnt, m, minpts, flag);
            MCBB_LAUNCHED();
            MCBB_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, flag, pos, (int)m, st));
            g_launch_count.fetch_add(1);
            int h_pos = 0, h_flag = 0;
            MCBB_CUDA_OK(cudaMemcpyAsync(&h_pos, pos + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
            MCBB_CUDA_OK(cudaMemcpyAsync(&h_flag, flag + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
            MCBB_CUDA_OK(cudaStreamSynchronize(st));
            const long long nu = (long long)h_pos + h_flag;
            if (nu == 0) return 0;
            k_compact_flagged<<<nblk(m), 256, 0, st>>>(flag, pos, m, uidx);  // indices relative to row0
            MCBB_LAUNCHED();
            double* Xu = (double*)scratch_get(24, sizeof(double) * (size_t)gt.F * (size_t)nu);
            int* cu = (int*)scratch_get(16, sizeof(int) * (size_t)nu);
            if (!Xu || !cu) MCBB_FAIL("dbscan: device scratch allocation failed");
            k_gather_cols<<<dim3(nblk(nu), gt.F), 256, 0, st>>>(X + row0, n, uidx, nu, gt.F, Xu);
            MCBB_LAUNCHED();
            k_gather_i32<<<nblk(nu), 256, 0, st>>>(cnt, uidx, nu, cu);  // carry the counts so far over
            MCBB_LAUNCHED();
            TileArgs b = a;
            b.Xr = Xu;
            b.ldr = nu;
            b.r0 = 0;
            b.r1 = nu;
            b.cnt = cu;
            b.rid = uidx;  // neighbour lists stay indexed by the original (row0-relative) row
            b.col_lo = bounds[stage];
            b.col_hi = bounds[stage + 1];
            if (int rc = launch_tiles_rows<TM_COUNT>(b, st)) return rc;
            k_scatter_i32<<<nblk(nu), 256, 0, st>>>(cu, uidx, nu, cnt);
            MCBB_LAUNCHED();
            if (pt) pt->lap(stage == 1 ? "pass 1 sample 5/32" : stage == 2 ? "pass 1 sample 1/2" : "pass 1 rest");
        }
        return 0;
    }
    if (int rc = launch_tiles_rows<TM_COUNT>(a, st)) return rc;
    if (pt) pt->lap("pass 1 count");
    return 0;
}

// union-find over the core-core edges (row in [row0,row1), col > row): band pass first, then the full pass, which
// skips every tile whose points already share a root
static int union_rows(const double* Xc, long long n_core, const GroupTables& gt, double eps, long long row0,
                      long long row1, int* parent, cudaStream_t st, PassTimer* pt) {
    if (row1 <= row0) return 0;
    TileArgs u = {};
    u.Xr = u.Xc = Xc;
    u.ldr = u.ldc = n_core;
    u.r0 = row0;
    u.r1 = row1;
    u.c0 = 0;
    u.c1 = n_core;
    u.F = gt.F;
    u.fw = gt.fw;
    u.fend = gt.fend;
    u.fmask = gt.fmask;
    u.eps = eps;
    u.symmetric = 1;
    u.early_exit = gt.early_exit;
    u.parent = parent;
    u.sat_union = 1;
    u.sat_seed = -1;
    if ((row0 % TILE) == 0 && n_core > (2 * BAND_TILES + 1) * TILE * 4) {
        u.band = BAND_TILES;
        if (int rc = launch_tiles_rows<TM_UNION>(u, st)) return rc;
        u.band = 0;
    }
    if (int rc = launch_tiles_rows<TM_UNION>(u, st)) return rc;
    if (pt) pt->lap("pass 2 union");
    return 0;
}

// border points from the neighbour lists recorded in pass 1: smallest seed among the adjacent core points
__global__ void k_border_from_lists(const int* non_idx, long long n_non, const int* nbr, const int* nfill, int cap,
                                    const int* core, const int* core_pos, const int* col_root, int* best) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_non) return;
    const int o = non_idx[k];
    const int m = min(nfill[o], cap);
    int b = INT_MAX;
    for (int s = 0; s < m; s++) {
        const int q = nbr[(long long)o * cap + s];
        if (core[q]) b = min(b, col_root[core_pos[q]]);
    }
    best[k] = b;
}

int distance_matrix_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                        const double* group_weight, double* D, cudaStream_t st) {
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    {  // large, evenly sized problems: the persistent TMA kernel (dist_mat.cu), same arithmetic per pair
        bool handled = false;
        if (int rc = distance_matrix_tma(X, n, gt.F, gt.fw, gt.fmask, D, st, &handled)) return rc;
        if (handled) return 0;
    }
    TileArgs a = {};
    a.Xr = a.Xc = X;
    a.ldr = a.ldc = n;
    a.r0 = a.c0 = 0;
    a.r1 = a.c1 = n;
    a.F = gt.F;
    a.fw = gt.fw;
    a.fend = gt.fend;
    a.fmask = gt.fmask;
    a.symmetric = 1;
    a.D = D;
    a.sat_seed = -1;
    return launch_tiles_rows<TM_MATERIALIZE>(a, st);
}

// rows [row0,row1) of D as (row1-row0) contiguous n-vectors: the unit distance_matrix_mmap writes to its file
// (src/eval_clustering.jl:351-394).  D is symmetric, so the block is produced as columns of the column-major tile
// kernel: out[(c - row0) * n + r] = d(r, c).
int distance_rows_dev(const double* X, long long n, int n_groups, const int32_t* group_len, const double* group_weight,
                      long long row0, long long row1, double* out, cudaStream_t st) {
    if (row0 < 0 || row1 > n || row0 > row1) MCBB_FAIL("distance_matrix: bad row range");
    if (row0 == row1) return 0;
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    TileArgs a = {};
    a.Xr = a.Xc = X;
    a.ldr = a.ldc = n;
    a.r0 = 0;
    a.r1 = n;
    a.c0 = row0;
    a.c1 = row1;
    a.F = gt.F;
    a.fw = gt.fw;
    a.fend = gt.fend;
    a.fmask = gt.fmask;
    a.symmetric = 0;
    a.self_zero = 1;
    a.D = out - n * row0;
    a.sat_seed = -1;
    return launch_tiles_rows<TM_MATERIALIZE>(a, st);
}

// seeds / assignments / counts from root labels (shared by the dense and the feature path)
static int finalize_labels(const int* root_label, const int* is_seed, long long n, long long* assignments,
                           long long* seeds, long long* counts, long long* n_clusters_host, cudaStream_t st) {
    int* seed_rank = (int*)scratch_get(20, sizeof(int) * (n + 1));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, is_seed, seed_rank, (int)n, st);
    void* tmp = scratch_get(21, tb);
    if (!seed_rank || !tmp) MCBB_FAIL("dbscan: device scratch allocation failed");
    MCBB_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, is_seed, seed_rank, (int)n, st));
    g_launch_count.fetch_add(1);
    k_fill_u64<<<nblk(n), 256, 0, st>>>((unsigned long long*)counts, n, 0ULL);
    MCBB_LAUNCHED();
    k_fill_u64<<<nblk(n), 256, 0, st>>>((unsigned long long*)seeds, n, 0ULL);
    MCBB_LAUNCHED();
    k_assign<<<nblk(n), 256, 0, st>>>(root_label, seed_rank, is_seed, n, assignments, seeds,
                                      (unsigned long long*)counts);
    MCBB_LAUNCHED();
    int last_rank = 0, last_flag = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(&last_rank, seed_rank + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(&last_flag, is_seed + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    *n_clusters_host = (long long)last_rank + last_flag;
    return 0;
}

static int check_dbscan_args(long long n, double eps, int minpts) {
    // argument checks of Clustering.dbscan, src/sparse_clustering.jl:126-136
    if (n < 2) MCBB_FAIL("At least two data points are required.");
    if (!(eps > 0)) MCBB_FAIL("eps must be a positive value.");
    if (minpts < 1) MCBB_FAIL("minpts must be positive integer.");
    if (n > 2147483647LL - 1) MCBB_FAIL("dbscan: n exceeds int32 index range");
    return 0;
}

int dbscan_features_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                        const double* group_weight, double eps, int minpts, long long* assignments,
                        long long* seeds, long long* counts, long long* n_clusters_host, cudaStream_t st) {
    if (int rc = check_dbscan_args(n, eps, minpts)) return rc;
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    const int F = gt.F;
    int* cnt = (int*)scratch_get(12, sizeof(int) * n);
    int* core = (int*)scratch_get(13, sizeof(int) * n);
    int* noncore = (int*)scratch_get(14, sizeof(int) * n);
    int* core_pos = (int*)scratch_get(15, sizeof(int) * n);
    int* non_pos = (int*)scratch_get(16, sizeof(int) * n);
    int* core_idx = (int*)scratch_get(17, sizeof(int) * n);
    int* non_idx = (int*)scratch_get(18, sizeof(int) * n);
    int* root_label = (int*)scratch_get(19, sizeof(int) * n);
    int* is_seed = (int*)scratch_get(22, sizeof(int) * n);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, core, core_pos, (int)n, st);
    void* tmp = scratch_get(21, tb);
    if (!cnt || !core || !noncore || !core_pos || !non_pos || !core_idx || !non_idx || !root_label || !is_seed ||
        !tmp)
        MCBB_FAIL("dbscan: device scratch allocation failed");

    PassTimer pt(st);
    // pass 1: neighbour counts (the diagonal pair, D_pp = 0 < eps, supplies the self count)
    // neighbour lists of the (eventual) non-core points: minpts entries per point are enough
    int *nbr = nullptr, *nfill = nullptr;
    if ((size_t)n * (size_t)minpts * sizeof(int) <= ((size_t)512 << 20) && !tune_get("dbscan_no_lists", 0)) {
        nbr = (int*)scratch_get(34, sizeof(int) * (size_t)n * (size_t)minpts);
        nfill = (int*)scratch_get(35, sizeof(int) * (size_t)n);
        if (!nbr || !nfill) MCBB_FAIL("dbscan: device scratch allocation failed");
    }
    if (int rc = count_rows(X, n, gt, eps, minpts, 0, n, cnt, st, &pt, nbr, nfill, minpts)) return rc;
    TileArgs a = {};
    a.F = F;
    a.fw = gt.fw;
    a.fend = gt.fend;
    a.fmask = gt.fmask;
    a.eps = eps;
    a.early_exit = gt.early_exit;
    a.sat_seed = -1;

    // core flags and stable compaction
    k_core_flags<<<nblk(n), 256, 0, st>>>(cnt, n, minpts, core, noncore);
    MCBB_LAUNCHED();
    MCBB_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, core, core_pos, (int)n, st));
    MCBB_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, noncore, non_pos, (int)n, st));
    g_launch_count.fetch_add(2);
    int h_pos = 0, h_flag = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(&h_pos, core_pos + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(&h_flag, core + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    const long long n_core = (long long)h_pos + h_flag, n_non = n - n_core;
    k_compact_idx<<<nblk(n), 256, 0, st>>>(core, core_pos, non_pos, n, core_idx, non_idx);
    MCBB_LAUNCHED();

    double* Xc = (double*)scratch_get(23, sizeof(double) * (size_t)F * (size_t)(n_core > 0 ? n_core : 1));
    double* Xn = (double*)scratch_get(24, sizeof(double) * (size_t)F * (size_t)(n_non > 0 ? n_non : 1));
    int* parent = (int*)scratch_get(25, sizeof(int) * (size_t)(n_core > 0 ? n_core : 1));
    int* col_root = (int*)scratch_get(26, sizeof(int) * (size_t)(n_core > 0 ? n_core : 1));
    int* best = (int*)scratch_get(27, sizeof(int) * (size_t)(n_non > 0 ? n_non : 1));
    if (!Xc || !Xn || !parent || !col_root || !best) MCBB_FAIL("dbscan: device scratch allocation failed");

    if (n_core > 0) {
        k_gather_cols<<<dim3(nblk(n_core), F), 256, 0, st>>>(X, n, core_idx, n_core, F, Xc);
        MCBB_LAUNCHED();
        // pass 2: components of the core graph
        k_iota_i32<<<nblk(n_core), 256, 0, st>>>(parent, n_core);
        MCBB_LAUNCHED();
        if (int rc = union_rows(Xc, n_core, gt, eps, 0, n_core, parent, st, &pt)) return rc;
        k_flatten_roots<<<nblk(n_core), 256, 0, st>>>(parent, n_core, core_idx, col_root);
        MCBB_LAUNCHED();
    }
    if (n_non > 0) {
        k_fill_i32<<<nblk(n_non), 256, 0, st>>>(best, n_non, INT_MAX);
        MCBB_LAUNCHED();
        if (n_core > 0 && nbr) {
            k_border_from_lists<<<nblk(n_non), 256, 0, st>>>(non_idx, n_non, nbr, nfill, minpts, core, core_pos, col_root, best);
            MCBB_LAUNCHED();
            pt.lap("border from the recorded neighbour lists");
        } else if (n_core > 0) {
            k_gather_cols<<<dim3(nblk(n_non), F), 256, 0, st>>>(X, n, non_idx, n_non, F, Xn);
            MCBB_LAUNCHED();
            // pass 3: border points
            TileArgs b = a;
            b.Xr = Xn;
            b.ldr = n_non;
            b.r0 = 0;
            b.r1 = n_non;
            b.Xc = Xc;
            b.ldc = n_core;
            b.c0 = 0;
            b.c1 = n_core;
            b.symmetric = 0;
            b.cnt = nullptr;
            b.col_root = col_root;
            b.best = best;
            {   // the globally smallest seed is the seed of the first core point (seeds are component minima)
                int h_seed = -1;
                MCBB_CUDA_OK(cudaMemcpyAsync(&h_seed, col_root, sizeof(int), cudaMemcpyDeviceToHost, st));
                MCBB_CUDA_OK(cudaStreamSynchronize(st));
                b.sat_seed = h_seed;
            }
            if (int rc = launch_tiles_rows<TM_BORDER>(b, st)) return rc;
            pt.lap("pass 3 border");
        }
    }
    k_root_labels<<<nblk(n), 256, 0, st>>>(n_core, core_idx, col_root, n_non, non_idx, best, root_label, is_seed);
    MCBB_LAUNCHED();
    return finalize_labels(root_label, is_seed, n, assignments, seeds, counts, n_clusters_host, st);
}

// ---- several devices in one process: the per-rank body of mcbb_dbscan_features_multi (csrc/multi.cu) ------------------
// Every rank holds the whole feature matrix X [F][n] and produces identical labels.  The pair space is sharded by
// INTERLEAVED row blocks (rank r owns blocks b = r, r + world, ... of MULTI_BLOCKS_PER_RANK * world equal blocks): runs
// arrive sorted by parameter and the cost of a row varies strongly along that order, so contiguous shards would be
// unbalanced, while interleaving needs no permutation of X.  Exchanges (all O(n) integers):
//   counts           every rank fills the rows it owns of a zeroed full-length array        -> all-reduce SUM
//   union-find       per-rank parent array over its blocks of the core pair triangle         -> all-gather + merge
//   border labels    owned non-core rows from their recorded neighbour lists (INT_MAX else)  -> all-reduce MIN
constexpr int MULTI_BLOCKS_PER_RANK = 16;
__global__ void k_merge_parents(int* parent, const int* other, long long m);

__global__ void k_border_owned_rows(const int* core, const int* core_pos, const int* col_root, const int* nbr,
                                    const int* nfill, int cap, long long r0, long long r1, int* best) {
    const long long i = r0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r1 || core[i]) return;
    const int m = min(nfill[i], cap);
    int b = INT_MAX;
    for (int s = 0; s < m; s++) {
        const int q = nbr[i * cap + s];
        if (core[q]) b = min(b, col_root[core_pos[q]]);
    }
    best[i] = b;
}
__global__ void k_root_labels_full(const int* core, const int* core_pos, const int* col_root, const int* best, long long n,
                                   int* root_label, int* is_seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (core[i]) {
        const int r = col_root[core_pos[i]];
        root_label[i] = r;
        is_seed[i] = (r == (int)i);
    } else {
        root_label[i] = best[i] == INT_MAX ? -1 : best[i];
        is_seed[i] = 0;
    }
}

// row block [lo, hi) number k of `parts` equal-AREA blocks of the pair triangle {col > row} over m points (tile aligned)
static void triangle_block(long long m, int k, int parts, long long* lo, long long* hi) {
    auto edge = [&](int q) -> long long {
        if (q >= parts) return m;
        const long long b = (long long)((double)m * (1.0 - std::sqrt(1.0 - (double)q / (double)parts)));
        return std::min<long long>(m, (b / TILE) * TILE);
    };
    *lo = edge(k);
    *hi = edge(k + 1);
}

int dbscan_features_multi_rank(const double* X, long long n, int n_groups, const int32_t* group_len,
                               const double* group_weight, double eps, int minpts, const MultiComm& mc, long long* assignments,
                               long long* seeds, long long* counts, long long* n_clusters_host, cudaStream_t st) {
    if (int rc = check_dbscan_args(n, eps, minpts)) return rc;
    if ((size_t)n * (size_t)minpts * sizeof(int) > ((size_t)2 << 30))
        MCBB_FAIL("dbscan (multi-device): n * minpts neighbour lists exceed 2 GB; lower minpts or use fewer points");
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    const int F = gt.F;
    const int world = mc.world, rank = mc.rank;
    int* cnt = (int*)scratch_get(12, sizeof(int) * n);
    int* core = (int*)scratch_get(75, sizeof(int) * n);
    int* noncore = (int*)scratch_get(76, sizeof(int) * n);
    int* core_pos = (int*)scratch_get(77, sizeof(int) * n);
    int* non_pos = (int*)scratch_get(78, sizeof(int) * n);
    int* core_idx = (int*)scratch_get(79, sizeof(int) * n);
    int* non_idx = (int*)scratch_get(18, sizeof(int) * n);
    int* root_label = (int*)scratch_get(19, sizeof(int) * n);
    int* is_seed = (int*)scratch_get(22, sizeof(int) * n);
    int* best = (int*)scratch_get(27, sizeof(int) * n);
    int* nbr = (int*)scratch_get(34, sizeof(int) * (size_t)n * (size_t)minpts);
    int* nfill = (int*)scratch_get(35, sizeof(int) * (size_t)n);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, core, core_pos, (int)n, st);
    void* tmp = scratch_get(80, tb);
    if (!cnt || !core || !noncore || !core_pos || !non_pos || !core_idx || !non_idx || !root_label || !is_seed || !best ||
        !nbr || !nfill || !tmp)
        MCBB_FAIL("dbscan: device scratch allocation failed");
    PassTimer pt(st);
    // ---- stage 1: neighbour counts (+ the first minpts neighbours) of the owned row blocks -----------------
    k_fill_i32<<<nblk(n), 256, 0, st>>>(cnt, n, 0);
    MCBB_LAUNCHED();
    k_fill_i32<<<nblk(n), 256, 0, st>>>(nfill, n, 0);
    MCBB_LAUNCHED();
    const int n_blocks = world * MULTI_BLOCKS_PER_RANK;
    const long long bs = ((n + n_blocks - 1) / n_blocks + TILE - 1) / TILE * TILE;  // tile-aligned block size
    for (long long b = rank; b * bs < n; b += world) {
        const long long r0 = b * bs, r1 = std::min<long long>(n, r0 + bs);
        if (int rc = count_rows(X, n, gt, eps, minpts, r0, r1, cnt + r0, st, nullptr, nbr + r0 * minpts, nfill + r0, minpts))
            return rc;
    }
    pt.lap("multi: count owned row blocks");
    if (int rc = mc.allreduce_i32(mc.ctx, cnt, n, 0, st)) return rc;
    pt.lap("multi: all-reduce counts");
    // ---- core flags and stable compaction (replicated: O(n)) ---------------------------------------------------
    k_core_flags<<<nblk(n), 256, 0, st>>>(cnt, n, minpts, core, noncore);
    MCBB_LAUNCHED();
    MCBB_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, core, core_pos, (int)n, st));
    MCBB_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, noncore, non_pos, (int)n, st));
    g_launch_count.fetch_add(2);
    int h_pos = 0, h_flag = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(&h_pos, core_pos + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(&h_flag, core + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    const long long n_core = (long long)h_pos + h_flag;
    k_compact_idx<<<nblk(n), 256, 0, st>>>(core, core_pos, non_pos, n, core_idx, non_idx);
    MCBB_LAUNCHED();
    int* col_root = (int*)scratch_get(26, sizeof(int) * (size_t)(n_core > 0 ? n_core : 1));
    if (!col_root) MCBB_FAIL("dbscan: device scratch allocation failed");
    if (n_core > 0) {
        // ---- stage 2: components of the core graph; the pair triangle in 4 * world equal-area blocks, dealt cyclically
        double* Xc = (double*)scratch_get(23, sizeof(double) * (size_t)F * (size_t)n_core);
        int* parent = (int*)scratch_get(25, sizeof(int) * (size_t)n_core);
        int* others = (int*)scratch_get(81, sizeof(int) * (size_t)n_core * (size_t)world);
        if (!Xc || !parent || !others) MCBB_FAIL("dbscan: device scratch allocation failed");
        k_gather_cols<<<dim3(nblk(n_core), F), 256, 0, st>>>(X, n, core_idx, n_core, F, Xc);
        MCBB_LAUNCHED();
        k_iota_i32<<<nblk(n_core), 256, 0, st>>>(parent, n_core);
        MCBB_LAUNCHED();
        const int parts = 4 * world;
        for (int k = rank; k < parts; k += world) {
            long long c0, c1;
            triangle_block(n_core, k, parts, &c0, &c1);
            if (int rc = union_rows(Xc, n_core, gt, eps, c0, c1, parent, st, nullptr)) return rc;
        }
        pt.lap("multi: union owned triangle blocks");
        if (world > 1) {
            if (int rc = mc.allgather_i32(mc.ctx, parent, others, n_core, st)) return rc;
            for (int r = 0; r < world; r++) {
                if (r == rank) continue;
                k_merge_parents<<<nblk(n_core), 256, 0, st>>>(parent, others + (size_t)r * n_core, n_core);
                MCBB_LAUNCHED();
            }
        }
        pt.lap("multi: all-gather + merge parents");
        // cores are compacted in ascending original order, so the smallest position of a component is its seed
        k_flatten_roots<<<nblk(n_core), 256, 0, st>>>(parent, n_core, core_idx, col_root);
        MCBB_LAUNCHED();
    }
    // ---- stage 3: border labels of the owned non-core rows from their neighbour lists ---------------------------
    k_fill_i32<<<nblk(n), 256, 0, st>>>(best, n, INT_MAX);
    MCBB_LAUNCHED();
    if (n_core > 0 && n_core < n) {
        for (long long b = rank; b * bs < n; b += world) {
            const long long r0 = b * bs, r1 = std::min<long long>(n, r0 + bs);
            k_border_owned_rows<<<nblk(r1 - r0), 256, 0, st>>>(core, core_pos, col_root, nbr, nfill, minpts, r0, r1, best);
            MCBB_LAUNCHED();
        }
        if (int rc = mc.allreduce_i32(mc.ctx, best, n, 1, st)) return rc;
    }
    pt.lap("multi: border labels");
    k_root_labels_full<<<nblk(n), 256, 0, st>>>(core, core_pos, col_root, best, n, root_label, is_seed);
    MCBB_LAUNCHED();
    return finalize_labels(root_label, is_seed, n, assignments, seeds, counts, n_clusters_host, st);
}

int dbscan_dense_dev(const double* D, long long n, double eps, int minpts, long long* assignments,
                     long long* seeds, long long* counts, long long* n_clusters_host, cudaStream_t st) {
    if (int rc = check_dbscan_args(n, eps, minpts)) return rc;
    if (n > 65535) MCBB_FAIL("dbscan_dense: n > 65535 does not fit a materialised matrix launch; use dbscan_features");
    int* cnt = (int*)scratch_get(12, sizeof(int) * n);
    int* core = (int*)scratch_get(13, sizeof(int) * n);
    int* noncore = (int*)scratch_get(14, sizeof(int) * n);
    int* parent = (int*)scratch_get(25, sizeof(int) * n);
    int* root = (int*)scratch_get(26, sizeof(int) * n);
    int* best = (int*)scratch_get(27, sizeof(int) * n);
    int* root_label = (int*)scratch_get(19, sizeof(int) * n);
    int* is_seed = (int*)scratch_get(22, sizeof(int) * n);
    if (!cnt || !core || !noncore || !parent || !root || !best || !root_label || !is_seed)
        MCBB_FAIL("dbscan: device scratch allocation failed");
    dim3 grid((unsigned)((n + 1023) / 1024), (unsigned)n);
    k_fill_i32<<<nblk(n), 256, 0, st>>>(cnt, n, 0);
    MCBB_LAUNCHED();
    k_dense_pass<DM_COUNT><<<grid, 256, 0, st>>>(D, n, eps, cnt, nullptr, nullptr, nullptr, nullptr);
    MCBB_LAUNCHED();
    k_core_flags<<<nblk(n), 256, 0, st>>>(cnt, n, minpts, core, noncore);
    MCBB_LAUNCHED();
    k_iota_i32<<<nblk(n), 256, 0, st>>>(parent, n);
    MCBB_LAUNCHED();
    k_dense_pass<DM_UNION><<<grid, 256, 0, st>>>(D, n, eps, nullptr, core, parent, nullptr, nullptr);
    MCBB_LAUNCHED();
    k_flatten_dense<<<nblk(n), 256, 0, st>>>(parent, core, n, root, is_seed, best);
    MCBB_LAUNCHED();
    k_dense_pass<DM_BORDER><<<grid, 256, 0, st>>>(D, n, eps, nullptr, core, nullptr, root, best);
    MCBB_LAUNCHED();
    k_labels_dense<<<nblk(n), 256, 0, st>>>(core, root, best, n, root_label);
    MCBB_LAUNCHED();
    return finalize_labels(root_label, is_seed, n, assignments, seeds, counts, n_clusters_host, st);
}

// ---- rank-sharded stages ------------------------------------------------------------------------------------
__global__ void k_merge_parents(int* parent, const int* other, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const int o = other[i];
        if (o != (int)i) uf_union(parent, (int)i, o);
    }
}
__global__ void k_seed_flags(const int* root_label, long long n, int* is_seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_seed[i] = (root_label[i] == (int)i);
}

int rows_count_dev(const double* X, long long n, int n_groups, const int32_t* group_len, const double* group_weight,
                   double eps, int minpts, long long row0, long long row1, int* count, int* nbr, int* nfill,
                   cudaStream_t st) {
    if (int rc = check_dbscan_args(n, eps, 1)) return rc;
    if (row0 < 0 || row1 > n || row0 > row1) MCBB_FAIL("dbscan rows: bad row range");
    if ((nbr == nullptr) != (nfill == nullptr)) MCBB_FAIL("dbscan rows: neighbour lists need both arrays");
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    PassTimer pt(st);
    return count_rows(X, n, gt, eps, minpts, row0, row1, count, st, &pt, nbr, nfill, minpts);
}

int rows_union_dev(const double* Xc, long long n_core, int n_groups, const int32_t* group_len,
                   const double* group_weight, double eps, long long row0, long long row1, int* parent, int reset,
                   cudaStream_t st) {
    if (n_core <= 0) return 0;
    if (row0 < 0 || row1 > n_core || row0 > row1) MCBB_FAIL("dbscan rows: bad row range");
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    if (reset) {
        k_iota_i32<<<nblk(n_core), 256, 0, st>>>(parent, n_core);
        MCBB_LAUNCHED();
    }
    PassTimer pt(st);
    return union_rows(Xc, n_core, gt, eps, row0, row1, parent, st, &pt);
}

int merge_parents_dev(int* parent, const int* other, long long n_core, cudaStream_t st) {
    if (n_core <= 0) return 0;
    k_merge_parents<<<nblk(n_core), 256, 0, st>>>(parent, other, n_core);
    MCBB_LAUNCHED();
    return 0;
}

int rows_border_dev(const double* Xn, long long n_non, long long row0, long long row1, const double* Xc,
                    long long n_core, const int* core_seed, int n_groups, const int32_t* group_len,
                    const double* group_weight, double eps, int* best, cudaStream_t st) {
    if (row0 < 0 || row1 > n_non || row0 > row1) MCBB_FAIL("dbscan rows: bad row range");
    if (row1 == row0) return 0;
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    k_fill_i32<<<nblk(row1 - row0), 256, 0, st>>>(best, row1 - row0, INT_MAX);
    MCBB_LAUNCHED();
    if (n_core <= 0) return 0;
    TileArgs a = {};
    a.Xr = Xn;
    a.ldr = n_non;
    a.r0 = row0;
    a.r1 = row1;
    a.Xc = Xc;
    a.ldc = n_core;
    a.c0 = 0;
    a.c1 = n_core;
    a.F = gt.F;
    a.fw = gt.fw;
    a.fend = gt.fend;
    a.fmask = gt.fmask;
    a.eps = eps;
    a.symmetric = 0;
    a.early_exit = gt.early_exit;
    a.col_root = core_seed;
    a.best = best - row0;
    {
        int h_seed = -1;  // seed of the first core point = the globally smallest seed
        MCBB_CUDA_OK(cudaMemcpyAsync(&h_seed, core_seed, sizeof(int), cudaMemcpyDeviceToHost, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        a.sat_seed = h_seed;
    }
    return launch_tiles_rows<TM_BORDER>(a, st);
}

int labels_dev(const int* root_label, long long n, long long* assignments, long long* seeds, long long* counts,
               long long* n_clusters_host, cudaStream_t st) {
    if (n < 1) MCBB_FAIL("dbscan labels: empty input");
    int* is_seed = (int*)scratch_get(22, sizeof(int) * n);
    if (!is_seed) MCBB_FAIL("dbscan: device scratch allocation failed");
    k_seed_flags<<<nblk(n), 256, 0, st>>>(root_label, n, is_seed);
    MCBB_LAUNCHED();
    return finalize_labels(root_label, is_seed, n, assignments, seeds, counts, n_clusters_host, st);
}

int gather_columns_dev(const double* X, long long ld, const int* idx, long long m, int F, double* Xout,
                       cudaStream_t st) {
    if (m <= 0 || F <= 0) return 0;
    k_gather_cols<<<dim3(nblk(m), F), 256, 0, st>>>(X, ld, idx, m, F, Xout);
    MCBB_LAUNCHED();
    return 0;
}

// ---- sparse interchange: NonzeroSparseMatrix (src/sparse_clustering.jl:37-110, built column by column in
// src/eval_clustering.jl:541-550): entries with d < threshold stored as d - threshold in compressed columns ----------
__global__ void k_add_i32(int* p, long long n, int v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += v;
}
__global__ void k_sparse_diag(const long long* colptr, int* fill, int* row, double* val, long long n, double thr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {  // D_ii = 0 < threshold: stored as -threshold
        const int s = atomicAdd(&fill[i], 1);
        row[colptr[i] + s] = (int)i;
        val[colptr[i] + s] = 0. - thr;
    }
}

static int sparse_tile_args(TileArgs* a, const double* X, long long n, const GroupTables& gt, double thr) {
    *a = TileArgs{};
    a->Xr = a->Xc = X;
    a->ldr = a->ldc = n;
    a->r0 = a->c0 = 0;
    a->r1 = a->c1 = n;
    a->F = gt.F;
    a->fw = gt.fw;
    a->fend = gt.fend;
    a->fmask = gt.fmask;
    a->eps = thr;
    a->symmetric = 1;
    a->early_exit = gt.early_exit;
    a->sat_seed = -1;
    return 0;
}

// count[c] = number of rows r with D[r,c] < threshold (diagonal included)
int distance_sparse_count_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                              const double* group_weight, double thr, int* count, cudaStream_t st) {
    if (n < 1) MCBB_FAIL("distance_matrix: empty input");
    if (!(thr > 0)) MCBB_FAIL("sparse_threshold must be positive");
    if (n > 2147483647LL - 1) MCBB_FAIL("distance_matrix: n exceeds int32 index range");
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    k_fill_i32<<<nblk(n), 256, 0, st>>>(count, n, 1);  // the diagonal
    MCBB_LAUNCHED();
    TileArgs a;
    sparse_tile_args(&a, X, n, gt, thr);
    a.cnt = count;
    return launch_tiles_rows<TM_COUNT>(a, st);
}

// fills rowval (0-based, ascending within each column) and nzval = d - threshold for the layout given by colptr
int distance_sparse_fill_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                             const double* group_weight, double thr, const long long* colptr, long long nnz,
                             int* rowval, double* nzval, cudaStream_t st) {
    if (n < 1) MCBB_FAIL("distance_matrix: empty input");
    if (!(thr > 0)) MCBB_FAIL("sparse_threshold must be positive");
    if (nnz > 2147483647LL - 1) MCBB_FAIL("sparse distance matrix: more than 2^31 stored entries");
    GroupTables gt;
    if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
    int* fill = (int*)scratch_get(12, sizeof(int) * n);
    int* row_u = (int*)scratch_get(28, sizeof(int) * (size_t)nnz);
    double* val_u = (double*)scratch_get(29, sizeof(double) * (size_t)nnz);
    if (!fill || !row_u || !val_u) MCBB_FAIL("sparse distance matrix: device scratch allocation failed");
    k_fill_i32<<<nblk(n), 256, 0, st>>>(fill, n, 0);
    MCBB_LAUNCHED();
    k_sparse_diag<<<nblk(n), 256, 0, st>>>(colptr, fill, row_u, val_u, n, thr);
    MCBB_LAUNCHED();
    TileArgs a;
    sparse_tile_args(&a, X, n, gt, thr);
    a.sp_colptr = colptr;
    a.sp_fill = fill;
    a.sp_row = row_u;
    a.sp_val = val_u;
    if (int rc = launch_tiles_rows<TM_SPARSE>(a, st)) return rc;
    // rows ascending within every column, like SparseMatrixCSC
    size_t tb = 0;
    cub::DeviceSegmentedSort::SortPairs(nullptr, tb, row_u, rowval, val_u, nzval, (int)nnz, (int)n, colptr, colptr + 1, st);
    void* tmp = scratch_get(47, tb);
    if (!tmp) MCBB_FAIL("sparse distance matrix: device scratch allocation failed");
    MCBB_CUDA_OK(cub::DeviceSegmentedSort::SortPairs(tmp, tb, row_u, rowval, val_u, nzval, (int)nnz, (int)n, colptr,
                                                     colptr + 1, st));
    g_launch_count.fetch_add(1);
    return 0;
}

// DBSCAN on a NonzeroSparseMatrix (src/sparse_clustering.jl:126-215): unstored entries read as `value`
__global__ void k_sp_count(const long long* colptr, const double* val, long long n, double value, double eps, int* cnt) {
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n) return;
    int k = 0;
    for (long long e = colptr[c] + lane; e < colptr[c + 1]; e += 32) k += (val[e] + value < eps);
    for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
    if (lane == 0) cnt[c] = k;
}
template <int PASS>  // 0: union over core-core edges, 1: smallest adjacent seed of the non-core columns
__global__ void k_sp_edges(const long long* colptr, const int* row, const double* val, long long n, double value,
                           double eps, const int* core, int* parent, int* best) {
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n) return;
    if (PASS == 0 && !core[c]) return;
    if (PASS == 1 && core[c]) return;
    for (long long e = colptr[c] + lane; e < colptr[c + 1]; e += 32) {
        const int r = row[e];
        if (!(val[e] + value < eps) || !core[r]) continue;
        if (PASS == 0) {
            if (r < (int)c) uf_union(parent, r, (int)c);
        } else {
            atomicMin(&best[c], uf_find(parent, r));
        }
    }
}
__global__ void k_sp_labels(const int* core, int* parent, const int* best, long long n, int* root_label, int* is_seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (core[i]) {
            const int r = uf_find(parent, (int)i);
            root_label[i] = r;
            is_seed[i] = (r == (int)i);
        } else {
            root_label[i] = best[i] == INT_MAX ? -1 : best[i];
            is_seed[i] = 0;
        }
    }
}
__global__ void k_all_one_cluster(long long n, int* root_label, int* is_seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        root_label[i] = 0;
        is_seed[i] = (i == 0);
    }
}

int dbscan_sparse_dev(long long n, const long long* colptr, const int* rowval, const double* nzval, double value,
                      double eps, int minpts, long long* assignments, long long* seeds, long long* counts,
                      long long* n_clusters_host, cudaStream_t st) {
    if (int rc = check_dbscan_args(n, eps, minpts)) return rc;
    int* cnt = (int*)scratch_get(12, sizeof(int) * n);
    int* core = (int*)scratch_get(13, sizeof(int) * n);
    int* noncore = (int*)scratch_get(14, sizeof(int) * n);
    int* root_label = (int*)scratch_get(19, sizeof(int) * n);
    int* is_seed = (int*)scratch_get(22, sizeof(int) * n);
    int* parent = (int*)scratch_get(25, sizeof(int) * n);
    int* best = (int*)scratch_get(27, sizeof(int) * n);
    if (!cnt || !core || !noncore || !root_label || !is_seed || !parent || !best)
        MCBB_FAIL("dbscan: device scratch allocation failed");
    if (value < eps) {
        // every unstored pair reads as `value` < eps and every stored one is below `value`: all points are mutual
        // neighbours, exactly what the reference's region query returns
        if (n >= minpts) {
            k_all_one_cluster<<<nblk(n), 256, 0, st>>>(n, root_label, is_seed);
        } else {
            k_fill_i32<<<nblk(n), 256, 0, st>>>(root_label, n, -1);
            MCBB_LAUNCHED();
            k_fill_i32<<<nblk(n), 256, 0, st>>>(is_seed, n, 0);
        }
        MCBB_LAUNCHED();
        return finalize_labels(root_label, is_seed, n, assignments, seeds, counts, n_clusters_host, st);
    }
    const unsigned wb = nblk(n * 32);
    k_sp_count<<<wb, 256, 0, st>>>(colptr, nzval, n, value, eps, cnt);
    MCBB_LAUNCHED();
    k_core_flags<<<nblk(n), 256, 0, st>>>(cnt, n, minpts, core, noncore);
    MCBB_LAUNCHED();
    k_iota_i32<<<nblk(n), 256, 0, st>>>(parent, n);
    MCBB_LAUNCHED();
    k_fill_i32<<<nblk(n), 256, 0, st>>>(best, n, INT_MAX);
    MCBB_LAUNCHED();
    k_sp_edges<0><<<wb, 256, 0, st>>>(colptr, rowval, nzval, n, value, eps, core, parent, best);
    MCBB_LAUNCHED();
    k_sp_edges<1><<<wb, 256, 0, st>>>(colptr, rowval, nzval, n, value, eps, core, parent, best);
    MCBB_LAUNCHED();
    k_sp_labels<<<nblk(n), 256, 0, st>>>(core, parent, best, n, root_label, is_seed);
    MCBB_LAUNCHED();
    return finalize_labels(root_label, is_seed, n, assignments, seeds, counts, n_clusters_host, st);
}

// ---- k_dist / KNN_dist (src/eval_clustering.jl:1504-1553): every row of D sorted ascending -----------------------
__global__ void k_row_offsets(long long* off, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) off[i] = i * n;
}
__global__ void k_row_offsets_ld(long long* off, long long m, long long ld) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= m) off[i] = i * ld;
}
// kth[i] = sorted_row_i[k-1];  cum[i] = sum of the K smallest entries of row i (the zero self-distance included,
// exactly like `sort(D[i,:])[1:K]` in the reference)
__global__ void k_knn_pick(const double* sorted, long long n, int k, int K, double* kth, double* cum,
                           long long n_rows = -1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (n_rows < 0 ? n : n_rows)) {
        const double* row = sorted + i * n;
        if (kth) kth[i] = row[k - 1];
        if (cum) {
            double s = 0.;
            for (int q = 0; q < K; q++) s += row[q];
            cum[i] = s;
        }
    }
}

int knn_dist_dense_dev(const double* D, long long n, int k, int K, double* kth, double* cum, cudaStream_t st) {
    if (n < 1) MCBB_FAIL("k_dist: empty matrix");
    if ((kth && (k < 1 || k > n)) || (cum && (K < 1 || K > n))) MCBB_FAIL("k_dist: k out of range");
    if (n > 46000) MCBB_FAIL("k_dist: n too large for a materialised sort");
    double* sorted = (double*)scratch_get(45, sizeof(double) * (size_t)n * (size_t)n);
    long long* off = (long long*)scratch_get(46, sizeof(long long) * (size_t)(n + 1));
    if (!sorted || !off) MCBB_FAIL("k_dist: device scratch allocation failed");
    k_row_offsets<<<nblk(n + 1), 256, 0, st>>>(off, n);
    MCBB_LAUNCHED();
    size_t tb = 0;
    // D is symmetric: row i == column i, contiguous in the column-major layout
    cub::DeviceSegmentedSort::SortKeys(nullptr, tb, D, sorted, (int)(n * n), (int)n, off, off + 1, st);
    void* tmp = scratch_get(47, tb);
    if (!tmp) MCBB_FAIL("k_dist: device scratch allocation failed");
    MCBB_CUDA_OK(cub::DeviceSegmentedSort::SortKeys(tmp, tb, D, sorted, (int)(n * n), (int)n, off, off + 1, st));
    g_launch_count.fetch_add(1);
    k_knn_pick<<<nblk(n), 256, 0, st>>>(sorted, n, k, K, kth, cum);
    MCBB_LAUNCHED();
    return 0;
}

// The same from the FEATURES, for ensembles whose D does not fit (C3-C5): blocks of rows of D are produced by the
// pair tiles (distance_rows_dev), sorted per row and reduced to (k-th, cumulative-K); only one block is ever resident.
// block_rows = 0 picks a block of about 2 GB.
int knn_dist_features_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                          const double* group_weight, int k, int K, long long block_rows, double* kth, double* cum,
                          cudaStream_t st) {
    if (n < 1) MCBB_FAIL("k_dist: empty input");
    if ((kth && (k < 1 || k > n)) || (cum && (K < 1 || K > n))) MCBB_FAIL("k_dist: k out of range");
    {   // fused per-row K-smallest (topk.cuh): no row of D is ever materialised or sorted
        const int KK = std::max(kth ? k : 1, cum ? K : 1);
        if (KK <= TOPK_MAX && !tune_get("no_fused_topk", 0)) {
            GroupTables gt;
            if (int rc = upload_groups(n_groups, group_len, group_weight, &gt, st)) return rc;
            TopkArgs a = {};
            a.X = X;
            a.n = n;
            a.F = gt.F;
            a.fw = gt.fw;
            a.fmask = gt.fmask;
            a.early_exit = gt.early_exit;
            a.KK = KK;
            a.k = k;
            a.K = K;
            MCBB_CUDA_OK(cudaFuncSetAttribute(k_topk_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TOPK_SMEM));
            const long long step = 65535LL * TILE;
            for (long long r0 = 0; r0 < n; r0 += step) {
                a.r0 = r0;
                a.r1 = std::min(n, r0 + step);
                a.kth = kth ? kth + r0 : nullptr;
                a.cum = cum ? cum + r0 : nullptr;
                k_topk_rows<<<(unsigned)((a.r1 - a.r0 + TILE - 1) / TILE), 256, TOPK_SMEM, st>>>(a);
                MCBB_LAUNCHED();
            }
            return 0;
        }
    }
    if (block_rows <= 0) block_rows = std::max<long long>(1, (1LL << 28) / n);
    block_rows = std::min<long long>(block_rows, n);
    if (block_rows * n >= (1LL << 31)) MCBB_FAIL("k_dist: block_rows * n must stay below 2^31");
    double* blk = (double*)scratch_get(66, sizeof(double) * (size_t)n * (size_t)block_rows);
    double* sorted = (double*)scratch_get(67, sizeof(double) * (size_t)n * (size_t)block_rows);
    long long* off = (long long*)scratch_get(46, sizeof(long long) * (size_t)(block_rows + 1));
    if (!blk || !sorted || !off) MCBB_FAIL("k_dist: device scratch allocation failed");
    k_row_offsets_ld<<<nblk(block_rows + 1), 256, 0, st>>>(off, block_rows, n);
    MCBB_LAUNCHED();
    size_t tb = 0;
    cub::DeviceSegmentedSort::SortKeys(nullptr, tb, blk, sorted, (int)(n * block_rows), (int)block_rows, off, off + 1, st);
    void* tmp = scratch_get(47, tb);
    if (!tmp) MCBB_FAIL("k_dist: device scratch allocation failed");
    for (long long r0 = 0; r0 < n; r0 += block_rows) {
        const long long r1 = std::min(n, r0 + block_rows), m = r1 - r0;
        // rows r0..r1 of D, each contiguous (column-major block of the symmetric matrix)
        if (int rc = distance_rows_dev(X, n, n_groups, group_len, group_weight, r0, r1, blk, st)) return rc;
        size_t tb2 = tb;
        MCBB_CUDA_OK(cub::DeviceSegmentedSort::SortKeys(tmp, tb2, blk, sorted, (int)(n * m), (int)m, off, off + 1, st));
        g_launch_count.fetch_add(1);
        k_knn_pick<<<nblk(m), 256, 0, st>>>(sorted, n, k, K, kth ? kth + r0 : nullptr, cum ? cum + r0 : nullptr, m);
        MCBB_LAUNCHED();
    }
    return 0;
}

}  // namespace mcbb
