// hist.cu — histogram-mode features on the device: what distance_matrix(sol, prob, ...; histograms = true) builds before
// it compares two runs (src/eval_clustering.jl:187-211).
//
//   * _compute_hist_edges (eval_clustering.jl:607-642) needs, over ALL runs and dimensions of one measure, the minimum, the
//     maximum and the inter-quartile range: `order_stats_dev` sorts a copy of the measure's slab (cub radix sort) and hands
//     back the order statistics the host asks for (the two neighbours of each type-7 quantile) — the edges themselves are a
//     Julia range, built on the host from those few numbers.
//   * _compute_hist_weights (:587-598) + ecdf_hist (:759-762): per run, the Float32 histogram of the N_dim values over the
//     common edges (closed = :left, out-of-range values dropped), accumulated to an ECDF in Float32 and normalised by its
//     last entry.  `hist_rows_dev`: one thread per run, counts in shared memory ([bin][thread]), rows written feature-major
//     ([bin][run]) — the layout the pair tiles and the on-the-fly DBSCAN read.
//   * normalize(sol) (setup_mc_prob.jl:635-660) and sort!(sol, prob) (:520-536) are folded in: an optional (x - lo) / range
//     applied to every value as it is read, an optional permutation of the runs.
//
// Arithmetic matches the host mirror operation by operation (no FMA contraction): bin = searchsortedlast on the range,
// i.e. round((x - first) / step + 1) clamped to 1 .. n_edges, minus one when x lies below that edge.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mcbb {

constexpr int HR_THREADS = 128;

__global__ void __launch_bounds__(HR_THREADS) k_hist_rows(const double* __restrict__ values, const long long n_mc, const int n_dim,
                                                          const long long* __restrict__ order, const int normalize,
                                                          const double lo, const double range, const double* __restrict__ edges,
                                                          const int n_edges, const int use_ecdf, double* __restrict__ out) {
    extern __shared__ unsigned char hr_smem[];
    double* s_edges = reinterpret_cast<double*>(hr_smem);                       // [n_edges]
    unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_edges + n_edges);  // [nb][HR_THREADS]
    const int nb = n_edges - 1;
    for (int i = threadIdx.x; i < n_edges; i += HR_THREADS) s_edges[i] = edges[i];
    for (int b = 0; b < nb; b++) s_cnt[b * HR_THREADS + threadIdx.x] = 0;
    __syncthreads();
    const long long r = (long long)blockIdx.x * HR_THREADS + threadIdx.x;
    if (r >= n_mc) return;
    const long long src = order ? order[r] : r;
    const double first = s_edges[0], step = __dsub_rn(s_edges[1], s_edges[0]);
    const double ne = (double)n_edges;
    for (int d0 = 0; d0 < n_dim; d0 += 8) {
        double x[8];
#pragma unroll
        for (int q = 0; q < 8; q++) x[q] = (d0 + q < n_dim) ? __ldg(values + (long long)(d0 + q) * n_mc + src) : nan("");
#pragma unroll
        for (int q = 0; q < 8; q++) {
            double v = x[q];
            if (normalize) v = __ddiv_rn(__dsub_rn(v, lo), range);
            if (v != v) continue;
            double f = __dadd_rn(__ddiv_rn(__dsub_rn(v, first), step), 1.0);
            f = fmin(fmax(f, 1.0), ne);
            int k = (int)rint(f);
            if (v < s_edges[k - 1]) k--;
            if (k >= 1 && k <= nb) s_cnt[(k - 1) * HR_THREADS + threadIdx.x] += 1;
        }
    }
    if (!use_ecdf) {
        for (int b = 0; b < nb; b++) out[(long long)b * n_mc + r] = (double)(float)s_cnt[b * HR_THREADS + threadIdx.x];
        return;
    }
    float tot = 0.f;
    for (int b = 0; b < nb; b++) tot = __fadd_rn(tot, (float)s_cnt[b * HR_THREADS + threadIdx.x]);
    float cum = 0.f;
    for (int b = 0; b < nb; b++) {
        cum = __fadd_rn(cum, (float)s_cnt[b * HR_THREADS + threadIdx.x]);
        out[(long long)b * n_mc + r] = (double)__fdiv_rn(cum, tot);  // 0 / 0 = NaN for a run without a value in range
    }
}

int hist_rows_dev(const double* values, long long n_mc, int n_dim, const long long* order, int normalize, double lo,
                  double range, const double* edges_host, int n_edges, int use_ecdf, double* out, cudaStream_t st) {
    if (n_mc <= 0) return 0;
    const size_t smem = sizeof(double) * (size_t)n_edges + sizeof(unsigned short) * (size_t)(n_edges - 1) * HR_THREADS;
    if (smem > 200 * 1024) MCBB_FAIL("hist_rows: too many histogram edges");
    if (n_dim > 65535) MCBB_FAIL("hist_rows: more than 65535 dimensions per run");
    double* d_edges = (double*)scratch_get(95, sizeof(double) * (size_t)n_edges);
    if (!d_edges) MCBB_FAIL("hist_rows: device allocation failed");
    MCBB_CUDA_OK(cudaMemcpyAsync(d_edges, edges_host, sizeof(double) * (size_t)n_edges, cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));  // edges_host belongs to the caller
    MCBB_CUDA_OK(cudaFuncSetAttribute(k_hist_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (n_mc + HR_THREADS - 1) / HR_THREADS;
    k_hist_rows<<<(unsigned)grid, HR_THREADS, smem, st>>>(values, n_mc, n_dim, order, normalize, lo, range, d_edges, n_edges,
                                                          use_ecdf, out);
    MCBB_LAUNCHED();
    return 0;
}

__global__ void k_count_nan(const double* __restrict__ v, const long long n, unsigned long long* out) {
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += (v[i] != v[i]);
    for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void k_pick_ranks(const double* __restrict__ sorted, const long long* __restrict__ ranks, const int n_ranks, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ranks) out[i] = sorted[ranks[i]];
}

// out_host[i] = the ranks_host[i]-th smallest value (0-based) of values[0 .. count); NaNs are counted, not ranked
// (radix order puts them at the ends: the caller decides what a measure with NaNs means).
int order_stats_dev(const double* values, long long count, int n_ranks, const long long* ranks_host, double* out_host,
                    long long* n_nan_host, cudaStream_t st) {
    if (count <= 0 || count > 0x7fffffffLL) MCBB_FAIL("order_stats: element count out of range");
    for (int i = 0; i < n_ranks; i++)
        if (ranks_host[i] < 0 || ranks_host[i] >= count) MCBB_FAIL("order_stats: rank out of range");
    double* d_sorted = (double*)scratch_get(96, sizeof(double) * (size_t)count);
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, values, d_sorted, (int)count, 0, 64, st);
    void* tmp = scratch_get(97, tb);
    long long* d_ranks = (long long*)scratch_get(98, sizeof(long long) * (size_t)(n_ranks + 1));
    double* d_out = (double*)scratch_get(99, sizeof(double) * (size_t)(n_ranks + 2));
    if (!d_sorted || !tmp || !d_ranks || !d_out) MCBB_FAIL("order_stats: device allocation failed");
    unsigned long long* d_nan = reinterpret_cast<unsigned long long*>(d_out + n_ranks);
    MCBB_CUDA_OK(cudaMemsetAsync(d_nan, 0, sizeof(unsigned long long), st));
    k_count_nan<<<592, 256, 0, st>>>(values, count, d_nan);
    MCBB_LAUNCHED();
    MCBB_CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp, tb, values, d_sorted, (int)count, 0, 64, st));
    g_launch_count.fetch_add(1);
    if (n_ranks > 0) {
        MCBB_CUDA_OK(cudaMemcpyAsync(d_ranks, ranks_host, sizeof(long long) * (size_t)n_ranks, cudaMemcpyHostToDevice, st));
        k_pick_ranks<<<(n_ranks + 127) / 128, 128, 0, st>>>(d_sorted, d_ranks, n_ranks, d_out);
        MCBB_LAUNCHED();
        MCBB_CUDA_OK(cudaMemcpyAsync(out_host, d_out, sizeof(double) * (size_t)n_ranks, cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_nan = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(&h_nan, d_nan, sizeof(h_nan), cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    if (n_nan_host) *n_nan_host = (long long)h_nan;
    return 0;
}

}  // namespace mcbb
