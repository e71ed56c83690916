// multi.cu — several B200s driven from ONE host process behind the C-ABI (SURVEY §8(b) "Threading": multi-GPU is internal,
// one host thread per device + an NCCL communicator created by the library).  A Julia host has no torch.distributed:
// `ccall(:mcbb_multi_init, ...)` once, then the *_multi entry points take the same HOST buffers as their single-device
// twins and shard the work themselves.
//
//   integrate + featurize   runs are independent: contiguous blocks of run indices per device, every device copies its
//                           columns of ic / par in, solves, and copies its columns of the outputs back — no collective
//   DBSCAN                  every device holds the feature matrix, the pair space is sharded by interleaved row blocks
//                           (cluster.cu: dbscan_features_multi_rank); counts / parents / border labels are exchanged
//                           with NCCL all-reduce / all-gather over NVLink
//   solve + cluster         the features never leave the devices: NCCL all-gather of the per-device feature blocks, a
//                           stable device sort by the first parameter (what MCBB.solve's sort! does on the host,
//                           setup_mc_prob.jl:520-536), sharded DBSCAN on the sorted runs
//
// NCCL is loaded with dlopen at mcbb_multi_init (libnccl.so.2): the single-device library keeps no link-time dependency
// on it, and a process that already loaded another NCCL (torch) shares that copy.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "multi.cuh"

namespace mcbb {

// ---- the few NCCL entry points used, bound at run time (prototypes as in nccl.h 2.x) --------------------------------
typedef struct ncclComm* ncclComm_t;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt32 = 2, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;

struct MultiState {
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;
    int nccl_version = 0;
};
static MultiState g_multi;
static std::mutex g_multi_mu;

static int load_nccl() {
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
    }
    if (!h) MCBB_FAIL(std::string("mcbb_multi_init: cannot load libnccl.so.2 (") + dlerror() + ")");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))dlsym(h, "ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    g_nccl.GetVersion = (decltype(g_nccl.GetVersion))dlsym(h, "ncclGetVersion");
    if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.GetErrorString) {
        dlclose(h);
        MCBB_FAIL("mcbb_multi_init: libnccl.so.2 lacks a required symbol");
    }
    g_nccl.handle = h;
    return 0;
}

#define MCBB_NCCL_OK(expr)                                                                                   \
    do {                                                                                                     \
        ncclResult_t _r = (expr);                                                                            \
        if (_r != ncclSuccess) {                                                                             \
            ::mcbb::set_error(std::string(#expr) + ": " + g_nccl.GetErrorString(_r) + " (" + __FILE__ + ":" + \
                              std::to_string(__LINE__) + ")");                                               \
            return 3;                                                                                        \
        }                                                                                                    \
    } while (0)

struct RankCtx {
    int rank;
    ncclComm_t comm;
};
static int comm_allreduce_i32(void* ctx, int* buf, long long count, int op, cudaStream_t st) {
    RankCtx* c = (RankCtx*)ctx;
    MCBB_NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t)count, ncclInt32, op == 1 ? ncclMin : ncclSum, c->comm, st));
    return 0;
}
static int comm_allgather_i32(void* ctx, const int* send, int* recv, long long count, cudaStream_t st) {
    RankCtx* c = (RankCtx*)ctx;
    MCBB_NCCL_OK(g_nccl.AllGather(send, recv, (size_t)count, ncclInt32, c->comm, st));
    return 0;
}

// runs fn(rank) on one host thread per device (bound to that device); returns the first non-zero status, with its text
static int run_on_devices(const std::function<int(int)>& fn) {
    const int G = (int)g_multi.devices.size();
    std::vector<int> rc(G, 0);
    std::vector<std::string> msg(G);
    std::vector<std::thread> th;
    for (int r = 0; r < G; r++)
        th.emplace_back([&, r]() {
            if (cudaSetDevice(g_multi.devices[r]) != cudaSuccess) {
                rc[r] = 2;
                msg[r] = "cudaSetDevice failed";
                return;
            }
            rc[r] = fn(r);
            if (rc[r]) {
                char buf[1024];
                mcbb_last_error(buf, sizeof(buf));  // thread-local text of THIS worker
                msg[r] = buf;
            }
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < G; r++)
        if (rc[r]) {
            set_error("device " + std::to_string(g_multi.devices[r]) + ": " + msg[r]);
            return rc[r];
        }
    return 0;
}

static void shard(long long n, int r, int G, long long* lo, long long* hi) {
    const long long base = n / G, rem = n % G;
    *lo = r * base + std::min<long long>(r, rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}

// X[f][lo + k] (ld = n) <-> block[f][k] (ld = m): strided 2-D copies between a column slab of a host matrix and a
// compact device block
static cudaError_t copy_cols_h2d(double* dst, const double* src, long long n, long long lo, long long m, int rows,
                                 cudaStream_t st) {
    return cudaMemcpy2DAsync(dst, sizeof(double) * m, src + lo, sizeof(double) * n, sizeof(double) * m, rows,
                             cudaMemcpyHostToDevice, st);
}
static cudaError_t copy_cols_d2h(double* dst, const double* src, long long n, long long lo, long long m, int rows,
                                 cudaStream_t st) {
    return cudaMemcpy2DAsync(dst + lo, sizeof(double) * n, src, sizeof(double) * m, sizeof(double) * m, rows,
                             cudaMemcpyDeviceToHost, st);
}

__global__ void k_unpad_blocks(const double* padded, long long mmax, const long long* lo, int G, int F, long long n,
                               double* X) {
    // padded [G][F][mmax] -> X [F][n], rank g holding columns [lo[g], lo[g+1])
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (i >= n || f >= F) return;
    int g = 0;
    while (g + 1 < G && i >= lo[g + 1]) g++;
    X[(long long)f * n + i] = padded[((long long)g * F + f) * mmax + (i - lo[g])];
}
__global__ void k_iota_i32_multi(int* p, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}
__global__ void k_gather_cols_multi(const double* X, long long ld, const int* idx, long long m, int F, double* Xout) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (k < m && f < F) Xout[(long long)f * m + k] = X[(long long)f * ld + idx[k]];
}
__global__ void k_nan_to_big(double* X, long long cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt && X[i] != X[i]) X[i] = 1e30;
}

}  // namespace mcbb

using namespace mcbb;

extern "C" {

int mcbb_multi_init(int32_t n_devices, const int32_t* devices) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    if (!g_multi.devices.empty()) MCBB_FAIL("mcbb_multi_init: already initialised (call mcbb_multi_finalize first)");
    int avail = 0;
    MCBB_CUDA_OK(cudaGetDeviceCount(&avail));
    if (n_devices < 1 || n_devices > avail || n_devices > 16)
        MCBB_FAIL("mcbb_multi_init: " + std::to_string(n_devices) + " devices requested, " + std::to_string(avail) + " visible");
    if (int rc = load_nccl()) return rc;
    std::vector<int> devs(n_devices);
    for (int r = 0; r < n_devices; r++) devs[r] = devices ? devices[r] : r;
    for (int r = 0; r < n_devices; r++) {
        if (devs[r] < 0 || devs[r] >= avail) MCBB_FAIL("mcbb_multi_init: device index out of range");
        if (int rc = mcbb_init(devs[r])) return rc;  // compute-capability check, context creation
    }
    std::vector<ncclComm_t> comms(n_devices);
    MCBB_NCCL_OK(g_nccl.CommInitAll(comms.data(), n_devices, devs.data()));
    std::vector<cudaStream_t> streams(n_devices);
    for (int r = 0; r < n_devices; r++) {
        MCBB_CUDA_OK(cudaSetDevice(devs[r]));
        MCBB_CUDA_OK(cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking));
    }
    MCBB_CUDA_OK(cudaSetDevice(devs[0]));
    if (g_nccl.GetVersion) g_nccl.GetVersion(&g_multi.nccl_version);
    g_multi.devices = devs;
    g_multi.comms = comms;
    g_multi.streams = streams;
    return 0;
}

int mcbb_multi_finalize(void) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    for (size_t r = 0; r < g_multi.devices.size(); r++) {
        cudaSetDevice(g_multi.devices[r]);
        cudaStreamSynchronize(g_multi.streams[r]);
        g_nccl.CommDestroy(g_multi.comms[r]);
        cudaStreamDestroy(g_multi.streams[r]);
    }
    if (!g_multi.devices.empty()) cudaSetDevice(g_multi.devices[0]);
    g_multi = MultiState();
    return 0;
}

int mcbb_multi_info(int32_t* n_devices, int32_t* nccl_version) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    if (n_devices) *n_devices = (int32_t)g_multi.devices.size();
    if (nccl_version) *nccl_version = g_multi.nccl_version;
    return 0;
}

int mcbb_solve_features_multi(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                              int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                              double* feat_out, double* matrix_out, double* global_out, int32_t* retcode_out,
                              int64_t* nsteps_out) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    if (g_multi.devices.empty()) MCBB_FAIL("mcbb_solve_features_multi: call mcbb_multi_init first");
    if (!sys || !opts || !feat) MCBB_FAIL("solve: NULL descriptor");
    if (n_mc < 1) MCBB_FAIL("solve: N_mc must be positive");
    if (!ic || !t_save || !feat_out) MCBB_FAIL("solve: NULL ic / t_save / feat_out");
    if (sys->n_par > 0 && !par) MCBB_FAIL("solve: NULL par with n_par > 0");
    const int G = (int)std::min<long long>((long long)g_multi.devices.size(), n_mc);
    const bool cplx = sys->system == MCBB_SYS_STUART_LANDAU;
    const int nf = feat->n_filter > 0 ? feat->n_filter : (cplx ? sys->n_dim / 2 : sys->n_dim);
    const int rows_feat = nf * feat->n_meas;
    const int nbins = feat->corr_nbins > 0 ? feat->corr_nbins : 30;
    const bool has_mat = feat->matrix_meas != MCBB_MAT_NONE;
    if (has_mat && !matrix_out) MCBB_FAIL("solve: a matrix measure is requested but matrix_out is NULL");
    if (feat->n_global > 0 && !global_out) MCBB_FAIL("solve: global measures requested but global_out is NULL");
    return run_on_devices([&](int r) -> int {
        if (r >= G) return 0;
        long long lo, hi;
        shard(n_mc, r, G, &lo, &hi);
        const long long m = hi - lo;
        cudaStream_t st = g_multi.streams[r];
        ApiGuard guard(st);  // this device's working state is ours until the results are back on the host
        double* d_ic = (double*)scratch_get(30, sizeof(double) * (size_t)m * sys->n_dim);
        double* d_par = (double*)scratch_get(31, sizeof(double) * (size_t)m * (sys->n_par > 0 ? sys->n_par : 1));
        double* d_ts = (double*)scratch_get(32, sizeof(double) * (size_t)n_t);
        double* d_feat = (double*)scratch_get(33, sizeof(double) * (size_t)m * rows_feat);
        double* d_glob = (double*)scratch_get(82, sizeof(double) * (size_t)m * (feat->n_global > 0 ? feat->n_global : 1));
        double* d_mat = (double*)scratch_get(37, sizeof(double) * (size_t)m * nbins);
        int* d_ret = (int*)scratch_get(83, sizeof(int) * (size_t)m);
        long long* d_ns = (long long*)scratch_get(36, sizeof(long long) * 2 * (size_t)m);
        if (!d_ic || !d_par || !d_ts || !d_feat || !d_glob || !d_mat || !d_ret || !d_ns) MCBB_FAIL("solve: device allocation failed");
        MCBB_CUDA_OK(copy_cols_h2d(d_ic, ic, n_mc, lo, m, sys->n_dim, st));
        if (sys->n_par > 0) MCBB_CUDA_OK(copy_cols_h2d(d_par, par, n_mc, lo, m, sys->n_par, st));
        MCBB_CUDA_OK(cudaMemcpyAsync(d_ts, t_save, sizeof(double) * (size_t)n_t, cudaMemcpyHostToDevice, st));
        if (has_mat) MCBB_CUDA_OK(cudaMemsetAsync(d_mat, 0xFF, sizeof(double) * (size_t)m * nbins, st));
        if (int rc = mcbb_solve_features_dev(sys, opts, feat, m, d_ic, d_par, d_ts, n_t, d_feat, has_mat ? d_mat : nullptr,
                                             feat->n_global > 0 ? d_glob : nullptr, d_ret, (int64_t*)d_ns, (void*)st))
            return rc;
        MCBB_CUDA_OK(copy_cols_d2h(feat_out, d_feat, n_mc, lo, m, rows_feat, st));
        if (has_mat) MCBB_CUDA_OK(copy_cols_d2h(matrix_out, d_mat, n_mc, lo, m, nbins, st));
        if (feat->n_global > 0) MCBB_CUDA_OK(copy_cols_d2h(global_out, d_glob, n_mc, lo, m, feat->n_global, st));
        if (retcode_out)
            MCBB_CUDA_OK(cudaMemcpyAsync(retcode_out + lo, d_ret, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, st));
        if (nsteps_out) {  // [2][n_mc]
            MCBB_CUDA_OK(cudaMemcpy2DAsync(nsteps_out + lo, sizeof(long long) * n_mc, d_ns, sizeof(long long) * m,
                                           sizeof(long long) * m, 2, cudaMemcpyDeviceToHost, st));
        }
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    });
}

int mcbb_dbscan_features_multi(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, double eps, int32_t minpts, int64_t* assignments,
                               int64_t* seeds, int64_t* counts, int64_t* n_clusters) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    if (g_multi.devices.empty()) MCBB_FAIL("mcbb_dbscan_features_multi: call mcbb_multi_init first");
    if (!X || !assignments || !seeds || !counts || !n_clusters || !group_len || !group_weight) MCBB_FAIL("dbscan: NULL argument");
    if (n < 2) MCBB_FAIL("At least two data points are required.");
    size_t F = 0;
    for (int g = 0; g < n_groups; g++) F += group_len[g] > 0 ? group_len[g] : 0;
    const int G = (int)g_multi.devices.size();
    std::vector<long long> ks(G, 0);
    int rc = run_on_devices([&](int r) -> int {
        cudaStream_t st = g_multi.streams[r];
        ApiGuard guard(st);
        double* d_X = (double*)scratch_get(40, sizeof(double) * F * (size_t)n);
        long long* d_a = (long long*)scratch_get(42, sizeof(long long) * (size_t)n);
        long long* d_s = (long long*)scratch_get(43, sizeof(long long) * (size_t)n);
        long long* d_c = (long long*)scratch_get(44, sizeof(long long) * (size_t)n);
        if (!d_X || !d_a || !d_s || !d_c) MCBB_FAIL("dbscan: device allocation failed");
        MCBB_CUDA_OK(cudaMemcpyAsync(d_X, X, sizeof(double) * F * (size_t)n, cudaMemcpyHostToDevice, st));
        RankCtx ctx{r, g_multi.comms[r]};
        MultiComm mc{r, G, comm_allreduce_i32, comm_allgather_i32, &ctx};
        if (int rc2 = dbscan_features_multi_rank(d_X, n, n_groups, group_len, group_weight, eps, minpts, mc, d_a, d_s, d_c,
                                                 &ks[r], st))
            return rc2;
        if (r == 0) {
            MCBB_CUDA_OK(cudaMemcpyAsync(assignments, d_a, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, st));
            MCBB_CUDA_OK(cudaMemcpyAsync(seeds, d_s, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, st));
            MCBB_CUDA_OK(cudaMemcpyAsync(counts, d_c, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, st));
        }
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    });
    if (rc) return rc;
    *n_clusters = ks[0];
    return 0;
}

int mcbb_solve_cluster_multi(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                             int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                             const double* weights, double eps, int32_t minpts, double* feat_out, int32_t* retcode_out,
                             int64_t* nsteps_out, int64_t* order_out, int64_t* assignments, int64_t* seeds, int64_t* counts,
                             int64_t* n_clusters) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    if (g_multi.devices.empty()) MCBB_FAIL("mcbb_solve_cluster_multi: call mcbb_multi_init first");
    if (!sys || !opts || !feat || !weights) MCBB_FAIL("solve_cluster: NULL descriptor");
    if (n_mc < 2) MCBB_FAIL("At least two data points are required.");
    if (!ic || !t_save || !assignments || !seeds || !counts || !n_clusters) MCBB_FAIL("solve_cluster: NULL argument");
    if (sys->n_par < 1 || !par) MCBB_FAIL("solve_cluster: the runs are sorted by the first varied parameter; n_par must be >= 1");
    if (feat->matrix_meas != MCBB_MAT_NONE || feat->n_global > 0)
        MCBB_FAIL("solve_cluster: matrix / global measures are not supported by the one-call pipeline; use the separate entries");
    const int G = (int)g_multi.devices.size();
    if (n_mc < G) MCBB_FAIL("solve_cluster: fewer runs than devices");
    const bool cplx = sys->system == MCBB_SYS_STUART_LANDAU;
    const int nf = feat->n_filter > 0 ? feat->n_filter : (cplx ? sys->n_dim / 2 : sys->n_dim);
    const int rows_feat = nf * feat->n_meas;
    const int F = rows_feat + sys->n_par;  // feature matrix of distance_matrix: measures, then the parameter columns
    std::vector<int32_t> glen;
    std::vector<double> gw;
    for (int m = 0; m < feat->n_meas; m++) {
        glen.push_back(nf);
        gw.push_back(weights[m]);
    }
    for (int p = 0; p < sys->n_par; p++) {
        glen.push_back(1);
        gw.push_back(weights[feat->n_meas + p]);
    }
    std::vector<long long> los(G + 1, 0);
    long long mmax = 0;
    for (int r = 0; r < G; r++) {
        long long lo, hi;
        shard(n_mc, r, G, &lo, &hi);
        los[r] = lo;
        los[r + 1] = hi;
        mmax = std::max(mmax, hi - lo);
    }
    std::vector<long long> ks(G, 0);
    int rc = run_on_devices([&](int r) -> int {
        cudaStream_t st = g_multi.streams[r];
        ApiGuard guard(st);
        const long long lo = los[r], m = los[r + 1] - los[r];
        // ---- integrate + featurize the owned block; local feature block [F][mmax] (padded for the all-gather) -----
        double* d_ic = (double*)scratch_get(30, sizeof(double) * (size_t)m * sys->n_dim);
        double* d_ts = (double*)scratch_get(32, sizeof(double) * (size_t)n_t);
        double* d_blk = (double*)scratch_get(84, sizeof(double) * (size_t)F * (size_t)mmax);
        double* d_feat = (double*)scratch_get(33, sizeof(double) * (size_t)m * rows_feat);
        double* d_par = (double*)scratch_get(31, sizeof(double) * (size_t)m * sys->n_par);
        int* d_ret = (int*)scratch_get(83, sizeof(int) * (size_t)m);
        long long* d_ns = (long long*)scratch_get(36, sizeof(long long) * 2 * (size_t)m);
        double* d_all = (double*)scratch_get(85, sizeof(double) * (size_t)G * F * (size_t)mmax);
        double* d_X = (double*)scratch_get(40, sizeof(double) * (size_t)F * (size_t)n_mc);
        double* d_Xs = (double*)scratch_get(86, sizeof(double) * (size_t)F * (size_t)n_mc);
        long long* d_lo = (long long*)scratch_get(87, sizeof(long long) * (G + 1));
        int* d_idx = (int*)scratch_get(88, sizeof(int) * (size_t)n_mc);
        int* d_ord = (int*)scratch_get(89, sizeof(int) * (size_t)n_mc);
        double* d_key = (double*)scratch_get(90, sizeof(double) * (size_t)n_mc);
        long long* d_a = (long long*)scratch_get(42, sizeof(long long) * (size_t)n_mc);
        long long* d_s = (long long*)scratch_get(43, sizeof(long long) * (size_t)n_mc);
        long long* d_c = (long long*)scratch_get(44, sizeof(long long) * (size_t)n_mc);
        if (!d_ic || !d_ts || !d_blk || !d_feat || !d_par || !d_ret || !d_ns || !d_all || !d_X || !d_Xs || !d_lo || !d_idx ||
            !d_ord || !d_key || !d_a || !d_s || !d_c)
            MCBB_FAIL("solve_cluster: device allocation failed");
        MCBB_CUDA_OK(copy_cols_h2d(d_ic, ic, n_mc, lo, m, sys->n_dim, st));
        MCBB_CUDA_OK(copy_cols_h2d(d_par, par, n_mc, lo, m, sys->n_par, st));
        MCBB_CUDA_OK(cudaMemcpyAsync(d_ts, t_save, sizeof(double) * (size_t)n_t, cudaMemcpyHostToDevice, st));
        MCBB_CUDA_OK(cudaMemcpyAsync(d_lo, los.data(), sizeof(long long) * (G + 1), cudaMemcpyHostToDevice, st));
        if (int rc2 = mcbb_solve_features_dev(sys, opts, feat, m, d_ic, d_par, d_ts, n_t, d_feat, nullptr, nullptr, d_ret,
                                              (int64_t*)d_ns, (void*)st))
            return rc2;
        if (feat_out) MCBB_CUDA_OK(copy_cols_d2h(feat_out, d_feat, n_mc, lo, m, rows_feat, st));
        if (retcode_out) MCBB_CUDA_OK(cudaMemcpyAsync(retcode_out + lo, d_ret, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, st));
        if (nsteps_out)
            MCBB_CUDA_OK(cudaMemcpy2DAsync(nsteps_out + lo, sizeof(long long) * n_mc, d_ns, sizeof(long long) * m,
                                           sizeof(long long) * m, 2, cudaMemcpyDeviceToHost, st));
        MCBB_CUDA_OK(cudaMemcpy2DAsync(d_blk, sizeof(double) * mmax, d_feat, sizeof(double) * m, sizeof(double) * m, rows_feat,
                                       cudaMemcpyDeviceToDevice, st));
        MCBB_CUDA_OK(cudaMemcpy2DAsync(d_blk + (size_t)rows_feat * mmax, sizeof(double) * mmax, d_par, sizeof(double) * m,
                                       sizeof(double) * m, sys->n_par, cudaMemcpyDeviceToDevice, st));
        // ---- all-gather the feature blocks over NVLink, unpad, failed runs (NaN features) pushed far away ------------
        MCBB_NCCL_OK(g_nccl.AllGather(d_blk, d_all, (size_t)F * (size_t)mmax, ncclFloat64, g_multi.comms[r], st));
        const unsigned nb = (unsigned)((n_mc + 255) / 256);
        k_unpad_blocks<<<dim3(nb, F), 256, 0, st>>>(d_all, mmax, d_lo, G, F, n_mc, d_X);
        MCBB_LAUNCHED();
        k_nan_to_big<<<(unsigned)(((long long)F * n_mc + 255) / 256), 256, 0, st>>>(d_X, (long long)F * n_mc);
        MCBB_LAUNCHED();
        // ---- stable sort by the first parameter (sort!(sol, prob)): radix sort of (par, index) -----------------------
        k_iota_i32_multi<<<nb, 256, 0, st>>>(d_idx, n_mc);
        MCBB_LAUNCHED();
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, d_X + (size_t)rows_feat * n_mc, d_key, d_idx, d_ord, (int)n_mc, 0, 64, st);
        void* tmp = scratch_get(91, tb);
        if (!tmp) MCBB_FAIL("solve_cluster: device allocation failed");
        MCBB_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tb, d_X + (size_t)rows_feat * n_mc, d_key, d_idx, d_ord, (int)n_mc, 0,
                                                     64, st));
        g_launch_count.fetch_add(1);
        k_gather_cols_multi<<<dim3(nb, F), 256, 0, st>>>(d_X, n_mc, d_ord, n_mc, F, d_Xs);
        MCBB_LAUNCHED();
        // ---- sharded DBSCAN on the sorted runs ------------------------------------------------------------------------
        RankCtx ctx{r, g_multi.comms[r]};
        MultiComm mc{r, G, comm_allreduce_i32, comm_allgather_i32, &ctx};
        if (int rc2 = dbscan_features_multi_rank(d_Xs, n_mc, (int)glen.size(), glen.data(), gw.data(), eps, minpts, mc, d_a, d_s,
                                                 d_c, &ks[r], st))
            return rc2;
        if (r == 0) {
            MCBB_CUDA_OK(cudaMemcpyAsync(assignments, d_a, sizeof(long long) * (size_t)n_mc, cudaMemcpyDeviceToHost, st));
            MCBB_CUDA_OK(cudaMemcpyAsync(seeds, d_s, sizeof(long long) * (size_t)n_mc, cudaMemcpyDeviceToHost, st));
            MCBB_CUDA_OK(cudaMemcpyAsync(counts, d_c, sizeof(long long) * (size_t)n_mc, cudaMemcpyDeviceToHost, st));
            if (order_out) {
                std::vector<int> h(n_mc);
                MCBB_CUDA_OK(cudaMemcpyAsync(h.data(), d_ord, sizeof(int) * (size_t)n_mc, cudaMemcpyDeviceToHost, st));
                MCBB_CUDA_OK(cudaStreamSynchronize(st));
                for (long long i = 0; i < n_mc; i++) order_out[i] = h[i];
            }
        }
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    });
    if (rc) return rc;
    *n_clusters = ks[0];
    return 0;
}

}  // extern "C"
