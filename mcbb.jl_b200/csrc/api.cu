// api.cu — the extern "C" surface declared in include/mcbb_b200.h.
#include <math.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mcbb {

static thread_local std::string t_last_error;
std::atomic<long long> g_launch_count{0};

void set_error(const std::string& msg) { t_last_error = msg; }

// ---- device scratch cache ------------------------------------------------------------------------------------
struct ScratchSlot {
    void* p = nullptr;
    size_t cap = 0;
};
constexpr int MAX_DEV = 16;
constexpr int N_SCRATCH = 112;
static ScratchSlot g_scratch[MAX_DEV][N_SCRATCH];
static std::mutex g_scratch_mu;

// Working state (scratch buffers, the trajectory queue counter, constant-bank tables) is kept PER DEVICE, not per
// stream.  ApiGuard makes that safe: every compute entry point holds the device's lock for the duration of the call
// (host threads driving different devices run concurrently, two threads on one device take turns), and a call that
// arrives on another stream than the previous one of that device first waits for the previous stream to drain, so a
// kernel still running there never sees its tables overwritten.  One stream per device costs nothing.
struct DevState {
    std::recursive_mutex mu;
    cudaStream_t last = nullptr;
    bool used = false;
};
static DevState g_dev[MAX_DEV];

static int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < MAX_DEV) ? dev : 0;
}

ApiGuard::ApiGuard(cudaStream_t st) : dev_(current_device_slot()) {
    DevState& d = g_dev[dev_];
    d.mu.lock();
    if (d.used && d.last != st) cudaStreamSynchronize(d.last);
    d.last = st;
    d.used = true;
}
ApiGuard::~ApiGuard() { g_dev[dev_].mu.unlock(); }

size_t scratch_size(int slot) {
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    return g_scratch[current_device_slot()][slot].cap;
}

void* scratch_get(int slot, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    if (bytes == 0) bytes = 16;
    ScratchSlot& s = g_scratch[current_device_slot()][slot];
    if (s.p && s.cap >= bytes) return s.p;
    if (s.p) {
        cudaDeviceSynchronize();
        cudaFree(s.p);
        s.p = nullptr;
        s.cap = 0;
    }
    size_t want = bytes + bytes / 8;
    if (cudaMalloc(&s.p, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        if (cudaMalloc(&s.p, want) != cudaSuccess) {
            cudaGetLastError();
            s.p = nullptr;
            return nullptr;
        }
    }
    s.cap = want;
    return s.p;
}

void scratch_free_all() {
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    int cur = 0, ndev = 0;
    cudaGetDevice(&cur);
    cudaGetDeviceCount(&ndev);
    for (int d = 0; d < MAX_DEV && d < ndev; d++) {
        bool any = false;
        for (auto& s : g_scratch[d]) any |= (s.p != nullptr);
        if (!any) continue;
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        for (auto& s : g_scratch[d]) {
            if (s.p) cudaFree(s.p);
            s = ScratchSlot();
        }
    }
    cudaSetDevice(cur);
    cudaGetLastError();
}

static std::mutex g_tune_mu;
static std::map<std::string, long long> g_tune;
long long tune_get(const char* key, long long dflt) {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    auto it = g_tune.find(key);
    return it == g_tune.end() ? dflt : it->second;
}

// defined in solve.cu / cluster.cu
int table_lengths(const mcbb_system_desc* s, int len[5]);
int resolve_feat(const mcbb_system_desc* sys, const mcbb_feat_opts* f, FeatDev* out, std::vector<int>* filter_host);
int solve_features_launch(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F,
                          const SolveIO& io, cudaStream_t st);
int distance_matrix_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                        const double* group_weight, double* D, cudaStream_t st);
int dbscan_features_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                        const double* group_weight, double eps, int minpts, long long* assignments,
                        long long* seeds, long long* counts, long long* n_clusters_host, cudaStream_t st);
int dbscan_dense_dev(const double* D, long long n, double eps, int minpts, long long* assignments,
                     long long* seeds, long long* counts, long long* n_clusters_host, cudaStream_t st);

int rows_count_dev(const double* X, long long n, int n_groups, const int32_t* group_len, const double* group_weight,
                   double eps, int minpts, long long row0, long long row1, int* count, int* nbr, int* nfill,
                   cudaStream_t st);
int rows_union_dev(const double* Xc, long long n_core, int n_groups, const int32_t* group_len,
                   const double* group_weight, double eps, long long row0, long long row1, int* parent, int reset,
                   cudaStream_t st);
int merge_parents_dev(int* parent, const int* other, long long n_core, cudaStream_t st);
int rows_border_dev(const double* Xn, long long n_non, long long row0, long long row1, const double* Xc,
                    long long n_core, const int* core_seed, int n_groups, const int32_t* group_len,
                    const double* group_weight, double eps, int* best, cudaStream_t st);
int labels_dev(const int* root_label, long long n, long long* assignments, long long* seeds, long long* counts,
               long long* n_clusters_host, cudaStream_t st);
int gather_columns_dev(const double* X, long long ld, const int* idx, long long m, int F, double* Xout,
                       cudaStream_t st);
int knn_dist_dense_dev(const double* D, long long n, int k, int K, double* kth, double* cum, cudaStream_t st);
int hist_rows_dev(const double* values, long long n_mc, int n_dim, const long long* order, int normalize, double lo,
                  double range, const double* edges_host, int n_edges, int use_ecdf, double* out, cudaStream_t st);
int order_stats_dev(const double* values, long long count, int n_ranks, const long long* ranks_host, double* out_host,
                    long long* n_nan_host, cudaStream_t st);
int knn_dist_features_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                          const double* group_weight, int k, int K, long long block_rows, double* kth, double* cum,
                          cudaStream_t st);
int distance_rows_dev(const double* X, long long n, int n_groups, const int32_t* group_len, const double* group_weight,
                      long long row0, long long row1, double* out, cudaStream_t st);
int distance_sparse_count_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                              const double* group_weight, double thr, int* count, cudaStream_t st);
int distance_sparse_fill_dev(const double* X, long long n, int n_groups, const int32_t* group_len,
                             const double* group_weight, double thr, const long long* colptr, long long nnz,
                             int* rowval, double* nzval, cudaStream_t st);
int pair_feature_counter(int enable, long long* value_out, cudaStream_t st);
int dbscan_sparse_dev(long long n, const long long* colptr, const int* rowval, const double* nzval, double value,
                      double eps, int minpts, long long* assignments, long long* seeds, long long* counts,
                      long long* n_clusters_host, cudaStream_t st);


// Julia Base (:)(start, step, stop) for Float64 (base/twiceprecision.jl): length and elements.  First the
// "rational lift" (exact arithmetic when start, step, stop are ratios of small integers, e.g. 0.0:0.1:0.7),
// then the literal fallback.
static void jl_rat(double x, long long* num, long long* den) {
    double y = x;
    long long a = 1, d = 1, b = 0, c = 0;
    const double m = 16777216.0; 
    while (fabs(y) <= m) {
        const long long f = (long long)y; 
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

// incidence matrix of second_order_kuramoto by edge, compressed (layout: systems.cuh): the kernels never scan zeros
void build_edge_csr(const mcbb_system_desc* sys, std::vector<double>* out) {
    const int N = sys->n_osc, E = sys->n_edges;
    std::vector<double> node, val;
    out->assign(E + 1, 0.);
    for (int e = 0; e < E; e++) {
        (*out)[e] = (double)node.size();
        for (int i = 0; i < N; i++) {
            const double w = sys->mat_a[i + (size_t)N * e];
            if (w != 0.) {
                node.push_back((double)i);
                val.push_back(w);
            }
        }
    }
    (*out)[E] = (double)node.size();
    out->insert(out->end(), node.begin(), node.end());
    out->insert(out->end(), val.begin(), val.end());
}

// 0/1 adjacency (non-zero = true) as compressed rows: [0..N] row starts, then ascending column indices (doubles)
void build_adj_csr(const double* A, int N, std::vector<double>* out) {
    std::vector<double> cols;
    out->assign(N + 1, 0.);
    for (int i = 0; i < N; i++) {
        (*out)[i] = (double)cols.size();
        for (int j = 0; j < N; j++)
            if (A[i + (size_t)N * j] != 0.) cols.push_back((double)j);
    }
    (*out)[N] = (double)cols.size();
    out->insert(out->end(), cols.begin(), cols.end());
}

static int build_sysdev(const mcbb_system_desc* sys, SysDev* out, cudaStream_t st) {
    int len[5];
    if (int rc = table_lengths(sys, len)) return rc;
    SysDev S = {};
    S.system = sys->system;
    S.n_osc = sys->n_osc;
    S.n_dim = sys->n_dim;
    S.n_edges = sys->n_edges;
    if (S.n_osc < 1 || S.n_dim < 1) MCBB_FAIL("system descriptor: N and n_dim must be positive");
    const double* src[5] = {sys->mat_a, sys->mat_b, sys->vec_a, sys->vec_b, sys->vec_c};
    const double* dst[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<double> edge_csr, adj1, adj2;
    if (sys->system == MCBB_SYS_SECOND_ORDER_KURAMOTO) {
        build_edge_csr(sys, &edge_csr);
        src[1] = edge_csr.data();
        len[1] = (int)edge_csr.size();
    }
    if (sys->system == MCBB_SYS_STUART_LANDAU) {
        build_adj_csr(sys->mat_a, sys->n_osc, &adj1);
        build_adj_csr(sys->mat_b, sys->n_osc, &adj2);
        src[0] = adj1.data();
        len[0] = (int)adj1.size();
        src[1] = adj2.data();
        len[1] = (int)adj2.size();
    }
    for (int k = 0; k < 5; k++) {
        if (len[k] == 0) continue;
        double* d = (double*)scratch_get(k, sizeof(double) * (size_t)len[k]);
        if (!d) MCBB_FAIL("solve: device allocation of a system table failed");
        MCBB_CUDA_OK(cudaMemcpyAsync(d, src[k], sizeof(double) * (size_t)len[k], cudaMemcpyHostToDevice, st));
        dst[k] = d;
    }
    MCBB_CUDA_OK(cudaStreamSynchronize(st));  // host tables are caller-owned: do not outlive the call
    S.mat_a = dst[0];
    S.mat_b = dst[1];
    S.vec_a = dst[2];
    S.vec_b = dst[3];
    S.vec_c = dst[4];
    S.len_mat_a = len[0];
    S.len_mat_b = len[1];
    S.len_vec_a = len[2];
    S.len_vec_b = len[3];
    S.len_vec_c = len[4];
    memcpy(S.scalars, sys->scalars, sizeof(S.scalars));
    S.n_par = sys->n_par;
    for (int p = 0; p < MCBB_MAX_PAR; p++) S.par_slot[p] = sys->par_slot[p];
    *out = S;
    return 0;
}

}  // namespace mcbb

using namespace mcbb;

extern "C" {

const char* mcbb_version(void) { return "mcbb_b200 0.1.0 (sm_100a)"; }

int mcbb_init(int device) {
    MCBB_CUDA_OK(cudaSetDevice(device));
    MCBB_CUDA_OK(cudaFree(0));
    // The library is compiled for sm_100a only (tcgen05 / tensor memory have no forward-compatible PTX): fail here,
    // with a readable message, rather than with "no kernel image" at the first launch.
    int major = 0, minor = 0;
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10 || minor != 0)
        MCBB_FAIL("mcbb_init: device " + std::to_string(device) + " has compute capability " + std::to_string(major) + "." +
                  std::to_string(minor) + "; this library is built for sm_100a (B200) only");
    return 0;
}

int mcbb_finalize(void) {
    scratch_free_all();
    return 0;
}

size_t mcbb_last_error(char* buf, size_t len) {
    const size_t n = t_last_error.size();
    if (buf && len > 0) {
        const size_t k = n < len - 1 ? n : len - 1;
        memcpy(buf, t_last_error.data(), k);
        buf[k] = 0;
    }
    return n;
}

int mcbb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* hbm_bytes) {
    int dev = 0;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    MCBB_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)p.totalGlobalMem;
    return 0;
}

int64_t mcbb_kernel_launch_count(void) { return g_launch_count.load(); }

int mcbb_pair_feature_counter(int32_t enable, int64_t* value_out, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    return pair_feature_counter(enable, (long long*)value_out, (cudaStream_t)stream);
}

int mcbb_test_hook_set(const char* key, int64_t value) {
    if (!key) MCBB_FAIL("test hook: NULL key");
    std::lock_guard<std::mutex> lk(g_tune_mu);
    if (value == INT64_MIN)
        g_tune.erase(key);
    else
        g_tune[key] = value;
    return 0;
}

int mcbb_tsave_array(const mcbb_solver_opts* o, int32_t n_t, double rel_transient_time, double* t_save,
                     int32_t* n_out) {
    if (!o || !n_out) MCBB_FAIL("tsave_array: NULL argument");
    if (o->alg == MCBB_ALG_FUNCTIONMAP) {  // setup_mc_prob.jl:1175-1180
        const double tot = nearbyint((o->t1 - o->t0) * (1 - rel_transient_time));
        const double start = o->t1 - tot;
        *n_out = julia_range(start, 1., o->t1 - 1, t_save, 1 << 30);
        return 0;
    }
    if (n_t < 1) MCBB_FAIL("tsave_array: N_t must be positive");
    const double tot = (o->t1 - o->t0) * (1 - rel_transient_time);  // setup_mc_prob.jl:1166-1172
    const double start = o->t1 - tot;
    const double dt = tot / n_t;
    *n_out = julia_range(start, dt, o->t1 - dt, t_save, 1 << 30);
    return 0;
}

int mcbb_solve_features_dev(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                            int64_t n_mc, const double* ic_dev, const double* par_dev, const double* t_save_dev,
                            int32_t n_t, double* feat_dev, double* matrix_dev, double* global_dev,
                            int32_t* retcode_dev, int64_t* nsteps_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    return mcbb_solve_features_saved_dev(sys, opts, feat, n_mc, ic_dev, par_dev, t_save_dev, n_t, feat_dev, matrix_dev,
                                         global_dev, retcode_dev, nsteps_dev, MCBB_SAVE_NONE, nullptr, stream);
}

// one-shot per-thread request set by mcbb_set_solver_parameter_var and consumed by the next solve of this thread
struct SolverVarRequest {
    int field = MCBB_SOLVER_NONE;
    const double* vals = nullptr;
    long long n = 0;
};
static thread_local SolverVarRequest t_solver_var;

int mcbb_set_solver_parameter_var(int32_t solver_field, const double* solver_vals, int64_t n_vals) {
    if (solver_field == MCBB_SOLVER_NONE) {
        t_solver_var = SolverVarRequest();
        return 0;
    }
    if (solver_field < MCBB_SOLVER_ABSTOL || solver_field > MCBB_SOLVER_DTMIN) MCBB_FAIL("unknown solver option id");
    if (!solver_vals || n_vals < 1) MCBB_FAIL("solver parameter variation: NULL or empty value array");
    t_solver_var.field = solver_field;
    t_solver_var.vals = solver_vals;
    t_solver_var.n = n_vals;
    return 0;
}

int mcbb_solve_features_saved_dev(const mcbb_system_desc* sys, const mcbb_solver_opts* opts,
                                  const mcbb_feat_opts* feat, int64_t n_mc, const double* ic_dev,
                                  const double* par_dev, const double* t_save_dev, int32_t n_t, double* feat_dev,
                                  double* matrix_dev, double* global_dev, int32_t* retcode_dev, int64_t* nsteps_dev,
                                  int32_t save_mode, double* saved_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    const SolverVarRequest sv = t_solver_var;  // applies to this call only
    t_solver_var = SolverVarRequest();
    if (!sys || !opts || !feat) MCBB_FAIL("solve: NULL descriptor");
    if (sv.field != MCBB_SOLVER_NONE && sv.n != n_mc) MCBB_FAIL("solver parameter variation: need one value per run");
    if (save_mode != MCBB_SAVE_NONE && save_mode != MCBB_SAVE_ALL && save_mode != MCBB_SAVE_LAST)
        MCBB_FAIL("solve: unknown save mode");
    if (save_mode != MCBB_SAVE_NONE && !saved_dev) MCBB_FAIL("solve: sample export requested but saved_out is NULL");
    if (n_mc < 1) MCBB_FAIL("solve: N_mc must be positive");
    if (!ic_dev || !t_save_dev || !feat_dev) MCBB_FAIL("solve: NULL ic / t_save / feat_out");
    if (sys->n_par > 0 && !par_dev) MCBB_FAIL("solve: NULL par with n_par > 0");
    cudaStream_t st = (cudaStream_t)stream;
    SysDev S;
    if (int rc = build_sysdev(sys, &S, st)) return rc;
    FeatDev F;
    std::vector<int> filter_host;
    if (int rc = resolve_feat(sys, feat, &F, &filter_host)) return rc;
    if (F.n_global > 0 && !global_dev) MCBB_FAIL("solve: global measures requested but global_out is NULL");
    if (F.matrix_meas != MCBB_MAT_NONE && !matrix_dev) MCBB_FAIL("solve: a matrix measure is requested but matrix_out is NULL");
    if (!filter_host.empty()) {
        int* fd = (int*)scratch_get(5, sizeof(int) * filter_host.size());
        if (!fd) MCBB_FAIL("solve: device allocation failed");
        MCBB_CUDA_OK(cudaMemcpyAsync(fd, filter_host.data(), sizeof(int) * filter_host.size(), cudaMemcpyHostToDevice, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        F.filter = fd;
    }
    SolverDev o = resolve_solver(opts);
    SolveIO io = {};
    io.n_mc = n_mc;
    io.ic = ic_dev;
    io.par = par_dev;
    io.t_save = t_save_dev;
    io.n_t = n_t;
    io.feat = feat_dev;
    io.matrix = matrix_dev;
    io.glob = global_dev;
    io.retcode = retcode_dev;
    io.nsteps = (long long*)nsteps_dev;
    io.work_counter = (unsigned long long*)scratch_get(6, 64);
    if (!io.work_counter) MCBB_FAIL("solve: device allocation failed");
    if (F.need_kl) {
        io.ckpt = (double*)scratch_get(7, sizeof(double) * (size_t)(2 * S.n_dim + 4) * (size_t)n_mc);
        if (!io.ckpt) MCBB_FAIL("solve: device allocation of the KL checkpoint buffer failed");
    }
    if (sv.field != MCBB_SOLVER_NONE) {
        io.solver_par = sv.vals;
        io.solver_field = sv.field;
    }
    if (save_mode != MCBB_SAVE_NONE) {
        io.saved = saved_dev;
        io.save_mode = save_mode;
        // samples a failed run never reached read as NaN (all-ones bytes are a quiet NaN)
        const size_t cnt = (size_t)n_mc * S.n_dim * (save_mode == MCBB_SAVE_ALL ? (size_t)n_t : 1);
        MCBB_CUDA_OK(cudaMemsetAsync(saved_dev, 0xFF, sizeof(double) * cnt, st));
    }
    return solve_features_launch(sys, S, o, F, io, st);
}

int mcbb_solve_features(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                        int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                        double* feat_out, double* matrix_out, double* global_out, int32_t* retcode_out,
                        int64_t* nsteps_out) {
    ApiGuard guard_((cudaStream_t)0);
    return mcbb_solve_features_saved(sys, opts, feat, n_mc, ic, par, t_save, n_t, feat_out, matrix_out, global_out,
                                     retcode_out, nsteps_out, MCBB_SAVE_NONE, nullptr);
}

int mcbb_solve_features_saved(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                              int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                              double* feat_out, double* matrix_out, double* global_out, int32_t* retcode_out,
                              int64_t* nsteps_out, int32_t save_mode, double* saved_out) {
    ApiGuard guard_((cudaStream_t)0);
    if (!sys || !opts || !feat) MCBB_FAIL("solve: NULL descriptor");
    if (save_mode != MCBB_SAVE_NONE && !saved_out) MCBB_FAIL("solve: sample export requested but saved_out is NULL");
    if (save_mode < MCBB_SAVE_NONE || save_mode > MCBB_SAVE_LAST) MCBB_FAIL("solve: unknown save mode");
    const SolverVarRequest sv_host = t_solver_var;  // host values: uploaded below, then handed to the device entry
    t_solver_var = SolverVarRequest();
    if (sv_host.field != MCBB_SOLVER_NONE && sv_host.n != n_mc) MCBB_FAIL("solver parameter variation: need one value per run");
    if (n_mc < 1) MCBB_FAIL("solve: N_mc must be positive");
    if (!ic || !t_save || !feat_out) MCBB_FAIL("solve: NULL ic / t_save / feat_out");
    const bool cplx = sys->system == MCBB_SYS_STUART_LANDAU;
    const int n_state = cplx ? sys->n_dim / 2 : sys->n_dim;
    const int nf = feat->n_filter > 0 ? feat->n_filter : n_state;
    const size_t b_ic = sizeof(double) * (size_t)n_mc * sys->n_dim;
    const size_t b_par = sizeof(double) * (size_t)n_mc * (sys->n_par > 0 ? sys->n_par : 1);
    const size_t b_ts = sizeof(double) * (size_t)n_t;
    const size_t b_feat = sizeof(double) * (size_t)n_mc * nf * feat->n_meas;
    const size_t b_glob = sizeof(double) * (size_t)n_mc * (feat->n_global > 0 ? feat->n_global : 1);
    double* d_ic = (double*)scratch_get(30, b_ic);
    double* d_par = (double*)scratch_get(31, b_par);
    double* d_ts = (double*)scratch_get(32, b_ts);
    double* d_feat = (double*)scratch_get(33, b_feat);
    double* d_glob = (double*)scratch_get(34, b_glob);
    const int nbins = feat->corr_nbins > 0 ? feat->corr_nbins : 30;
    const size_t b_mat = sizeof(double) * (size_t)n_mc * nbins;
    double* d_mat = (double*)scratch_get(37, b_mat);
    if (!d_mat) MCBB_FAIL("solve: device allocation failed");
    int* d_ret = (int*)scratch_get(35, sizeof(int) * (size_t)n_mc);
    long long* d_ns = (long long*)scratch_get(36, sizeof(long long) * 2 * (size_t)n_mc);
    if (!d_ic || !d_par || !d_ts || !d_feat || !d_glob || !d_ret || !d_ns) MCBB_FAIL("solve: device allocation failed");
    const size_t b_saved = save_mode == MCBB_SAVE_NONE
                               ? 0
                               : sizeof(double) * (size_t)n_mc * sys->n_dim * (save_mode == MCBB_SAVE_ALL ? (size_t)n_t : 1);
    double* d_saved = nullptr;
    if (b_saved) {
        d_saved = (double*)scratch_get(64, b_saved);
        if (!d_saved) MCBB_FAIL("solve: device allocation of the sample export buffer failed");
    }
    cudaStream_t st = 0;
    // a matrix measure no kernel wrote must read as NaN, never as stale scratch (all-ones bytes = quiet NaN)
    if (feat->matrix_meas != MCBB_MAT_NONE) MCBB_CUDA_OK(cudaMemsetAsync(d_mat, 0xFF, b_mat, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_ic, ic, b_ic, cudaMemcpyHostToDevice, st));
    if (sys->n_par > 0) {
        if (!par) MCBB_FAIL("solve: NULL par with n_par > 0");
        MCBB_CUDA_OK(cudaMemcpyAsync(d_par, par, b_par, cudaMemcpyHostToDevice, st));
    }
    MCBB_CUDA_OK(cudaMemcpyAsync(d_ts, t_save, b_ts, cudaMemcpyHostToDevice, st));
    if (sv_host.field != MCBB_SOLVER_NONE) {
        double* d_sv = (double*)scratch_get(65, sizeof(double) * (size_t)n_mc);
        if (!d_sv) MCBB_FAIL("solve: device allocation failed");
        MCBB_CUDA_OK(cudaMemcpyAsync(d_sv, sv_host.vals, sizeof(double) * (size_t)n_mc, cudaMemcpyHostToDevice, st));
        t_solver_var.field = sv_host.field;
        t_solver_var.vals = d_sv;
        t_solver_var.n = n_mc;
    }
    if (int rc = mcbb_solve_features_saved_dev(sys, opts, feat, n_mc, d_ic, d_par, d_ts, n_t, d_feat,
                                               feat->matrix_meas != MCBB_MAT_NONE ? d_mat : nullptr,
                                               feat->n_global > 0 ? d_glob : nullptr, d_ret, (int64_t*)d_ns,
                                               save_mode, d_saved, (void*)st))
        return rc;
    if (b_saved) MCBB_CUDA_OK(cudaMemcpyAsync(saved_out, d_saved, b_saved, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(feat_out, d_feat, b_feat, cudaMemcpyDeviceToHost, st));
    if (global_out && feat->n_global > 0)
        MCBB_CUDA_OK(cudaMemcpyAsync(global_out, d_glob, b_glob, cudaMemcpyDeviceToHost, st));
    if (retcode_out) MCBB_CUDA_OK(cudaMemcpyAsync(retcode_out, d_ret, sizeof(int) * (size_t)n_mc, cudaMemcpyDeviceToHost, st));
    if (nsteps_out)
        MCBB_CUDA_OK(cudaMemcpyAsync(nsteps_out, d_ns, sizeof(long long) * 2 * (size_t)n_mc, cudaMemcpyDeviceToHost, st));
    if (matrix_out && feat->matrix_meas != MCBB_MAT_NONE)
        MCBB_CUDA_OK(cudaMemcpyAsync(matrix_out, d_mat, b_mat, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_distance_matrix_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                             const double* group_weight, double* D_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!X_dev || !D_dev || !group_len || !group_weight) MCBB_FAIL("distance_matrix: NULL argument");
    if (n < 1 || n_groups < 1) MCBB_FAIL("distance_matrix: empty input");
    return distance_matrix_dev(X_dev, n, n_groups, group_len, group_weight, D_dev, (cudaStream_t)stream);
}

int mcbb_distance_matrix(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                         const double* group_weight, double* D_out) {
    ApiGuard guard_((cudaStream_t)0);
    if (!X || !D_out || !group_len || !group_weight) MCBB_FAIL("distance_matrix: NULL argument");
    if (n < 1 || n_groups < 1) MCBB_FAIL("distance_matrix: empty input");
    size_t F = 0;
    for (int g = 0; g < n_groups; g++) F += group_len[g] > 0 ? group_len[g] : 0;
    double* d_X = (double*)scratch_get(40, sizeof(double) * F * (size_t)n);
    double* d_D = (double*)scratch_get(41, sizeof(double) * (size_t)n * (size_t)n);
    if (!d_X || !d_D) MCBB_FAIL("distance_matrix: device allocation failed (n x n does not fit?)");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_X, X, sizeof(double) * F * (size_t)n, cudaMemcpyHostToDevice, st));
    if (int rc = distance_matrix_dev(d_X, n, n_groups, group_len, group_weight, d_D, st)) return rc;
    MCBB_CUDA_OK(cudaMemcpyAsync(D_out, d_D, sizeof(double) * (size_t)n * (size_t)n, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_dbscan_dense_dev(const double* D_dev, int64_t n, double eps, int32_t minpts, int64_t* assignments_dev,
                          int64_t* seeds_dev, int64_t* counts_dev, int64_t* n_clusters_host, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!D_dev || !assignments_dev || !seeds_dev || !counts_dev || !n_clusters_host) MCBB_FAIL("dbscan: NULL argument");
    return dbscan_dense_dev(D_dev, n, eps, minpts, (long long*)assignments_dev, (long long*)seeds_dev,
                            (long long*)counts_dev, (long long*)n_clusters_host, (cudaStream_t)stream);
}

static int dbscan_copy_back(long long n, long long* d_a, long long* d_s, long long* d_c, int64_t* assignments,
                            int64_t* seeds, int64_t* counts, cudaStream_t st) {
    MCBB_CUDA_OK(cudaMemcpyAsync(assignments, d_a, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(seeds, d_s, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(counts, d_c, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_dbscan_dense(const double* D, int64_t n, double eps, int32_t minpts, int64_t* assignments, int64_t* seeds,
                      int64_t* counts, int64_t* n_clusters) {
    ApiGuard guard_((cudaStream_t)0);
    if (!D || !assignments || !seeds || !counts || !n_clusters) MCBB_FAIL("dbscan: NULL argument");
    if (n < 2) MCBB_FAIL("At least two data points are required.");
    double* d_D = (double*)scratch_get(41, sizeof(double) * (size_t)n * (size_t)n);
    long long* d_a = (long long*)scratch_get(42, sizeof(long long) * (size_t)n);
    long long* d_s = (long long*)scratch_get(43, sizeof(long long) * (size_t)n);
    long long* d_c = (long long*)scratch_get(44, sizeof(long long) * (size_t)n);
    if (!d_D || !d_a || !d_s || !d_c) MCBB_FAIL("dbscan: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_D, D, sizeof(double) * (size_t)n * (size_t)n, cudaMemcpyHostToDevice, st));
    long long k = 0;
    if (int rc = dbscan_dense_dev(d_D, n, eps, minpts, d_a, d_s, d_c, &k, st)) return rc;
    *n_clusters = k;
    return dbscan_copy_back(n, d_a, d_s, d_c, assignments, seeds, counts, st);
}

int mcbb_dbscan_features_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                             const double* group_weight, double eps, int32_t minpts, int64_t* assignments_dev,
                             int64_t* seeds_dev, int64_t* counts_dev, int64_t* n_clusters_host, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!X_dev || !assignments_dev || !seeds_dev || !counts_dev || !n_clusters_host || !group_len || !group_weight)
        MCBB_FAIL("dbscan: NULL argument");
    return dbscan_features_dev(X_dev, n, n_groups, group_len, group_weight, eps, minpts, (long long*)assignments_dev,
                               (long long*)seeds_dev, (long long*)counts_dev, (long long*)n_clusters_host,
                               (cudaStream_t)stream);
}

int mcbb_dbscan_features(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                         const double* group_weight, double eps, int32_t minpts, int64_t* assignments,
                         int64_t* seeds, int64_t* counts, int64_t* n_clusters) {
    ApiGuard guard_((cudaStream_t)0);
    if (!X || !assignments || !seeds || !counts || !n_clusters || !group_len || !group_weight)
        MCBB_FAIL("dbscan: NULL argument");
    if (n < 2) MCBB_FAIL("At least two data points are required.");
    size_t F = 0;
    for (int g = 0; g < n_groups; g++) F += group_len[g] > 0 ? group_len[g] : 0;
    double* d_X = (double*)scratch_get(40, sizeof(double) * F * (size_t)n);
    long long* d_a = (long long*)scratch_get(42, sizeof(long long) * (size_t)n);
    long long* d_s = (long long*)scratch_get(43, sizeof(long long) * (size_t)n);
    long long* d_c = (long long*)scratch_get(44, sizeof(long long) * (size_t)n);
    if (!d_X || !d_a || !d_s || !d_c) MCBB_FAIL("dbscan: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_X, X, sizeof(double) * F * (size_t)n, cudaMemcpyHostToDevice, st));
    long long k = 0;
    if (int rc = dbscan_features_dev(d_X, n, n_groups, group_len, group_weight, eps, minpts, d_a, d_s, d_c, &k, st))
        return rc;
    *n_clusters = k;
    return dbscan_copy_back(n, d_a, d_s, d_c, assignments, seeds, counts, st);
}

int mcbb_dbscan_rows_count_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, double eps, int32_t minpts, int64_t row0, int64_t row1,
                               int32_t* count_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!X_dev || !count_dev || !group_len || !group_weight) MCBB_FAIL("dbscan rows: NULL argument");
    return rows_count_dev(X_dev, n, n_groups, group_len, group_weight, eps, minpts, row0, row1, count_dev, nullptr,
                          nullptr, (cudaStream_t)stream);
}
int mcbb_dbscan_rows_count_lists_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                                     const double* group_weight, double eps, int32_t minpts, int64_t row0, int64_t row1,
                                     int32_t* count_dev, int32_t* nbr_dev, int32_t* nfill_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!X_dev || !count_dev || !nbr_dev || !nfill_dev || !group_len || !group_weight) MCBB_FAIL("dbscan rows: NULL argument");
    return rows_count_dev(X_dev, n, n_groups, group_len, group_weight, eps, minpts, row0, row1, count_dev, nbr_dev,
                          nfill_dev, (cudaStream_t)stream);
}
int mcbb_dbscan_rows_union_dev(const double* Xcore_dev, int64_t n_core, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, double eps, int64_t row0, int64_t row1,
                               int32_t* parent_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if ((n_core > 0 && (!Xcore_dev || !parent_dev)) || !group_len || !group_weight) MCBB_FAIL("dbscan rows: NULL argument");
    return rows_union_dev(Xcore_dev, n_core, n_groups, group_len, group_weight, eps, row0, row1, parent_dev, 1,
                          (cudaStream_t)stream);
}
int mcbb_dbscan_rows_union_more_dev(const double* Xcore_dev, int64_t n_core, int32_t n_groups, const int32_t* group_len,
                                    const double* group_weight, double eps, int64_t row0, int64_t row1,
                                    int32_t* parent_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if ((n_core > 0 && (!Xcore_dev || !parent_dev)) || !group_len || !group_weight) MCBB_FAIL("dbscan rows: NULL argument");
    return rows_union_dev(Xcore_dev, n_core, n_groups, group_len, group_weight, eps, row0, row1, parent_dev, 0,
                          (cudaStream_t)stream);
}
int mcbb_dbscan_merge_parents_dev(int32_t* parent_dev, const int32_t* other_parent_dev, int64_t n_core, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (n_core > 0 && (!parent_dev || !other_parent_dev)) MCBB_FAIL("dbscan merge: NULL argument");
    return merge_parents_dev(parent_dev, other_parent_dev, n_core, (cudaStream_t)stream);
}
int mcbb_dbscan_rows_border_dev(const double* Xnon_dev, int64_t n_non, int64_t row0, int64_t row1,
                                const double* Xcore_dev, int64_t n_core, const int32_t* core_seed_dev,
                                int32_t n_groups, const int32_t* group_len, const double* group_weight, double eps,
                                int32_t* best_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!group_len || !group_weight) MCBB_FAIL("dbscan rows: NULL argument");
    return rows_border_dev(Xnon_dev, n_non, row0, row1, Xcore_dev, n_core, core_seed_dev, n_groups, group_len,
                           group_weight, eps, best_dev, (cudaStream_t)stream);
}
int mcbb_dbscan_labels_dev(const int32_t* root_label_dev, int64_t n, int64_t* assignments_dev, int64_t* seeds_dev,
                           int64_t* counts_dev, int64_t* n_clusters_host, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!root_label_dev || !assignments_dev || !seeds_dev || !counts_dev || !n_clusters_host)
        MCBB_FAIL("dbscan labels: NULL argument");
    return labels_dev(root_label_dev, n, (long long*)assignments_dev, (long long*)seeds_dev, (long long*)counts_dev,
                      (long long*)n_clusters_host, (cudaStream_t)stream);
}
int mcbb_gather_columns_dev(const double* X_dev, int64_t ld, const int32_t* idx_dev, int64_t m, int32_t F,
                            double* Xout_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (m > 0 && (!X_dev || !idx_dev || !Xout_dev)) MCBB_FAIL("gather: NULL argument");
    return gather_columns_dev(X_dev, ld, idx_dev, m, F, Xout_dev, (cudaStream_t)stream);
}

int mcbb_knn_dist_dense_dev(const double* D_dev, int64_t n, int32_t k, int32_t K, double* kth_dev, double* cum_dev,
                            void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!D_dev || (!kth_dev && !cum_dev)) MCBB_FAIL("k_dist: NULL argument");
    return knn_dist_dense_dev(D_dev, n, k, K, kth_dev, cum_dev, (cudaStream_t)stream);
}
int mcbb_knn_dist_dense(const double* D, int64_t n, int32_t k, int32_t K, double* kth_out, double* cum_out) {
    ApiGuard guard_((cudaStream_t)0);
    if (!D || (!kth_out && !cum_out)) MCBB_FAIL("k_dist: NULL argument");
    if (n < 1) MCBB_FAIL("k_dist: empty matrix");
    double* d_D = (double*)scratch_get(41, sizeof(double) * (size_t)n * (size_t)n);
    double* d_k = (double*)scratch_get(48, sizeof(double) * (size_t)n);
    double* d_c = (double*)scratch_get(49, sizeof(double) * (size_t)n);
    if (!d_D || !d_k || !d_c) MCBB_FAIL("k_dist: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_D, D, sizeof(double) * (size_t)n * (size_t)n, cudaMemcpyHostToDevice, st));
    if (int rc = knn_dist_dense_dev(d_D, n, k, K, kth_out ? d_k : nullptr, cum_out ? d_c : nullptr, st)) return rc;
    if (kth_out) MCBB_CUDA_OK(cudaMemcpyAsync(kth_out, d_k, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (cum_out) MCBB_CUDA_OK(cudaMemcpyAsync(cum_out, d_c, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_order_stats_dev(const double* values_dev, int64_t count, int32_t n_ranks, const int64_t* ranks, double* out,
                         int64_t* n_nan, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!values_dev || (n_ranks > 0 && (!ranks || !out))) MCBB_FAIL("order_stats: NULL argument");
    return order_stats_dev(values_dev, count, n_ranks, (const long long*)ranks, out, (long long*)n_nan, (cudaStream_t)stream);
}
int mcbb_order_stats(const double* values, int64_t count, int32_t n_ranks, const int64_t* ranks, double* out, int64_t* n_nan) {
    ApiGuard guard_((cudaStream_t)0);
    if (!values || (n_ranks > 0 && (!ranks || !out))) MCBB_FAIL("order_stats: NULL argument");
    if (count < 1) MCBB_FAIL("order_stats: empty input");
    double* d_v = (double*)scratch_get(100, sizeof(double) * (size_t)count);
    if (!d_v) MCBB_FAIL("order_stats: device allocation failed");
    MCBB_CUDA_OK(cudaMemcpyAsync(d_v, values, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, 0));
    return order_stats_dev(d_v, count, n_ranks, (const long long*)ranks, out, (long long*)n_nan, 0);
}
int mcbb_hist_rows_dev(const double* values_dev, int64_t n_mc, int32_t n_dim, const int64_t* order_dev, int32_t normalize,
                       double lo, double range, const double* edges, int32_t n_edges, int32_t use_ecdf, double* rows_dev,
                       void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!values_dev || !edges || !rows_dev) MCBB_FAIL("hist_rows: NULL argument");
    if (n_edges < 2 || n_dim < 1) MCBB_FAIL("hist_rows: needs at least two edges and one dimension");
    return hist_rows_dev(values_dev, n_mc, n_dim, (const long long*)order_dev, normalize, lo, range, edges, n_edges, use_ecdf,
                         rows_dev, (cudaStream_t)stream);
}
int mcbb_hist_rows(const double* values, int64_t n_mc, int32_t n_dim, const int64_t* order, int32_t normalize, double lo,
                   double range, const double* edges, int32_t n_edges, int32_t use_ecdf, double* rows_out) {
    ApiGuard guard_((cudaStream_t)0);
    if (!values || !edges || !rows_out) MCBB_FAIL("hist_rows: NULL argument");
    if (n_edges < 2 || n_dim < 1 || n_mc < 1) MCBB_FAIL("hist_rows: needs at least two edges, one dimension and one run");
    const size_t nv = (size_t)n_mc * (size_t)n_dim, no = (size_t)n_mc * (size_t)(n_edges - 1);
    double* d_v = (double*)scratch_get(100, sizeof(double) * nv);
    double* d_o = (double*)scratch_get(101, sizeof(double) * no);
    long long* d_ord = order ? (long long*)scratch_get(102, sizeof(long long) * (size_t)n_mc) : nullptr;
    if (!d_v || !d_o || (order && !d_ord)) MCBB_FAIL("hist_rows: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_v, values, sizeof(double) * nv, cudaMemcpyHostToDevice, st));
    if (order) MCBB_CUDA_OK(cudaMemcpyAsync(d_ord, order, sizeof(long long) * (size_t)n_mc, cudaMemcpyHostToDevice, st));
    if (int rc = hist_rows_dev(d_v, n_mc, n_dim, d_ord, normalize, lo, range, edges, n_edges, use_ecdf, d_o, st)) return rc;
    MCBB_CUDA_OK(cudaMemcpyAsync(rows_out, d_o, sizeof(double) * no, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_knn_dist_features_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, int32_t k, int32_t K, int64_t block_rows, double* kth_dev,
                               double* cum_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!X_dev || !group_len || !group_weight || (!kth_dev && !cum_dev)) MCBB_FAIL("k_dist: NULL argument");
    if (n_groups < 1) MCBB_FAIL("k_dist: empty input");
    return knn_dist_features_dev(X_dev, n, n_groups, group_len, group_weight, k, K, block_rows, kth_dev, cum_dev,
                                 (cudaStream_t)stream);
}
int mcbb_knn_dist_features(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                           const double* group_weight, int32_t k, int32_t K, int64_t block_rows, double* kth_out,
                           double* cum_out) {
    ApiGuard guard_((cudaStream_t)0);
    if (!X || !group_len || !group_weight || (!kth_out && !cum_out)) MCBB_FAIL("k_dist: NULL argument");
    if (n < 1 || n_groups < 1) MCBB_FAIL("k_dist: empty input");
    size_t F = 0;
    for (int g = 0; g < n_groups; g++) F += group_len[g] > 0 ? group_len[g] : 0;
    double* d_X = (double*)scratch_get(40, sizeof(double) * F * (size_t)n);
    double* d_k = (double*)scratch_get(48, sizeof(double) * (size_t)n);
    double* d_c = (double*)scratch_get(49, sizeof(double) * (size_t)n);
    if (!d_X || !d_k || !d_c) MCBB_FAIL("k_dist: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_X, X, sizeof(double) * F * (size_t)n, cudaMemcpyHostToDevice, st));
    if (int rc = knn_dist_features_dev(d_X, n, n_groups, group_len, group_weight, k, K, block_rows,
                                       kth_out ? d_k : nullptr, cum_out ? d_c : nullptr, st))
        return rc;
    if (kth_out) MCBB_CUDA_OK(cudaMemcpyAsync(kth_out, d_k, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (cum_out) MCBB_CUDA_OK(cudaMemcpyAsync(cum_out, d_c, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_distance_sparse_count_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                                   const double* group_weight, double sparse_threshold, int32_t* count_dev,
                                   void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!X_dev || !count_dev || !group_len || !group_weight) MCBB_FAIL("distance_matrix: NULL argument");
    return distance_sparse_count_dev(X_dev, n, n_groups, group_len, group_weight, sparse_threshold, count_dev,
                                     (cudaStream_t)stream);
}
int mcbb_distance_sparse_fill_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                                  const double* group_weight, double sparse_threshold, const int64_t* colptr_dev,
                                  int64_t nnz, int32_t* rowval_dev, double* nzval_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!X_dev || !colptr_dev || !rowval_dev || !nzval_dev || !group_len || !group_weight)
        MCBB_FAIL("distance_matrix: NULL argument");
    return distance_sparse_fill_dev(X_dev, n, n_groups, group_len, group_weight, sparse_threshold,
                                    (const long long*)colptr_dev, nnz, rowval_dev, nzval_dev, (cudaStream_t)stream);
}
int mcbb_dbscan_sparse_dev(int64_t n, const int64_t* colptr_dev, const int32_t* rowval_dev, const double* nzval_dev,
                           double value, double eps, int32_t minpts, int64_t* assignments_dev, int64_t* seeds_dev,
                           int64_t* counts_dev, int64_t* n_clusters_host, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!colptr_dev || !assignments_dev || !seeds_dev || !counts_dev || !n_clusters_host) MCBB_FAIL("dbscan: NULL argument");
    return dbscan_sparse_dev(n, (const long long*)colptr_dev, rowval_dev, nzval_dev, value, eps, minpts,
                             (long long*)assignments_dev, (long long*)seeds_dev, (long long*)counts_dev,
                             (long long*)n_clusters_host, (cudaStream_t)stream);
}

static size_t total_features(int32_t n_groups, const int32_t* group_len) {
    size_t F = 0;
    for (int g = 0; g < n_groups; g++) F += group_len[g] > 0 ? group_len[g] : 0;
    return F;
}

int mcbb_distance_matrix_rows_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                                  const double* group_weight, int64_t row0, int64_t row1, double* rows_dev, void* stream) {
    ApiGuard guard_((cudaStream_t)stream);
    if (!X_dev || !rows_dev || !group_len || !group_weight) MCBB_FAIL("distance_matrix: NULL argument");
    return distance_rows_dev(X_dev, n, n_groups, group_len, group_weight, row0, row1, rows_dev, (cudaStream_t)stream);
}
int mcbb_distance_matrix_rows(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                              const double* group_weight, int64_t row0, int64_t row1, double* rows_out) {
    ApiGuard guard_((cudaStream_t)0);
    if (!X || !rows_out || !group_len || !group_weight) MCBB_FAIL("distance_matrix: NULL argument");
    if (n < 1 || n_groups < 1) MCBB_FAIL("distance_matrix: empty input");
    if (row0 < 0 || row1 > n || row0 > row1) MCBB_FAIL("distance_matrix: bad row range");
    const size_t F = total_features(n_groups, group_len);
    const size_t nb = (size_t)(row1 - row0) * (size_t)n;
    double* d_X = (double*)scratch_get(40, sizeof(double) * F * (size_t)n);
    double* d_D = (double*)scratch_get(41, sizeof(double) * (nb > 0 ? nb : 1));
    if (!d_X || !d_D) MCBB_FAIL("distance_matrix: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_X, X, sizeof(double) * F * (size_t)n, cudaMemcpyHostToDevice, st));
    if (int rc = distance_rows_dev(d_X, n, n_groups, group_len, group_weight, row0, row1, d_D, st)) return rc;
    MCBB_CUDA_OK(cudaMemcpyAsync(rows_out, d_D, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_distance_sparse_count(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, double sparse_threshold, int32_t* count_out) {
    ApiGuard guard_((cudaStream_t)0);
    if (!X || !count_out || !group_len || !group_weight) MCBB_FAIL("distance_matrix: NULL argument");
    if (n < 1 || n_groups < 1) MCBB_FAIL("distance_matrix: empty input");
    const size_t F = total_features(n_groups, group_len);
    double* d_X = (double*)scratch_get(40, sizeof(double) * F * (size_t)n);
    int* d_c = (int*)scratch_get(30, sizeof(int) * (size_t)n);
    if (!d_X || !d_c) MCBB_FAIL("distance_matrix: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_X, X, sizeof(double) * F * (size_t)n, cudaMemcpyHostToDevice, st));
    if (int rc = distance_sparse_count_dev(d_X, n, n_groups, group_len, group_weight, sparse_threshold, d_c, st)) return rc;
    MCBB_CUDA_OK(cudaMemcpyAsync(count_out, d_c, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_distance_sparse_fill(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                              const double* group_weight, double sparse_threshold, const int64_t* colptr,
                              int32_t* rowval_out, double* nzval_out) {
    ApiGuard guard_((cudaStream_t)0);
    if (!X || !colptr || !rowval_out || !nzval_out || !group_len || !group_weight) MCBB_FAIL("distance_matrix: NULL argument");
    if (n < 1 || n_groups < 1) MCBB_FAIL("distance_matrix: empty input");
    const long long nnz = colptr[n];
    if (colptr[0] != 0 || nnz < n) MCBB_FAIL("sparse distance matrix: colptr must start at 0 and hold the diagonal");
    const size_t F = total_features(n_groups, group_len);
    double* d_X = (double*)scratch_get(40, sizeof(double) * F * (size_t)n);
    long long* d_p = (long long*)scratch_get(31, sizeof(long long) * (size_t)(n + 1));
    int* d_r = (int*)scratch_get(32, sizeof(int) * (size_t)nnz);
    double* d_v = (double*)scratch_get(33, sizeof(double) * (size_t)nnz);
    if (!d_X || !d_p || !d_r || !d_v) MCBB_FAIL("distance_matrix: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_X, X, sizeof(double) * F * (size_t)n, cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_p, colptr, sizeof(long long) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    if (int rc = distance_sparse_fill_dev(d_X, n, n_groups, group_len, group_weight, sparse_threshold, d_p, nnz, d_r, d_v, st))
        return rc;
    MCBB_CUDA_OK(cudaMemcpyAsync(rowval_out, d_r, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(nzval_out, d_v, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int mcbb_dbscan_sparse(int64_t n, const int64_t* colptr, const int32_t* rowval, const double* nzval, double value,
                       double eps, int32_t minpts, int64_t* assignments, int64_t* seeds, int64_t* counts,
                       int64_t* n_clusters) {
    ApiGuard guard_((cudaStream_t)0);
    if (!colptr || !assignments || !seeds || !counts || !n_clusters) MCBB_FAIL("dbscan: NULL argument");
    if (n < 2) MCBB_FAIL("At least two data points are required.");
    const long long nnz = colptr[n];
    if (nnz > 0 && (!rowval || !nzval)) MCBB_FAIL("dbscan: NULL argument");
    long long* d_p = (long long*)scratch_get(31, sizeof(long long) * (size_t)(n + 1));
    int* d_r = (int*)scratch_get(32, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    double* d_v = (double*)scratch_get(33, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    long long* d_a = (long long*)scratch_get(42, sizeof(long long) * (size_t)n);
    long long* d_s = (long long*)scratch_get(43, sizeof(long long) * (size_t)n);
    long long* d_c = (long long*)scratch_get(44, sizeof(long long) * (size_t)n);
    if (!d_p || !d_r || !d_v || !d_a || !d_s || !d_c) MCBB_FAIL("dbscan: device allocation failed");
    cudaStream_t st = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(d_p, colptr, sizeof(long long) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    if (nnz > 0) {
        MCBB_CUDA_OK(cudaMemcpyAsync(d_r, rowval, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
        MCBB_CUDA_OK(cudaMemcpyAsync(d_v, nzval, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    }
    long long k = 0;
    if (int rc = dbscan_sparse_dev(n, d_p, d_r, d_v, value, eps, minpts, d_a, d_s, d_c, &k, st)) return rc;
    *n_clusters = k;
    return dbscan_copy_back(n, d_a, d_s, d_c, assignments, seeds, counts, st);
}

}  // extern "C"