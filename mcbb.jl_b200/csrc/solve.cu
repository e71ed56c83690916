// solve.cu — launch side of the ensemble integrator + fused featurizer (C-ABI hot entry 1).
#include <algorithm>
#include <vector>

#include "lane_core.cuh"

namespace mcbb {

struct DevWorkQ {
    unsigned long long* ctr;
    __device__ __forceinline__ long long next() { return (long long)atomicAdd(ctr, 1ULL); }
    __device__ __forceinline__ bool all_done(bool f) { return __all_sync(0xffffffffu, f); }
};

struct LaneLayout {  // byte offsets into dynamic shared memory
    int off_tsave, off_tabs[5], off_state, off_acc, off_bins, off_como;
    int stage_tabs;
    int n_slots, n_acc;
};

template <int SYS, bool SOLVAR>
__global__ void __launch_bounds__(128) k_solve_lane(const SysDev S, const SolverDev o, const FeatDev F,
                                                    const SolveIO io, const LaneLayout lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, bs = blockDim.x;
    double* s_tsave = (double*)(smem + lay.off_tsave);
    for (int i = tid; i < io.n_t; i += bs) s_tsave[i] = io.t_save[i];
    LaneShared L;
    L.t_save = s_tsave;
    const double* src[5] = {S.mat_a, S.mat_b, S.vec_a, S.vec_b, S.vec_c};
    const int len[5] = {S.len_mat_a, S.len_mat_b, S.len_vec_a, S.len_vec_b, S.len_vec_c};
    const double* tp[5];
    for (int k = 0; k < 5; k++) {
        if (lay.stage_tabs && len[k] > 0) {
            double* dst = (double*)(smem + lay.off_tabs[k]);
            for (int i = tid; i < len[k]; i += bs) dst[i] = src[k][i];
            tp[k] = dst;
        } else {
            tp[k] = src[k];
        }
    }
    L.tabs.mat_a = tp[0];
    L.tabs.mat_b = tp[1];
    L.tabs.vec_a = tp[2];
    L.tabs.vec_b = tp[3];
    L.tabs.vec_c = tp[4];
    __syncthreads();
    double* st = (double*)(smem + lay.off_state) + tid;
    double* acc = (double*)(smem + lay.off_acc) + tid;
    unsigned short* bins = (unsigned short*)(smem + lay.off_bins) + tid;
    double* como = (double*)(smem + lay.off_como) + tid;
    DevWorkQ wq{io.work_counter};
    lane_main<SYS, DevWorkQ, SOLVAR>(S, o, F, io, L, st, acc, bins, como, bs, wq);
}

// Large systems (state does not fit shared memory, no batched kernel for the system): same lane state machine,
// per-lane state vectors in an L2/HBM-resident scratch laid out [slot][dim][lane] (coalesced across lanes).
// Correct for any size; throughput is bounded by L2/HBM traffic of the stage vectors, not by the FP64 pipe.
template <int SYS, bool SOLVAR>
__global__ void __launch_bounds__(128) k_solve_lane_gmem(const SysDev S, const SolverDev o, const FeatDev F,
                                                         const SolveIO io, double* state, double* accs,
                                                         unsigned short* bins_g, double* como_g) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    double* s_tsave = (double*)smem;
    for (int i = tid; i < io.n_t; i += blockDim.x) s_tsave[i] = io.t_save[i];
    __syncthreads();
    LaneShared L;
    L.t_save = s_tsave;
    L.tabs = SysTabs{S.mat_a, S.mat_b, S.vec_a, S.vec_b, S.vec_c};
    const int lanes = gridDim.x * blockDim.x;
    const int lane = blockIdx.x * blockDim.x + tid;
    DevWorkQ wq{io.work_counter};
    lane_main<SYS, DevWorkQ, SOLVAR>(S, o, F, io, L, state + lane, accs + lane, bins_g + lane, como_g + lane, lanes, wq);
}

SolverDev resolve_solver(const mcbb_solver_opts* o) {
    SolverDev s;
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

static int n_scratch_for(int sys) {
    switch (sys) {
    case MCBB_SYS_KURAMOTO_NETWORK:
    case MCBB_SYS_KURAMOTO:
    case MCBB_SYS_NONLOCAL_KURAMOTO_RING:
        return 2;
    default:
        return 0;
    }
}

int table_lengths(const mcbb_system_desc* s, int len[5]) {
    const int N = s->n_osc;
    for (int k = 0; k < 5; k++) len[k] = 0;
    switch (s->system) {
    case MCBB_SYS_LOGISTIC:
        if (s->n_dim != 1) MCBB_FAIL("logistic: n_dim must be 1");
        break;
    case MCBB_SYS_KURAMOTO:
        len[2] = N;
        if (s->n_dim != N) MCBB_FAIL("kuramoto: n_dim must equal N");
        break;
    case MCBB_SYS_KURAMOTO_NETWORK:
        len[0] = N * N;
        len[2] = N;
        if (s->n_dim != N) MCBB_FAIL("kuramoto_network: n_dim must equal N");
        break;
    case MCBB_SYS_SECOND_ORDER_KURAMOTO:
        len[0] = N * s->n_edges;
        len[2] = N;
        if (s->n_dim != 2 * N) MCBB_FAIL("second_order_kuramoto: n_dim must equal 2N");
        break;
    case MCBB_SYS_NONLOCAL_KURAMOTO_RING:
        len[0] = 2 * N - 1;
        if (s->n_dim != N) MCBB_FAIL("non_local_kuramoto_ring: n_dim must equal N");
        break;
    case MCBB_SYS_STUART_LANDAU:
        len[0] = N * N;
        len[1] = N * N;
        if (s->n_dim != 2 * N) MCBB_FAIL("stuart_landau: n_dim must equal 2N (interleaved re, im)");
        break;
    case MCBB_SYS_ROESSLER_NETWORK:
        len[0] = N * N;
        len[2] = len[3] = len[4] = N;
        if (s->n_dim != 3 * N) MCBB_FAIL("roessler_network: n_dim must equal 3N");
        break;
    default:
        MCBB_FAIL("unknown system id");
    }
    const double* p[5] = {s->mat_a, s->mat_b, s->vec_a, s->vec_b, s->vec_c};
    for (int k = 0; k < 5; k++)
        if (len[k] > 0 && p[k] == nullptr) MCBB_FAIL("system descriptor: a required table pointer is NULL");
    for (int q = 0; q < s->n_par; q++)
        if (s->par_slot[q] < 0 || s->par_slot[q] >= MCBB_N_SCALARS) MCBB_FAIL("par_slot out of range");
    if (s->n_par < 0 || s->n_par > MCBB_MAX_PAR) MCBB_FAIL("n_par out of range");
    return 0;
}

int resolve_feat(const mcbb_system_desc* sys, const mcbb_feat_opts* f, FeatDev* out, std::vector<int>* filter_host) {
    FeatDev F = {};
    const bool cplx = sys->system == MCBB_SYS_STUART_LANDAU;
    const int n_state = cplx ? sys->n_dim / 2 : sys->n_dim;
    if (f->n_meas < 1) MCBB_FAIL("No per dimension measures");  // eval_mc_prob.jl:79-81
    if (f->n_meas > MCBB_MAX_MEAS) MCBB_FAIL("too many per dimension measures");
    F.n_meas = f->n_meas;
    F.n_comp = cplx ? 2 : 1;
    bool have_mean[2] = {false, false}, have_std[2] = {false, false};
    for (int m = 0; m < f->n_meas; m++) {
        F.meas[m] = f->meas[m];
        F.comp[m] = cplx ? f->meas_comp[m] : 0;
        if (F.comp[m] < 0 || F.comp[m] > 1) MCBB_FAIL("meas_comp must be 0 or 1");
        const int c = F.comp[m];
        switch (f->meas[m]) {
        case MCBB_MEAS_MEAN:
            have_mean[c] = true;
            break;
        case MCBB_MEAS_STD:
            have_std[c] = true;
            break;
        case MCBB_MEAS_KL_HIST:
            if (!have_mean[c] || !have_std[c])
                MCBB_FAIL("KL_HIST needs a preceding MEAN and STD of the same component (past_measures)");
            F.need_kl = 1;
            break;
        default:
            MCBB_FAIL("unknown measure id");
        }
    }
    filter_host->clear();
    if (f->n_filter > 0) {
        if (!f->state_filter) MCBB_FAIL("state_filter is NULL");
        for (int i = 0; i < f->n_filter; i++) {
            if (f->state_filter[i] < 1 || f->state_filter[i] > n_state) MCBB_FAIL("state_filter index out of range");
            filter_host->push_back(f->state_filter[i]);
        }
        F.n_filter = f->n_filter;
    } else {
        F.n_filter = n_state;
    }
    F.n_ch = F.n_filter * F.n_comp;
    F.cyclic = f->cyclic_setback != 0;
    F.matrix_meas = f->matrix_meas;
    F.corr_nbins = f->corr_nbins > 0 ? f->corr_nbins : 30;
    if (F.matrix_meas != MCBB_MAT_NONE && F.matrix_meas != MCBB_MAT_CORR_ECDF && F.matrix_meas != MCBB_MAT_CORR_HIST)
        MCBB_FAIL("unknown matrix measure id");
    if (F.matrix_meas != MCBB_MAT_NONE) {
        if (F.corr_nbins > 64) MCBB_FAIL("correlation_ecdf: at most 64 bins");
    }
    F.n_global = f->n_global;
    if (F.n_global < 0 || F.n_global > MCBB_MAX_GLOBAL) MCBB_FAIL("n_global out of range");
    for (int g = 0; g < F.n_global; g++) {
        F.glob[g] = f->global_meas[g];
        if (F.glob[g] != MCBB_GLOB_SUM) MCBB_FAIL("unknown global measure id");
    }
    F.kl_bins = f->kl_bins > 0 ? f->kl_bins : 31;
    if (F.kl_bins % 2 == 0) F.kl_bins += 1;  // eval_mc_prob.jl:320-322
    if (F.kl_bins > 127) MCBB_FAIL("kl_bins > 127");
    F.kl_nstd = f->kl_n_stds > 0 ? f->kl_n_stds : 3.;
    F.kl_tol = f->kl_sig_tol > 0 ? f->kl_sig_tol : 1e-4;
    *out = F;
    return 0;
}

template <int SYS, bool SOLVAR>
static int launch_lane_impl(const SysDev& S, const SolverDev& o, const FeatDev& F, const SolveIO& io, cudaStream_t st) {
    const int n = S.n_dim;
    LaneLayout lay = {};
    lay.n_slots = 9 + n_scratch_for(SYS);
    const bool corr = F.matrix_meas != MCBB_MAT_NONE;
    const int n_pairs = corr ? F.n_ch * (F.n_ch - 1) / 2 : 0;
    lay.n_acc = corr ? ACC_SLOTS_CORR : (F.need_kl ? ACC_SLOTS_KL : ACC_SLOTS);
    const int lens[5] = {S.len_mat_a, S.len_mat_b, S.len_vec_a, S.len_vec_b, S.len_vec_c};
    size_t tab_bytes = 0;
    for (int k = 0; k < 5; k++) tab_bytes += sizeof(double) * (size_t)lens[k];
    const size_t per_lane = sizeof(double) * ((size_t)lay.n_slots * n + (size_t)lay.n_acc * F.n_ch + (size_t)n_pairs) +
                            (F.need_kl ? sizeof(unsigned short) * (size_t)F.n_ch * F.kl_bins : 0);
    const size_t fixed = sizeof(double) * (size_t)io.n_t + 64;
    const size_t cap = 227 * 1024;
    lay.stage_tabs = (fixed + tab_bytes + per_lane * 32 <= cap) && tab_bytes <= 48 * 1024;
    const size_t base = fixed + (lay.stage_tabs ? tab_bytes : 0);
    int bs = 0;
    for (int cand : {128, 64, 32}) {
        if (base + per_lane * cand + 16 <= cap) {
            bs = cand;
            break;
        }
    }
    if (tune_get("force_gmem", 0)) bs = 0;  // experiment switch: L2/HBM-resident lane state
    if (bs == 0) {
        // state in global memory.  The kernel is latency bound (dependent loads of the stage vectors), so fill the
        // machine: 8 CTAs of 64 lanes per SM (~76 k lanes; at most 8 GB of stage vectors, HBM-resident)
        int dev = 0, sms = 148;
        MCBB_CUDA_OK(cudaGetDevice(&dev));
        MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int block = 64;
        long long lanes = std::min<long long>((long long)sms * 8 * block, std::max<long long>(block, (8LL << 30) / (long long)per_lane));
        lanes = std::min<long long>(lanes, (io.n_mc + block - 1) / block * block);
        lanes = std::max<long long>(block, lanes / block * block);
        const int grid = (int)(lanes / block);
        double* state = (double*)scratch_get(8, sizeof(double) * (size_t)lay.n_slots * n * lanes);
        double* accs = (double*)scratch_get(9, sizeof(double) * (size_t)lay.n_acc * F.n_ch * lanes);
        unsigned short* bins_g = (unsigned short*)scratch_get(28, sizeof(unsigned short) * (size_t)(F.need_kl ? F.n_ch * F.kl_bins : 1) * lanes);
        double* como_g = (double*)scratch_get(29, sizeof(double) * (size_t)(n_pairs > 0 ? n_pairs : 1) * lanes);
        if (!state || !accs || !bins_g || !como_g) MCBB_FAIL("solve: device allocation of the lane state failed");
        const size_t smem_g = sizeof(double) * (size_t)io.n_t + 16;
        MCBB_CUDA_OK(cudaFuncSetAttribute(k_solve_lane_gmem<SYS, SOLVAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        MCBB_CUDA_OK(cudaMemsetAsync(io.work_counter, 0, sizeof(unsigned long long), st));
        k_solve_lane_gmem<SYS, SOLVAR><<<grid, block, smem_g, st>>>(S, o, F, io, state, accs, bins_g, como_g);
        MCBB_LAUNCHED();
        return 0;
    }
    // prefer two resident CTAs per SM when that costs nothing
    while (bs > 32 && base + per_lane * bs > cap / 2 && base + per_lane * (bs / 2) <= cap / 2) bs /= 2;
    {   // small ensembles (C1: 2 500 runs): 128-lane CTAs would occupy 20 of the 148 SMs — spread the lanes out
        int dev0 = 0, sms0 = 148;
        MCBB_CUDA_OK(cudaGetDevice(&dev0));
        MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0));
        while (bs > 32 && (io.n_mc + bs - 1) / bs < (long long)sms0) bs /= 2;
    }
    size_t off = 0;
    lay.off_tsave = (int)off;
    off += sizeof(double) * (size_t)io.n_t;
    for (int k = 0; k < 5; k++) {
        lay.off_tabs[k] = (int)off;
        if (lay.stage_tabs) off += sizeof(double) * (size_t)lens[k];
    }
    lay.off_state = (int)off;
    off += sizeof(double) * (size_t)lay.n_slots * n * bs;
    lay.off_acc = (int)off;
    off += sizeof(double) * (size_t)lay.n_acc * F.n_ch * bs;
    lay.off_como = (int)off;
    off += sizeof(double) * (size_t)n_pairs * bs;
    lay.off_bins = (int)off;
    off += F.need_kl ? sizeof(unsigned short) * (size_t)F.n_ch * F.kl_bins * bs : 0;
    const size_t smem = off;
    MCBB_CUDA_OK(cudaFuncSetAttribute(k_solve_lane<SYS, SOLVAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, per_sm = 1;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MCBB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_lane<SYS, SOLVAR>, bs, smem));
    if (per_sm < 1) per_sm = 1;
    long long want = (io.n_mc + bs - 1) / bs;
    long long grid = std::min<long long>(want, (long long)sms * per_sm);  // persistent CTAs, dynamic queue
    MCBB_CUDA_OK(cudaMemsetAsync(io.work_counter, 0, sizeof(unsigned long long), st));
    k_solve_lane<SYS, SOLVAR><<<(unsigned)grid, bs, smem, st>>>(S, o, F, io, lay);
    MCBB_LAUNCHED();
    return 0;
}

template <int SYS>
static int launch_lane(const SysDev& S, const SolverDev& o, const FeatDev& F, const SolveIO& io, cudaStream_t st) {
    return io.solver_par ? launch_lane_impl<SYS, true>(S, o, F, io, st) : launch_lane_impl<SYS, false>(S, o, F, io, st);
}

int launch_batched_umma(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F,
                        const SolveIO& io, cudaStream_t st, bool* handled);
int launch_batched_imma(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F,
                        const SolveIO& io, cudaStream_t st, bool* handled);
int launch_batched_umma_sl(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F,
                           const SolveIO& io, cudaStream_t st, bool* handled);
int launch_group(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F, const SolveIO& io,
                 cudaStream_t st, bool* handled);
int launch_batched_kuramoto(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F,
                            const SolveIO& io, cudaStream_t st, bool* handled);

// everything on the device already; tables in S are device pointers
int solve_features_launch(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F,
                          const SolveIO& io, cudaStream_t st) {
    if (o.alg == MCBB_ALG_FUNCTIONMAP && S.system != MCBB_SYS_LOGISTIC)
        MCBB_FAIL("FunctionMap stepping is only defined for the map systems (logistic)");
    if (o.alg != MCBB_ALG_FUNCTIONMAP && S.system == MCBB_SYS_LOGISTIC)
        MCBB_FAIL("logistic is a map: use MCBB_ALG_FUNCTIONMAP");
    if (o.alg == MCBB_ALG_RK4 && !(o.dt_user > 0) && !(io.solver_par && io.solver_field == MCBB_SOLVER_DT))
        MCBB_FAIL("RK4 is fixed-step: opts.dt must be > 0");
    if (o.alg < 0 || o.alg > 2) MCBB_FAIL("unknown algorithm id");
    if (!(o.t1 > o.t0)) MCBB_FAIL("tspan must be increasing");
    if (io.n_t < 2) MCBB_FAIL("need at least two save times (the first saved sample is dropped)");
    bool handled = false;
    if (io.solver_par && (io.solver_field < MCBB_SOLVER_ABSTOL || io.solver_field > MCBB_SOLVER_DTMIN))
        MCBB_FAIL("unknown solver option id for the per-run solver parameter");
    // the batched kernels keep no per-sample state to export and take their solver options per launch: sample export
    // and per-run solver options run the lane kernels
    if (!io.saved && !io.solver_par) {
        if (int rc = launch_batched_umma(hs, S, o, F, io, st, &handled)) return rc;  // tcgen05 + tensor memory
        if (handled) return 0;
        if (int rc = launch_batched_umma_sl(hs, S, o, F, io, st, &handled)) return rc;  // Stuart-Landau on the same path
        if (handled) return 0;
        if (int rc = launch_batched_imma(hs, S, o, F, io, st, &handled)) return rc;
        if (handled) return 0;
        if (int rc = launch_batched_kuramoto(hs, S, o, F, io, st, &handled)) return rc;
        if (handled) return 0;
        if (int rc = launch_group(hs, S, o, F, io, st, &handled)) return rc;  // small networks: a lane group per trajectory
        if (handled) return 0;
    }
    switch (S.system) {
    case MCBB_SYS_LOGISTIC:
        return launch_lane<MCBB_SYS_LOGISTIC>(S, o, F, io, st);
    case MCBB_SYS_KURAMOTO:
        return launch_lane<MCBB_SYS_KURAMOTO>(S, o, F, io, st);
    case MCBB_SYS_KURAMOTO_NETWORK:
        return launch_lane<MCBB_SYS_KURAMOTO_NETWORK>(S, o, F, io, st);
    case MCBB_SYS_SECOND_ORDER_KURAMOTO:
        return launch_lane<MCBB_SYS_SECOND_ORDER_KURAMOTO>(S, o, F, io, st);
    case MCBB_SYS_NONLOCAL_KURAMOTO_RING:
        return launch_lane<MCBB_SYS_NONLOCAL_KURAMOTO_RING>(S, o, F, io, st);
    case MCBB_SYS_STUART_LANDAU:
        return launch_lane<MCBB_SYS_STUART_LANDAU>(S, o, F, io, st);
    case MCBB_SYS_ROESSLER_NETWORK:
        return launch_lane<MCBB_SYS_ROESSLER_NETWORK>(S, o, F, io, st);
    }
    MCBB_FAIL("unknown system id");
}

}  // namespace mcbb
