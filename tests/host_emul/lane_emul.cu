// lane_emul.cu — DEVELOPMENT / TEST HARNESS ONLY (never shipped, never loaded by the product).
// Compiles the very same per-lane integrator source (csrc/lane_core.cuh) for the HOST and runs it one lane at
// a time, so the state machine can be checked against the oracle in the CPU-only container before a GPU run.
#include <vector>

#include "../../mcbb.jl_b200/csrc/lane_core.cuh"

namespace mcbb {
int table_lengths(const mcbb_system_desc* s, int len[5]);
int resolve_feat(const mcbb_system_desc* sys, const mcbb_feat_opts* f, FeatDev* out, std::vector<int>* filter_host);
void build_edge_csr(const mcbb_system_desc* sys, std::vector<double>* out);
void build_adj_csr(const double* A, int N, std::vector<double>* out);
}
using namespace mcbb;

struct HostWorkQ {
    long long ctr = 0;
    long long next() { return ctr++; }
    bool all_done(bool f) { return f; }
};

template <int SYS, bool SOLVAR>
static void run_host_impl(const SysDev& S, const SolverDev& o, const FeatDev& F, const SolveIO& io) {
    const int n = S.n_dim;
    std::vector<double> st((size_t)(9 + SysTraits<SYS>::n_scratch) * n + 8), acc((size_t)ACC_SLOTS_CORR * F.n_ch + 8);
    std::vector<double> como((size_t)F.n_ch * F.n_ch + 8);
    std::vector<unsigned short> bins((size_t)F.n_ch * F.kl_bins + 8);
    LaneShared L;
    L.t_save = io.t_save;
    L.tabs = SysTabs{S.mat_a, S.mat_b, S.vec_a, S.vec_b, S.vec_c};
    HostWorkQ wq;
    lane_main<SYS, HostWorkQ, SOLVAR>(S, o, F, io, L, st.data(), acc.data(), bins.data(), como.data(), 1, wq);
}

template <int SYS>
static void run_host(const SysDev& S, const SolverDev& o, const FeatDev& F, const SolveIO& io) {
    if (io.solver_par)
        run_host_impl<SYS, true>(S, o, F, io);
    else
        run_host_impl<SYS, false>(S, o, F, io);
}

extern "C" int emul_solve_features(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                                   int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                                   double* feat_out, double* global_out, int32_t* retcode_out, int64_t* nsteps_out,
                                   double* matrix_out, int32_t save_mode, double* saved_out, int32_t solver_field,
                                   const double* solver_vals) {
    int len[5];
    if (table_lengths(sys, len)) return 1;
    SysDev S = {};
    S.system = sys->system; S.n_osc = sys->n_osc; S.n_dim = sys->n_dim; S.n_edges = sys->n_edges;
    S.mat_a = sys->mat_a; S.mat_b = sys->mat_b; S.vec_a = sys->vec_a; S.vec_b = sys->vec_b; S.vec_c = sys->vec_c;
    for (int i = 0; i < MCBB_N_SCALARS; i++) S.scalars[i] = sys->scalars[i];
    S.n_par = sys->n_par;
    for (int i = 0; i < MCBB_MAX_PAR; i++) S.par_slot[i] = sys->par_slot[i];
    std::vector<double> csr;
    if (S.system == MCBB_SYS_SECOND_ORDER_KURAMOTO) {
        build_edge_csr(sys, &csr);
        S.mat_b = csr.data();
    }
    std::vector<double> adj1, adj2;
    if (S.system == MCBB_SYS_STUART_LANDAU) {
        build_adj_csr(sys->mat_a, sys->n_osc, &adj1);
        build_adj_csr(sys->mat_b, sys->n_osc, &adj2);
        S.mat_a = adj1.data();
        S.mat_b = adj2.data();
    }
    FeatDev F;
    std::vector<int> filt;
    if (resolve_feat(sys, feat, &F, &filt)) return 1;
    if (!filt.empty()) F.filter = filt.data();
    SolverDev o = resolve_solver(opts);
    std::vector<double> ckpt((size_t)(2 * S.n_dim + 4) * n_mc);
    SolveIO io = {};
    io.n_mc = n_mc; io.ic = ic; io.par = par; io.t_save = t_save; io.n_t = n_t; io.feat = feat_out;
    io.glob = global_out; io.matrix = matrix_out; io.retcode = retcode_out; io.nsteps = (long long*)nsteps_out; io.ckpt = ckpt.data();
    io.saved = save_mode ? saved_out : nullptr;
    io.save_mode = save_mode;
    io.solver_par = solver_field ? solver_vals : nullptr;
    io.solver_field = solver_field;
    switch (S.system) {
    case MCBB_SYS_LOGISTIC: run_host<MCBB_SYS_LOGISTIC>(S, o, F, io); break;
    case MCBB_SYS_KURAMOTO: run_host<MCBB_SYS_KURAMOTO>(S, o, F, io); break;
    case MCBB_SYS_KURAMOTO_NETWORK: run_host<MCBB_SYS_KURAMOTO_NETWORK>(S, o, F, io); break;
    case MCBB_SYS_SECOND_ORDER_KURAMOTO: run_host<MCBB_SYS_SECOND_ORDER_KURAMOTO>(S, o, F, io); break;
    case MCBB_SYS_NONLOCAL_KURAMOTO_RING: run_host<MCBB_SYS_NONLOCAL_KURAMOTO_RING>(S, o, F, io); break;
    case MCBB_SYS_STUART_LANDAU: run_host<MCBB_SYS_STUART_LANDAU>(S, o, F, io); break;
    case MCBB_SYS_ROESSLER_NETWORK: run_host<MCBB_SYS_ROESSLER_NETWORK>(S, o, F, io); break;
    default: return 1;
    }
    return 0;
}
