// common.cuh — error plumbing, launch accounting and the flat device-side descriptors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/mcbb_b200.h"

#ifndef MCBB_HD
#define MCBB_HD __host__ __device__ __forceinline__
#endif

namespace mcbb {

void set_error(const std::string& msg);
extern std::atomic<long long> g_launch_count;

#define MCBB_CUDA_OK(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::mcbb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                              ":" + std::to_string(__LINE__) + ")");                                \
            return 2;                                                                               \
        }                                                                                           \
    } while (0)

#define MCBB_FAIL(msg)          \
    do {                        \
        ::mcbb::set_error(msg); \
        return 1;               \
    } while (0)

#define MCBB_LAUNCHED()                      \
    do {                                     \
        ::mcbb::g_launch_count.fetch_add(1); \
        MCBB_CUDA_OK(cudaGetLastError());    \
    } while (0)

// ---- flat descriptors handed to kernels by value -----------------------------------------------------
struct SysDev {
    int system, n_osc, n_dim, n_edges;
    const double *mat_a, *mat_b, *vec_a, *vec_b, *vec_c;  // device (global) tables
    int len_mat_a, len_mat_b, len_vec_a, len_vec_b, len_vec_c;
    double scalars[MCBB_N_SCALARS];
    int n_par;
    int par_slot[MCBB_MAX_PAR];
};

struct SolverDev {  // mcbb_solver_opts with the DiffEq defaults resolved
    int alg;
    double t0, t1, abstol, reltol, dt_user, dtmin, dtmax, qmin, qmax, gamma, beta1, beta2, qoldinit;
    long long maxiters;
};

struct FeatDev {
    int n_meas;
    int meas[MCBB_MAX_MEAS];
    int comp[MCBB_MAX_MEAS];
    int n_filter;         // resolved (>0)
    const int* filter;    // device, 1-based; nullptr = identity
    int n_comp;           // 1, or 2 for complex states
    int n_ch;             // channels = n_filter * n_comp
    int cyclic;
    int need_kl;          // any KL_HIST measure
    int matrix_meas, corr_nbins;
    int n_global;
    int glob[MCBB_MAX_GLOBAL];
    int kl_bins;          // odd
    double kl_nstd, kl_tol;
};

struct SolveIO {
    long long n_mc;
    const double* ic;      // [n_dim][n_mc]
    const double* par;     // [n_par][n_mc]
    const double* t_save;  // [n_t]
    int n_t;
    double* feat;          // [n_meas][n_filter][n_mc]
    double* matrix;        // [nbins][n_mc] or null
    double* glob;          // [n_global][n_mc] or null
    int* retcode;          // [n_mc] or null
    long long* nsteps;     // [2][n_mc] or null
    double* ckpt;          // [2*n_dim + 4][n_mc] checkpoint scratch for the KL pass, or null
    unsigned long long* work_counter;  // dynamic trajectory queue
    // optional export of the saved samples themselves (get_trajectory, eval_clustering.jl:1484-1492; the end state
    // the continuation analysis restarts from, eval_mc_prob.jl:210-216).  Lane kernels only.
    double* saved;   // save_mode 1: [n_mc][n_t][n_dim]; save_mode 2: the last sample, [n_dim][n_mc] (the ic layout)
    int save_mode;   // 0 = none
    // optional per-run value of ONE solver option (SolverParameterVar, src/solver_parametervar.jl).  Lane kernels only.
    const double* solver_par;  // [n_mc] or null
    int solver_field;          // enum mcbb_solver_field
};

SolverDev resolve_solver(const mcbb_solver_opts* o);

// host-side cache of device scratch buffers (freed by mcbb_finalize)
void* scratch_get(int slot, size_t bytes);
size_t scratch_size(int slot);  // bytes currently cached for the slot on this device
void scratch_free_all();

// test / experiment switches: set only through mcbb_test_hook_set (never from the environment), 0 / default in production
long long tune_get(const char* key, long long dflt);

// serialises the library's per-device working state across host threads and streams (api.cu)
struct ApiGuard {
    explicit ApiGuard(cudaStream_t st);
    ~ApiGuard();
    ApiGuard(const ApiGuard&) = delete;
    ApiGuard& operator=(const ApiGuard&) = delete;
  private:
    int dev_;
};

}  // namespace mcbb
