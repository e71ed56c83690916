/*
 * mcbb_b200.h — C-ABI of libmcbb_b200.so, the B200 (sm_100a) drop-in for the MCBB.jl hot path
 *
 *     DEMCBBProblem / solve  ->  eval_ode_run statistics  ->  distance_matrix  ->  dbscan
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and returns an int status
 * (0 = ok, non-zero = error; text via mcbb_last_error).  No torch / C++ types cross this boundary.
 * The reference is pure Julia and has no FFI of its own for this path; each entry point below names the
 * reference function (file:line under the reference tree) whose work it replaces.  The Julia `ccall`
 * stubs a maintainer would add are in INTEGRATION.md and mcbb.jl_b200/julia/MCBBB200.jl.
 *
 * Conventions
 *   - All matrices follow Julia's column-major layout.  `ic` is N_mc x N_dim with run index fastest
 *     (prob.ic, src/setup_mc_prob.jl:334-368), `par` is N_mc x N_par (prob.par).  That layout is
 *     struct-of-arrays over runs and is consumed as is.
 *   - Host entry points (no suffix) take HOST pointers: the caller owns every buffer, the library copies
 *     to / from the device internally and retains nothing.  `_dev` entry points take DEVICE pointers on the
 *     current device plus a cudaStream_t (passed as void*) and are asynchronous on that stream.
 *   - Labels follow Clustering.jl: 0 = noise, 1..k = cluster id (src/sparse_clustering.jl:138-164).
 *   - Per-trajectory integrator failures are data (retcode_out), not errors
 *     (failure_handling, src/eval_mc_prob.jl:95-126).
 *   - Threads and streams: working state (device scratch, the trajectory queue, constant-bank tables) is kept per
 *     DEVICE.  Every compute entry point takes that device's lock for the duration of the call, and a call arriving on
 *     a different stream than the device's previous call first waits for that stream to drain.  So: any number of host
 *     threads may call concurrently (different devices overlap, one device takes turns), results are never corrupted
 *     by mixing streams, but only ONE stream per device overlaps with the host — use one stream per device.
 *   - Device memory: the library keeps its work buffers cached per device between calls and releases them in
 *     mcbb_finalize.  Measure sets that need the whole saved series of a run (KL-hist, the correlation measures) keep the
 *     saved samples of a CHUNK of runs on the device and reduce them after the chunk is integrated; the chunk is sized to
 *     at most 64 GB and at most half of the device memory that is free at the time of the call (ensembles whose samples
 *     take less than 2 GB never ask), so the results do not depend on it and a busy device only means more chunks.
 */
#ifndef MCBB_B200_H
#define MCBB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCBB_MAX_MEAS 8
#define MCBB_MAX_PAR 4
#define MCBB_MAX_GLOBAL 4
#define MCBB_N_SCALARS 8

/* ---- systems (src/systems.jl) -------------------------------------------------------------------------- */
enum mcbb_system {
    MCBB_SYS_LOGISTIC = 0,               /* systems.jl:34-36    scalars[0]=r                                    */
    MCBB_SYS_KURAMOTO_NETWORK = 1,       /* systems.jl:119-130  scalars[0]=K; vec_a=w[N]; mat_a=A[N x N]          */
    MCBB_SYS_SECOND_ORDER_KURAMOTO = 2,  /* systems.jl:198-202  scalars[0]=damping, [1]=coupling; vec_a=drive[N];
                                            mat_a=incidence E[N x n_edges]; state = (theta[N], omega[N])        */
    MCBB_SYS_NONLOCAL_KURAMOTO_RING = 3, /* systems.jl:230-239  scalars[0]=omega_0, [1]=phase_delay, [2]=fak;
                                            mat_a=G table, 2N-1 values, G[(k-j)+N-1]=coupling_function(fak*(k-j)) */
    MCBB_SYS_STUART_LANDAU = 4,          /* paper/stuartlandau/stuart_landau_sathiyadevi-standard-config.ipynb cell 1
                                            scalars[0]=eps, [1]=lambda, [2]=omega, [3]=P_1, [4]=P_2;
                                            mat_a=A_1[N x N], mat_b=A_2[N x N] (non-zero = true);
                                            state = interleaved (re,im) of Complex{Float64}[N]                  */
    MCBB_SYS_ROESSLER_NETWORK = 5,       /* systems.jl:286-298  scalars[0]=K; vec_a=a[N], vec_b=b[N], vec_c=c[N];
                                            mat_a=L[N x N]; state = (x1,y1,z1,x2,...)                          */
    MCBB_SYS_KURAMOTO = 6                /* systems.jl:85-94    scalars[0]=K; vec_a=w[N] (all-to-all)           */
};

typedef struct mcbb_system_desc {
    int32_t system;                     /* enum mcbb_system                                                     */
    int32_t n_osc;                      /* N: oscillators / nodes (1 for the logistic map)                     */
    int32_t n_dim;                      /* real state dimension: N, 2N (2nd order, Stuart-Landau) or 3N          */
    int32_t n_edges;                    /* SECOND_ORDER_KURAMOTO only: columns of the incidence matrix          */
    const double* mat_a;
    const double* mat_b;
    const double* vec_a;
    const double* vec_b;
    const double* vec_c;
    double scalars[MCBB_N_SCALARS];     /* base values of the scalar parameter fields                           */
    int32_t n_par;                      /* number of swept parameters (columns of `par`)                        */
    int32_t par_slot[MCBB_MAX_PAR];     /* par[:,p] overrides scalars[par_slot[p]] per run — the flat form of
                                           reconstruct(pars; name=>par[i,p]), src/setup_mc_prob.jl:722-762      */
} mcbb_system_desc;

/* ---- integrator (OrdinaryDiffEq 5.26.7 semantics restated; call site src/setup_mc_prob.jl:1106) --------- */
enum mcbb_alg {
    MCBB_ALG_TSIT5 = 0,        /* adaptive Tsit5, PI controller, saveat through the Tsit5 interpolant          */
    MCBB_ALG_RK4 = 1,          /* classical RK4, fixed step `dt` (solve(..., RK4(), adaptive=false, dt=h))     */
    MCBB_ALG_FUNCTIONMAP = 2   /* DiscreteProblem stepping u_{n+1} = f(u_n), dt = 1                            */
};

enum mcbb_retcode {            /* per trajectory, like sol.retcode                                             */
    MCBB_RET_SUCCESS = 0,
    MCBB_RET_MAXITERS = 1,
    MCBB_RET_DTLESSTHANMIN = 2,
    MCBB_RET_UNSTABLE = 3
};

typedef struct mcbb_solver_opts {
    int32_t alg;               /* enum mcbb_alg                                                               */
    int32_t reserved;
    double t0, t1;             /* prob.tspan                                                                  */
    double abstol, reltol;     /* <=0 -> DiffEq defaults 1e-6 / 1e-3                                          */
    double dt;                 /* RK4: fixed step (required). TSIT5: >0 overrides the automatic initial dt    */
    double dtmin, dtmax;       /* <=0 -> eps(Float64) / (t1 - t0)                                             */
    int64_t maxiters;          /* <=0 -> 100000                                                               */
    double qmin, qmax, gamma, beta1, beta2, qoldinit; /* <=0 -> 1/5, 10, 9/10, 7/50, 2/25, 1e-4              */
} mcbb_solver_opts;

/* ---- featurizer: the closure `eval_ode_run` flattened into a descriptor (src/eval_mc_prob.jl:75-198) ----- */
enum mcbb_measure {
    MCBB_MEAS_MEAN = 0,        /* StatsBase.mean                                       eval_mc_prob.jl:194     */
    MCBB_MEAS_STD = 1,         /* StatsBase.std(u; mean=past mean, corrected=true)     eval_mc_prob.jl:195     */
    MCBB_MEAS_KL_HIST = 2      /* empirical_1D_KL_divergence_hist(u, [mu, sig])        eval_mc_prob.jl:310-337 */
};
enum mcbb_matrix_measure {
    MCBB_MAT_NONE = 0,
    MCBB_MAT_CORR_ECDF = 1,    /* correlation_ecdf(sol, nbins)                         eval_mc_prob.jl:462-488 */
    MCBB_MAT_CORR_HIST = 2     /* correlation_hist(sol, nbins): the raw bin weights    eval_mc_prob.jl:472-481 */
};
enum mcbb_global_measure {
    MCBB_GLOB_SUM = 0          /* (x)->sum(x) over sol[state_filter, 2:end]     test/ode_prob_example.jl:54   */
};

typedef struct mcbb_feat_opts {
    int32_t n_meas;                     /* number of per-dimension measures, in output order                   */
    int32_t meas[MCBB_MAX_MEAS];        /* enum mcbb_measure; STD / KL_HIST use the nearest preceding MEAN (and
                                           STD) with the same component, i.e. flag_past_measures=true wiring   */
    int32_t meas_comp[MCBB_MAX_MEAS];   /* complex systems: 0 = real part, 1 = imaginary part; else 0          */
    int32_t n_filter;                   /* length of state_filter; 0 = all dimensions                          */
    const int32_t* state_filter;        /* 1-based indices (oscillator index for complex states)               */
    int32_t cyclic_setback;             /* _cyclic_setback!, eval_mc_prob.jl:286-296                           */
    int32_t matrix_meas;                /* enum mcbb_matrix_measure                                            */
    int32_t corr_nbins;                 /* <=0 -> 30                                                           */
    int32_t n_global;
    int32_t global_meas[MCBB_MAX_GLOBAL];
    int32_t kl_bins;                    /* <=0 -> 31                                                           */
    double kl_n_stds;                   /* <=0 -> 3                                                            */
    double kl_sig_tol;                  /* <=0 -> 1e-4                                                         */
} mcbb_feat_opts;

/* ---- library ------------------------------------------------------------------------------------------- */
const char* mcbb_version(void);
int mcbb_init(int device);                     /* binds the calling thread to `device`; allocates nothing       */
int mcbb_finalize(void);                       /* frees cached device scratch                                   */
size_t mcbb_last_error(char* buf, size_t len); /* copies the last error text (thread-local), returns its length */
int mcbb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* hbm_bytes);
/* number of library kernels launched by this process so far (bench.py's gpu_launches) */
int64_t mcbb_kernel_launch_count(void);

/* Test / experiment hook — NOT part of the drop-in surface.  Sets a named integer switch read by the kernel dispatch
 * (e.g. "no_umma", "umma_jb", "force_gmem", "trace"); value INT64_MIN erases the key.  Nothing in the library reads the
 * process environment to choose a kernel: production callers never call this and get the default dispatch. */
int mcbb_test_hook_set(const char* key, int64_t value);

/* Diagnostic behind bench.py's cluster.roofline: counts the pair x feature terms the distance tiles actually execute
 * on the current device (early exits and saturation skips excluded).  The current value is written to *value_out (may
 * be NULL; synchronises `stream`) BEFORE `enable` is applied: enable > 0 zeroes and switches the counter on, 0 switches
 * it off, < 0 leaves it as it is.  Off by default (one atomic per warp and tile when on). */
int mcbb_pair_feature_counter(int32_t enable, int64_t* value_out, void* stream);

/* FP64 FMA-pipe peak of the current device in TFLOP/s (DFMA micro-benchmark, best of 5) — the roofline
 * denominator of the ensemble integrator; MEASURED_PEAKS.json carries no FP64 figure. */
int mcbb_measure_fp64_fma_peak(double* tflops_out, double* sm_mhz_est_out);

/* number of saved samples a solve will consume; fills t_save (may be NULL to query the length).
 * Restates tsave_array, src/setup_mc_prob.jl:1166-1180 (ODE and Discrete variants). */
int mcbb_tsave_array(const mcbb_solver_opts* opts, int32_t n_t, double rel_transient_time,
                     double* t_save, int32_t* n_out);

/* ---- hot entry 1: ensemble integrate + featurize -------------------------------------------------------
 * Replaces the body of solve(prob::DEMCBBProblem, ...) between t_save and DEMCBBSol construction:
 * DiffEqBase ensemble solve + output_func=eval_ode_run per trajectory (src/setup_mc_prob.jl:1098-1116;
 * plug-in contract custom_solve(prob, t_save), :1102-1104).
 *   ic        N_mc x n_dim   column-major (run fastest); complex states as interleaved (re,im) PER RUN,
 *                            i.e. Julia's Matrix{Complex{Float64}} reinterpreted: dimension index 2*k+c
 *   par       N_mc x n_par   column-major
 *   t_save    n_t save times (the first saved sample is dropped by the featurizer, eval_mc_prob.jl:152)
 *   feat_out  N_mc x n_filter x n_meas   column-major: feat_out[(m*n_filter + d)*N_mc + i]
 *   matrix_out N_mc x corr_nbins (or NULL), global_out N_mc x n_global (or NULL)
 *   retcode_out N_mc (enum mcbb_retcode), nsteps_out N_mc x 2 column-major: [accepted+rejected, rejected]
 * Results are in the caller's run order; the stable sort by parameter (setup_mc_prob.jl:520-536) stays in
 * the host language. */
int mcbb_solve_features(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                        int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                        double* feat_out, double* matrix_out, double* global_out,
                        int32_t* retcode_out, int64_t* nsteps_out);

/* Same with ic / par / t_save and all outputs resident on the current device.  The descriptor structs
 * and the small tables they point to (A, w, ...) stay on the host. */
int mcbb_solve_features_dev(const mcbb_system_desc* sys, const mcbb_solver_opts* opts,
                            const mcbb_feat_opts* feat, int64_t n_mc, const double* ic_dev,
                            const double* par_dev, const double* t_save_dev, int32_t n_t, double* feat_dev,
                            double* matrix_dev, double* global_dev, int32_t* retcode_dev,
                            int64_t* nsteps_dev, void* stream);

/* The same solve, additionally exporting the saved samples themselves — what `sol.solve_command(prob_i)` returns to
 * get_trajectory (src/eval_clustering.jl:1484-1492; setup_mc_prob.jl:1108-1110: saveat = t_save, savestart = false,
 * save_everystep = false), and the end state `sol[end]` the continuation / response analysis restarts its three
 * short runs from (src/eval_mc_prob.jl:210-216).
 *   MCBB_SAVE_ALL   saved_out = N_mc blocks of n_dim x n_t, column-major: saved_out[(i*n_t + k)*n_dim + d]
 *   MCBB_SAVE_LAST  saved_out = N_mc x n_dim column-major, the ic layout: saved_out[d*N_mc + i] = sol_i[end][d]
 * Samples a failed run never reached are NaN.  Complex states are interleaved (re, im) like `ic`.  The export runs
 * the one-trajectory-per-lane kernels (the batched tensor-core kernels keep no per-sample state). */
enum mcbb_save_mode { MCBB_SAVE_NONE = 0, MCBB_SAVE_ALL = 1, MCBB_SAVE_LAST = 2 };
/* Per-run solver options — SolverParameterVar / SolverVarEnsembleProblem (src/solver_parametervar.jl:5-90;
 * test/solver_parametervar.jl varies :reltol): `solve(prob_i, alg; name => par[i], kwargs...)`.  One option of
 * mcbb_solver_opts is overridden per run by solver_vals[i]; the system parameters then usually stay fixed
 * (sys->n_par = 0, par = NULL), as in the reference, but a simultaneous system-parameter sweep is allowed.
 * Set with mcbb_set_solver_parameter_var before a solve_features* call of the SAME host thread; it applies to that one
 * call and is cleared by it.  solver_vals is a HOST array of n_mc values for the host entry points and a DEVICE array
 * for the *_dev ones and must stay valid until that solve returns (the library keeps the pointer, not a copy).
 * Runs the lane kernels. */
enum mcbb_solver_field {
    MCBB_SOLVER_NONE = 0,
    MCBB_SOLVER_ABSTOL = 1,
    MCBB_SOLVER_RELTOL = 2,
    MCBB_SOLVER_DT = 3,
    MCBB_SOLVER_DTMAX = 4,
    MCBB_SOLVER_DTMIN = 5
};
int mcbb_set_solver_parameter_var(int32_t solver_field, const double* solver_vals, int64_t n_vals);

int mcbb_solve_features_saved(const mcbb_system_desc* sys, const mcbb_solver_opts* opts,
                              const mcbb_feat_opts* feat, int64_t n_mc, const double* ic, const double* par,
                              const double* t_save, int32_t n_t, double* feat_out, double* matrix_out,
                              double* global_out, int32_t* retcode_out, int64_t* nsteps_out, int32_t save_mode,
                              double* saved_out);
int mcbb_solve_features_saved_dev(const mcbb_system_desc* sys, const mcbb_solver_opts* opts,
                                  const mcbb_feat_opts* feat, int64_t n_mc, const double* ic_dev,
                                  const double* par_dev, const double* t_save_dev, int32_t n_t, double* feat_dev,
                                  double* matrix_dev, double* global_dev, int32_t* retcode_dev,
                                  int64_t* nsteps_dev, int32_t save_mode, double* saved_dev, void* stream);

/* ---- hot entry 2: distance matrix ----------------------------------------------------------------------
 * Replaces distance_matrix(sol, prob, weights) with the default L1 distance_func
 * (src/eval_clustering.jl:164-254):
 *     D[i,j] = sum_g group_weight[g] * sum_{f in group g} |X[f,i] - X[f,j]|
 * `X` is the feature matrix, n x F column-major (run fastest), F = sum(group_len).  A group is one measure
 * (length n_filter), one matrix measure (length nbins), one global measure or one parameter column
 * (length 1, eval_clustering.jl:218-224); groups are accumulated in order, like the reference's
 * per-measure passes over mat_elements.  D_out is n x n, symmetric, zero diagonal. */
int mcbb_distance_matrix(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                         const double* group_weight, double* D_out);
int mcbb_distance_matrix_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                             const double* group_weight, double* D_dev, void* stream);

/* ---- hot entry 3: DBSCAN -------------------------------------------------------------------------------
 * Replaces Clustering.dbscan(D, eps, minpts) (src/eval_clustering.jl:84; algorithm as vendored in
 * src/sparse_clustering.jl:138-209): neighbourhood {i : D[i,p] < eps} (strict, self included), core iff
 * |neighbourhood| >= minpts.  Output is identical to the sequential algorithm with visit order 1:n:
 * clusters are the connected components of the core graph numbered by ascending smallest core index
 * (= seeds), a non-core point takes the smallest cluster id among cores within eps, else 0.
 *   assignments[n]; seeds[k], counts[k] (buffers of length n); *n_clusters = k.  seeds are 1-based. */
int mcbb_dbscan_dense(const double* D, int64_t n, double eps, int32_t minpts, int64_t* assignments,
                      int64_t* seeds, int64_t* counts, int64_t* n_clusters);
int mcbb_dbscan_dense_dev(const double* D_dev, int64_t n, double eps, int32_t minpts,
                          int64_t* assignments_dev, int64_t* seeds_dev, int64_t* counts_dev,
                          int64_t* n_clusters_host, void* stream);

/* Same labels as mcbb_dbscan_dense(mcbb_distance_matrix(X, ...)), but the n x n matrix is never
 * materialised: distances are recomputed on the fly in shared-memory tiles (the reference's answer to
 * "N^2 does not fit" is distance_matrix_mmap / distance_matrix_sparse, src/eval_clustering.jl:317, 451). */
int mcbb_dbscan_features(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                         const double* group_weight, double eps, int32_t minpts, int64_t* assignments,
                         int64_t* seeds, int64_t* counts, int64_t* n_clusters);
int mcbb_dbscan_features_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                             const double* group_weight, double eps, int32_t minpts,
                             int64_t* assignments_dev, int64_t* seeds_dev, int64_t* counts_dev,
                             int64_t* n_clusters_host, void* stream);

/* ---- eps-selection helpers -----------------------------------------------------------------------------------
 * Replace k_dist / KNN_dist / KNN_dist_relative (src/eval_clustering.jl:1504-1553): every row of the (symmetric)
 * distance matrix is sorted ascending on the device; kth_out[i] = sort(D[i,:])[k] (1-based k, the zero self-distance
 * counts, as in the reference), cum_out[i] = sum(sort(D[i,:])[1:K]).  Either output may be NULL. */
int mcbb_knn_dist_dense(const double* D, int64_t n, int32_t k, int32_t K, double* kth_out, double* cum_out);
int mcbb_knn_dist_dense_dev(const double* D_dev, int64_t n, int32_t k, int32_t K, double* kth_dev,
                            double* cum_dev, void* stream);
/* The same from the feature matrix (layout of mcbb_distance_matrix), for ensembles whose D is never materialised:
 * D is produced in blocks of `block_rows` rows (0 = about 2 GB per block), each block sorted per row and reduced.
 * Exact — identical to the dense entry on the same features. */
int mcbb_knn_dist_features(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                           const double* group_weight, int32_t k, int32_t K, int64_t block_rows, double* kth_out,
                           double* cum_out);
int mcbb_knn_dist_features_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, int32_t k, int32_t K, int64_t block_rows,
                               double* kth_dev, double* cum_dev, void* stream);

/* ---- histogram-mode features (distance_matrix(...; histograms = true), src/eval_clustering.jl:187-211) -------------
 * `values` is the slab of ONE per-dimension measure in the layout mcbb_solve_features writes it: [n_dim][n_mc].
 * mcbb_order_stats: out[i] = the ranks[i]-th smallest value (0-based) over all n_dim * n_mc values — what
 * _compute_hist_edges (:607-642) needs: rank 0 and count - 1 are the extrema, the two neighbours of a type-7 quantile give
 * the IQR; *n_nan receives the number of NaNs (they are not ranked).  The edges themselves (a Julia range) are built by the
 * host from those numbers.
 * mcbb_hist_rows: replaces _compute_hist_weights (:587-598) + ecdf_hist (:759-762): per run the Float32 histogram of its
 * n_dim values over the equidistant `edges` (host array; closed = :left, values outside dropped), with use_ecdf != 0 its
 * Float32 cumulative sum divided by its last entry.  rows_out[b][r] (n_edges - 1 rows of n_mc, the feature layout of
 * mcbb_distance_matrix / mcbb_dbscan_features) is the row of run order[r] (order NULL: r itself — sort!(sol, prob),
 * setup_mc_prob.jl:520-536, folded in); normalize != 0 applies (x - lo) / range to every value first (normalize(sol),
 * setup_mc_prob.jl:635-660).  Bit-identical to the reference arithmetic (no contraction). */
int mcbb_order_stats(const double* values, int64_t count, int32_t n_ranks, const int64_t* ranks, double* out, int64_t* n_nan);
int mcbb_order_stats_dev(const double* values_dev, int64_t count, int32_t n_ranks, const int64_t* ranks, double* out,
                         int64_t* n_nan, void* stream);
int mcbb_hist_rows(const double* values, int64_t n_mc, int32_t n_dim, const int64_t* order, int32_t normalize, double lo,
                   double range, const double* edges, int32_t n_edges, int32_t use_ecdf, double* rows_out);
int mcbb_hist_rows_dev(const double* values_dev, int64_t n_mc, int32_t n_dim, const int64_t* order_dev, int32_t normalize,
                       double lo, double range, const double* edges, int32_t n_edges, int32_t use_ecdf, double* rows_dev,
                       void* stream);

/* ---- sparse interchange ------------------------------------------------------------
 * Replace the NonzeroSparseMatrix branch of distance_matrix (src/eval_clustering.jl:541-550 builds it column by
 * column from a dense row; src/sparse_clustering.jl:37-110 is the type): only entries with d < sparse_threshold are
 * stored, as d - sparse_threshold, in compressed columns; unstored entries read as the threshold.  Here the columns
 * are produced from the feature matrix without ever materialising D:
 *   1  mcbb_distance_sparse_count   count_out[c] = stored entries of column c (the diagonal, d = 0, included)
 *      -- caller forms colptr (exclusive prefix sum, 0-based, n+1 entries) and allocates rowval / nzval
 *   2  mcbb_distance_sparse_fill    rowval (0-based, ascending within a column) and nzval = d - sparse_threshold
 * mcbb_dbscan_sparse replaces dbscan(::AbstractNonzeroSparseMatrix, eps, minpts) (src/sparse_clustering.jl:126-215):
 * same labels as the sequential algorithm on the matrix whose unstored entries equal `value`. */
int mcbb_distance_sparse_count(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, double sparse_threshold, int32_t* count_out);
/* rows [row0,row1) of D as contiguous n-vectors (rows_out[(i-row0)*n + j] = D[i,j]): the unit that
 * distance_matrix_mmap (src/eval_clustering.jl:317-406) appends to its file, so that D never has to fit in memory */
int mcbb_distance_matrix_rows(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                              const double* group_weight, int64_t row0, int64_t row1, double* rows_out);
int mcbb_distance_matrix_rows_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                                  const double* group_weight, int64_t row0, int64_t row1, double* rows_dev,
                                  void* stream);
int mcbb_distance_sparse_count_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                                   const double* group_weight, double sparse_threshold, int32_t* count_dev,
                                   void* stream);
int mcbb_distance_sparse_fill(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                              const double* group_weight, double sparse_threshold, const int64_t* colptr,
                              int32_t* rowval_out, double* nzval_out);
int mcbb_distance_sparse_fill_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                                  const double* group_weight, double sparse_threshold, const int64_t* colptr_dev,
                                  int64_t nnz, int32_t* rowval_dev, double* nzval_dev, void* stream);
int mcbb_dbscan_sparse(int64_t n, const int64_t* colptr, const int32_t* rowval, const double* nzval, double value,
                       double eps, int32_t minpts, int64_t* assignments, int64_t* seeds, int64_t* counts,
                       int64_t* n_clusters);
int mcbb_dbscan_sparse_dev(int64_t n, const int64_t* colptr_dev, const int32_t* rowval_dev, const double* nzval_dev,
                           double value, double eps, int32_t minpts, int64_t* assignments_dev, int64_t* seeds_dev,
                           int64_t* counts_dev, int64_t* n_clusters_host, void* stream);

/* ---- rank-sharded DBSCAN stages (one process per GPU; the collectives stay with the caller) --------------
 * The pair space shards by point rows; rank r owns rows [row0,row1).  All pointers are device pointers.
 *   1  mcbb_dbscan_rows_count_dev    neighbour counts (self included) of the owned rows against all n columns
 *      -- caller all-gathers the counts, derives core flags and the ascending core / non-core index lists,
 *         and gathers their feature columns (mcbb_gather_columns_dev)
 *   2  mcbb_dbscan_rows_union_dev    union-find over core-core edges (row, col > row) seen from the owned
 *                                    rows of the CORE matrix, into a full-length parent array (reset here)
 *      -- caller all-gathers the R parent arrays; mcbb_dbscan_merge_parents_dev folds each into the local one
 *   3  mcbb_dbscan_rows_border_dev   owned non-core rows x all core columns: smallest adjacent seed
 *      -- caller all-gathers, builds root_label[n] (seed index of the point's cluster, -1 = noise)
 *   4  mcbb_dbscan_labels_dev        seeds / assignments / counts from root labels (any rank)          */
int mcbb_dbscan_rows_count_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, double eps, int32_t minpts, int64_t row0,
                               int64_t row1, int32_t* count_dev /* row1-row0, exact below minpts and a lower
                               bound >= minpts above it: counting stops once core-ness is decided */,
                               void* stream);
/* Same counts, plus the first `minpts` neighbours found for every owned row (column indices, unordered):
 * nbr_dev[(i - row0) * minpts + s], s < min(nfill_dev[i - row0], minpts).  A row that ends below minpts has all its
 * neighbours in the list, so the border stage (3) can be replaced by a lookup of their seeds. */
int mcbb_dbscan_rows_count_lists_dev(const double* X_dev, int64_t n, int32_t n_groups, const int32_t* group_len,
                                     const double* group_weight, double eps, int32_t minpts, int64_t row0,
                                     int64_t row1, int32_t* count_dev, int32_t* nbr_dev, int32_t* nfill_dev,
                                     void* stream);
int mcbb_dbscan_rows_union_dev(const double* Xcore_dev, int64_t n_core, int32_t n_groups,
                               const int32_t* group_len, const double* group_weight, double eps, int64_t row0,
                               int64_t row1, int32_t* parent_dev /* n_core */, void* stream);
/* adds the edges seen from a further row block to a parent array that already holds some (no reset): lets a rank own
 * several blocks of the pair triangle, interleaved with the other ranks' */
int mcbb_dbscan_rows_union_more_dev(const double* Xcore_dev, int64_t n_core, int32_t n_groups,
                                    const int32_t* group_len, const double* group_weight, double eps, int64_t row0,
                                    int64_t row1, int32_t* parent_dev, void* stream);
int mcbb_dbscan_merge_parents_dev(int32_t* parent_dev, const int32_t* other_parent_dev, int64_t n_core,
                                  void* stream);
int mcbb_dbscan_rows_border_dev(const double* Xnon_dev, int64_t n_non, int64_t row0, int64_t row1,
                                const double* Xcore_dev, int64_t n_core, const int32_t* core_seed_dev,
                                int32_t n_groups, const int32_t* group_len, const double* group_weight,
                                double eps, int32_t* best_dev /* row1-row0; INT32_MAX = none */, void* stream);
int mcbb_dbscan_labels_dev(const int32_t* root_label_dev, int64_t n, int64_t* assignments_dev,
                           int64_t* seeds_dev, int64_t* counts_dev, int64_t* n_clusters_host, void* stream);
/* Xout[f][k] = X[f][idx[k]] for k < m: compaction of feature columns (idx 0-based) */
int mcbb_gather_columns_dev(const double* X_dev, int64_t ld, const int32_t* idx_dev, int64_t m, int32_t F,
                            double* Xout_dev, void* stream);

/* ---- several devices driven from ONE host process (SURVEY §8(b) "Threading"; docs/src/hpc.md:27-35 and the
 * EnsembleDistributed() default of solve, src/setup_mc_prob.jl:1098, are the reference's way to scale out) ---------------
 * mcbb_multi_init binds `n_devices` GPUs (devices == NULL: 0 .. n_devices-1), creates one NCCL communicator per device
 * inside the library (libnccl.so.2 is loaded with dlopen, no link-time dependency) and one stream per device.  The *_multi
 * entry points take the same HOST buffers as their single-device twins, run one host thread per device and return when
 * all devices are done.  One *_multi call at a time per process. */
int mcbb_multi_init(int32_t n_devices, const int32_t* devices);
int mcbb_multi_finalize(void);
int mcbb_multi_info(int32_t* n_devices, int32_t* nccl_version);
/* mcbb_solve_features over all devices: contiguous blocks of runs per device, no collective (runs are independent);
 * every device writes its block straight into the caller's arrays.  Same arguments and results as mcbb_solve_features. */
int mcbb_solve_features_multi(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                              int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                              double* feat_out, double* matrix_out, double* global_out, int32_t* retcode_out,
                              int64_t* nsteps_out);
/* mcbb_dbscan_features over all devices: X is copied to every device, the pair space is sharded by interleaved row
 * blocks, neighbour counts / union-find parents / border labels are exchanged with NCCL.  Labels are identical to
 * mcbb_dbscan_features (and so to Clustering.dbscan on the materialised matrix). */
int mcbb_dbscan_features_multi(const double* X, int64_t n, int32_t n_groups, const int32_t* group_len,
                               const double* group_weight, double eps, int32_t minpts, int64_t* assignments,
                               int64_t* seeds, int64_t* counts, int64_t* n_clusters);
/* solve -> sort -> distance -> dbscan in one call, the features never leave the devices: shard, integrate + featurize,
 * NCCL all-gather of the per-device feature blocks, STABLE sort of the runs by par[:,0] on the device (what sort!(sol,
 * prob) does on the host, src/setup_mc_prob.jl:520-536), sharded on-the-fly DBSCAN with the distance of
 * distance_matrix(sol, prob, weights) (per-dimension measures + parameter columns; `weights` has n_meas + n_par entries).
 *   feat_out / retcode_out / nsteps_out   in the CALLER's run order (any may be NULL)
 *   order_out[n_mc]                       0-based sortperm: sorted position k holds run order_out[k]
 *   assignments / seeds / counts          refer to the SORTED runs, as dbscan(distance_matrix(sol, prob, ...)) would */
int mcbb_solve_cluster_multi(const mcbb_system_desc* sys, const mcbb_solver_opts* opts, const mcbb_feat_opts* feat,
                             int64_t n_mc, const double* ic, const double* par, const double* t_save, int32_t n_t,
                             const double* weights, double eps, int32_t minpts, double* feat_out, int32_t* retcode_out,
                             int64_t* nsteps_out, int64_t* order_out, int64_t* assignments, int64_t* seeds,
                             int64_t* counts, int64_t* n_clusters);

#ifdef __cplusplus
}
#endif
#endif /* MCBB_B200_H */
