// batched_umma_sl.cu — ensemble Tsit5 for the Stuart-Landau ring of paper/stuartlandau/stuart_landau_sathiyadevi-
// standard-config.ipynb (cell 1; BASELINE config C4: N = 100 complex oscillators), coupling sums on the tensor cores.
//
//     dz_i = (lambda + i omega - |z_i|^2) z_i + eps/(2 P1) sum_{A1_ij} (Re z_j - Re z_i) - i eps/(2 P2) sum_{A2_ij} (Im z_j - Im z_i)
//
// The coupling is LINEAR in the state: sum_{A1}(x_j - x_i) = (A1 x)_i - deg1_i x_i, and A1 / A2 are 0/1 matrices
// (Bool adjacency of Watts-Strogatz graphs, degree 2 P1 = 2 and 2 P2 = 44).  So the two mat-vecs of every right-hand-side
// evaluation are exact fixed-point integer GEMMs, the same formulation as the Kuramoto kernel (batched_umma.cu,
// fixedpoint.cuh) without any transcendental: x in (-R, R) is written as u = round(x 2^50 / R) + 2^51 (52 bits), split
// into seven radix-256 digits, multiplied by the 0/1 matrix with tcgen05.mma kind::i8 (int32 accumulators, exact), and
// recombined in int64 — the sum carries ONE rounding instead of deg_i.
//
// Machine mapping (one CTA = 128 threads = one thread per oscillator, JB trajectories in lock step):
//   * tensor memory, 16 JB + 64 columns: D1 = A1 X-digits (8 JB columns), D2 = A2 Y-digits (8 JB columns), then the two
//     A operands (row r in lane r, four K bytes per 32-bit column) — both resident for the whole kernel.
//   * B operands in shared memory, MN-major: a 16-column chunk holds the 8 digit bytes of TWO trajectories, so the
//     thread of oscillator r publishes Re with one 8-byte store into the X panel and Im with one into the Y panel;
//     two MMA chains per evaluation (A1 x X panel -> D1, A2 x Y panel -> D2), one commit.
//   * a SPARSE A1 (every degree <= 8; the notebook's ring has 2 P1 = 2) is not sent through the tensor core at all: Re of
//     the stage value is staged in shared memory and each oscillator adds its neighbours' differences in ascending order
//     (template flag A1D: only D2 / the Y panel / A2 exist).  Measured equal in speed to the two-GEMM form — the kernel
//     is bound by per-warp latency, not by the tensor pipe — but it needs half the tensor memory and digit panels.
//   * every Runge-Kutta stage vector is thread-private shared memory [oscillator][trajectory][re | im][u, k1..k7];
//     sums come back with tcgen05.ld into the thread that owns the oscillator: no CTA barrier inside the stage loop.
//   * statistics of eval_ode_run fused: mean / std of Re and Im in registers (Welford); for the KL-histogram measures
//     and correlation_ecdf / correlation_hist (eval_mc_prob.jl:312-330, 462-488) the saved samples of a CHUNK of runs
//     go to a device scratch [run][n_t][2N] and are reduced after the chunk is integrated: k_sl_post (KL per oscillator
//     thread, one CTA per run) and k_sl_corr_post (the N x N complex correlation as 4 x 4 FP64 register tiles over the
//     upper triangle, samples tiled through shared memory).
//   * error control as DiffEqBase does it for a Complex{Float64} state: one scale abstol + max(|z0|, |z1|) reltol per
//     complex element (complex modulus), RMS over N elements (calculate_residuals / ODE_DEFAULT_NORM).
//
// A state component outside (-R, R) cannot be represented: R is chosen from lambda and the initial conditions with a
// factor 4 of head room, stage values are clamped, and a run that ever ACCEPTS a step whose stages were clamped is
// counted — the host then repeats the ensemble on the lane kernels (never observed: |z| stays near sqrt(lambda)).
#include <algorithm>
#include <cmath>
#include <vector>

#include "umma_common.cuh"

namespace mcbb {

// what the (out-of-line, once per run) finish needs: read from device memory, so that the call costs the hot loop no stack
struct SlFin {
    int N, n_t, n_meas, meas[MCBB_MAX_MEAS], comp[MCBB_MAX_MEAS], n_filter, kl_bins, matrix_meas, corr_nbins;
    long long n_mc;
    double kl_nstd, kl_tol;
    double* feat;
    double* matrix;
};

struct SlArgs {
    int N, KS, n_sms;
    double inv_n, ln_qoldinit;
    const unsigned char* a1_img;  // [128][128] row-major bytes
    const unsigned char* a2_img;
    const int* deg1;              // [128]
    const int* deg2;
    const unsigned char* nbr1;    // [128][8] neighbours of A1 in ascending order (A1D instantiation: sparse A1 summed directly)
    int nb1max;                   // largest degree of A1
    const int* filter_pos;        // [128] position of oscillator r in the state_filter, -1 = not reported
    double scalars[MCBB_N_SCALARS];
    int n_par;
    int par_slot[MCBB_MAX_PAR];
    SolverDev o;
    int n_meas, meas[MCBB_MAX_MEAS], comp[MCBB_MAX_MEAS], n_filter;
    SolveIO io;
    int need_kl, kl_bins;
    double kl_nstd, kl_tol;
    int matrix_meas, corr_nbins;
    double* smp;                  // [runs of the chunk][n_t][2N] saved samples (KL / correlation: reduced by k_sl_post)
    double* stat;                 // [runs of the chunk][N][4] mean / M2 of Re and Im as the run finished
    int* okflag;                  // [runs of the chunk] run succeeded and delivered all samples
    long long run_begin, run_end; // the runs this launch integrates (one chunk of the ensemble when samples are kept)
    double fx_scale, fx_inv, fx_max;  // 2^50 / R, R / 2^50, largest representable magnitude
    unsigned int* ovf;            // number of runs that accepted a clamped stage
    const SlFin* fin;             // device copy of the finish parameters
};

struct SlSlotCtl {
    long long run, fin_run, iter, nrej, natt;
    double t, dt, dts, lq, d0, d1, dt0, t_old, dts_old, hstep;
    double eP1, eP2, lam, om;  // eps / (2 P1), eps / (2 P2), lambda, omega of the run in this slot
    int mode, isave, ret, clipped, accept, fresh, fin, fin_ok, s0, s1, ovf;
};

// 8 columns (one trajectory's digit sums of one component): 32 lanes x 8 columns
__device__ __forceinline__ void sl_tmem_ld16(const unsigned taddr, unsigned (&v)[16]) { um_tmem_ld16(taddr, v); }

// |cor| histogram of the complex series of one finished run (Statistics.cor(x, dims = 2) of a Complex matrix: rows
// centred, x x' with the second factor conjugated, divided by the roots of the diagonal; eval_mc_prob.jl:462-481).
// All 128 threads of the CTA.  The run's saved samples are first centred IN PLACE in the device scratch (they are dead
// afterwards); then every thread accumulates 4 x 4 tiles of the upper triangle (pairs i < j) straight from the scratch
// — each sample row is 2N doubles shared by all threads, so the loads hit L1 — without any CTA barrier in the sample loop.
// hist receives 2 per pair ((i,j) and (j,i)).
static __device__ __noinline__ void sl_corr_hist(double* __restrict__ smp, const int n_t, const int N, const double* s_mx,
                                                 const double* s_my, const double* s_var, int* hist, const int nb) {
    const int tid = threadIdx.x;
    const int NBk = (N + 3) >> 2;
    const int n_blocks = NBk * (NBk + 1) / 2;
    const int n2 = 2 * N;
    for (int b = tid; b < nb; b += 128) hist[b] = 0;
    // the samples were written long ago and have left the L2 (the scratch of all slots is ~1 GB): keep 16 independent
    // loads in flight per thread, or the loop is bound by one HBM round trip per row
    for (int d = tid; d < n2; d += 128) {
        const double mu = (d & 1) ? s_my[d >> 1] : s_mx[d >> 1];
        for (int k0 = 1; k0 < n_t; k0 += 16) {
            double v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] = (k0 + q < n_t) ? __ldcs(smp + (size_t)(k0 + q) * n2 + d) : 0.;
#pragma unroll
            for (int q = 0; q < 16; q++)
                if (k0 + q < n_t) smp[(size_t)(k0 + q) * n2 + d] = v[q] - mu;
        }
    }
    __syncthreads();  // centred samples (global) and the zeroed histogram (shared) visible to the whole CTA
    const double step = 1. / (double)nb;
    for (int id = tid; id < n_blocks; id += 128) {
        int bi = 0, rem = id;  // id -> (bi, bj), bi <= bj, rows enumerated bi = 0, 1, ... with NBk - bi entries each
        while (rem >= NBk - bi) {
            rem -= NBk - bi;
            bi++;
        }
        const int bj = bi + rem;
        // the last tile row / column may reach past N: clamp the loads to valid rows, the pairs are skipped below
        const int ia = min(4 * bi, N - 4 < 0 ? 0 : N - 4), ic = min(4 * bj, N - 4 < 0 ? 0 : N - 4);
        const int sa = 4 * bi - ia, sc = 4 * bj - ic;  // shift of the tile inside the loaded window (0 unless clamped)
        double re[4][4], im[4][4];
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = 0; q < 4; q++) re[p][q] = im[p][q] = 0.;
#pragma unroll 4
        for (int k = 1; k < n_t; k++) {
            const double2* ra = reinterpret_cast<const double2*>(smp + (size_t)k * n2 + 2 * ia);
            const double2* rc = reinterpret_cast<const double2*>(smp + (size_t)k * n2 + 2 * ic);
            double2 za[4], zc[4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                za[p] = ra[p];
                zc[p] = rc[p];
            }
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) {  // (a + ib)(c - id)
                    re[p][q] = fma(za[p].x, zc[q].x, re[p][q]);
                    re[p][q] = fma(za[p].y, zc[q].y, re[p][q]);
                    im[p][q] = fma(za[p].y, zc[q].x, im[p][q]);
                    im[p][q] = fma(-za[p].x, zc[q].y, im[p][q]);
                }
        }
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = ia + p, j = ic + q;
                if (p < sa || q < sc) continue;  // rows of a clamped window that belong to the previous tile
                if (i >= j || j >= N) continue;
                const double den = sqrt(s_var[i]) * sqrt(s_var[j]);
                const double c = hypot(re[p][q] / den, im[p][q] / den);
                if (c != c) continue;  // zero-variance oscillator: dropped, as in the lane kernels and the oracle
                double f = c / step + 1.;  // Base.searchsortedlast on the range 0:1/nbins:1
                f = fmin(fmax(f, 1.), (double)(nb + 1));
                const int n0 = (int)rint(f);
                const int idx = (c < (double)(n0 - 1) / (double)nb) ? n0 - 1 : n0;
                if (idx >= 1 && idx <= nb) atomicAdd(&hist[idx - 1], 2);
            }
    }
    __syncthreads();
}

// results of one finished run: the per-oscillator measures (every thread its own oscillator), the KL-hist measures from the
// run's saved samples, the correlation histogram by the whole CTA.  `scratch` = the digit panels, dead between two rounds.
template <bool SMP>
static __device__ __noinline__ void sl_finish_run(const SlFin* __restrict__ fp, double* smp_run, const bool ok, const long long run,
                                                  const int fpos, const double amx, const double a2x, const double amy,
                                                  const double a2y, unsigned char* scratch, const int scratch_bytes,
                                                  double* sm_mx, double* sm_my, double* sm_var, int* sm_hist) {
    const SlFin a = *fp;
    const int tid = threadIdx.x, N = a.N, n_t = a.n_t;
    const long long n_mc = a.n_mc;
    if (fpos >= 0) {
        // std(u; mean, corrected = true) = sqrt(sum (u - mean)^2 / (n - 1)), n = n_t - 1 samples
        const double inv = 1. / (double)(n_t - 2);
        const double mean_c[2] = {ok ? amx : nan(""), ok ? amy : nan("")};
        const double sd_c[2] = {ok ? sqrt((a2x < 0. ? 0. : a2x) * inv) : nan(""), ok ? sqrt((a2y < 0. ? 0. : a2y) * inv) : nan("")};
        for (int m = 0; m < a.n_meas; m++) {
            const int cp = a.comp[m];
            double v;
            if (a.meas[m] == MCBB_MEAS_MEAN) {
                v = mean_c[cp];
            } else if (a.meas[m] == MCBB_MEAS_STD) {
                v = sd_c[cp];
            } else {
                v = nan("");
                if (SMP && ok)
                    v = um_kl_hist(smp_run + 2 * tid + cp, 2 * N, n_t, mean_c[cp], sd_c[cp], a.kl_bins, a.kl_nstd, a.kl_tol,
                                   (scratch_bytes >= a.kl_bins * 256) ? reinterpret_cast<unsigned short*>(scratch) + tid : nullptr);
            }
            a.feat[((long long)m * a.n_filter + fpos) * n_mc + run] = v;
        }
    }
    if (SMP && a.matrix_meas != MCBB_MAT_NONE) {
        const int nbn = a.corr_nbins;
        if (!ok) {
            for (int b = tid; b < nbn; b += 128) a.matrix[(long long)b * n_mc + run] = nan("");
        } else {
            // correlation runs over the FILTERED oscillators in filter order; only the identity filter is dispatched here,
            // so oscillator index = position
            sm_mx[tid] = amx;
            sm_my[tid] = amy;
            sm_var[tid] = a2x + a2y;
            __syncthreads();
            sl_corr_hist(smp_run, n_t, N, sm_mx, sm_my, sm_var, sm_hist, nbn);
            if (tid == 0) {
                double tot = 0., cum = 0.;
                for (int b = 0; b < nbn; b++) tot += (double)sm_hist[b];
                for (int b = 0; b < nbn; b++) {
                    cum += (double)sm_hist[b];
                    a.matrix[(long long)b * n_mc + run] = (a.matrix_meas == MCBB_MAT_CORR_HIST) ? (double)sm_hist[b] : cum / tot;
                }
            }
        }
    }
    if (SMP) __syncthreads();  // the digit panels served as scratch: all done before they are rewritten
}

// SMP = true: the instantiation that keeps the saved samples (KL-hist measures and / or a correlation matrix measure).
// A1D = true: A1 has at most 8 neighbours per oscillator (the notebook's ring: 2 P1 = 2), so its coupling sum is taken
// directly — Re of every stage value is staged in shared memory (double-buffered over the stages) and every oscillator
// adds its neighbours' differences in ascending order, the very expression of the right-hand side — and only the dense
// A2 (2 P2 = 44) goes through the tensor core: half the digit panels, half the MMAs, half the tensor-memory read-back.
template <int JB, int MINB, bool SMP, bool A1D>
__global__ void __launch_bounds__(128, MINB) k_solve_batched_umma_sl(const SlArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int N = a.N;
    const int row = tid;
    const bool valid = row < N;
    const int rr = valid ? row : N;  // idle lanes work on a dummy row: no branches in the stage loop
    const SolverDev& o = a.o;
    const int n_t = a.io.n_t;
    const long long n_mc = a.io.n_mc;
    constexpr unsigned NH = 8 * JB;   // columns of each accumulator block (D1, D2): 8 per trajectory (7 digits, 1 zero)
    constexpr int RS = 2 * UM_NSLOT * JB + 2;  // row stride (doubles) = 2 (mod 4): conflict-free 128-bit accesses
    static_assert(JB % 2 == 0, "two trajectories share a 16-column chunk of the digit panels");

    constexpr int PANEL = A1D ? (JB / 2) * 2048 : JB * 2048;    // digit panels
    constexpr int XST = A1D ? 2 * JB * 128 * 8 : 0;             // Re of the stage values, [2][JB][128] doubles
    unsigned char* sBx = smem;                                  // [JB/2][128][16]  digits of Re, two trajectories per chunk (not A1D)
    unsigned char* sBy = smem + (A1D ? 0 : (JB / 2) * 2048);    // [JB/2][128][16]  digits of Im
    double* sm_xst = (double*)(smem + PANEL);
    double* sm_k = (double*)(smem + PANEL + XST);               // [N + 1][JB][2][8] (+2 pad)
    const double* const sm_ts = c_um_ts;                        // [n_t] (constant bank)
    double* sm_red0 = sm_k + (size_t)(N + 1) * RS;              // [4][JB]
    double* sm_red1 = sm_red0 + 4 * JB;                         // [4][JB]
    double* sm_mx = sm_red1 + 4 * JB;                           // [128] means / variance of a finishing run (correlation)
    double* sm_my = sm_mx + 128;
    double* sm_var = sm_my + 128;
    SlSlotCtl* ctl = (SlSlotCtl*)(sm_var + 128);                // [JB]
    unsigned long long* mbar_p = (unsigned long long*)(ctl + JB);
    unsigned long long* full_p = mbar_p + 1;
    unsigned* tmem_slot = (unsigned*)(mbar_p + 2);
    int* sm_flag = (int*)(tmem_slot + 2);                       // [4][JB] clamp flags of the current step, per warp
    int* sm_hist = sm_flag + 4 * JB;                            // [64] correlation histogram
    const unsigned mbar_full = um_smem_u32(full_p);
    const unsigned mbar = um_smem_u32(mbar_p);
    double* const kb = sm_k + (size_t)rr * RS;

#define KX(slot, j) kb[(2 * (j)) * UM_NSLOT + (slot)]
#define KY(slot, j) kb[(2 * (j) + 1) * UM_NSLOT + (slot)]

    for (int i = tid; i < (PANEL + XST) / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < (N + 1) * RS; i += 128) sm_k[i] = 0.;
    const int cw = (blockIdx.x / a.n_sms) & 3;  // the controller lanes sit in a different warp for each resident CTA
    const bool is_ctl = (wid == cw) && lane < JB;
    if (is_ctl) {
        SlSlotCtl& c = ctl[lane];
        c.mode = MODE_FETCH;
        c.accept = c.fresh = c.fin = 0;
        c.hstep = 0.;
        c.dts = 0.;
        c.eP1 = c.eP2 = c.lam = c.om = 0.;
        c.ovf = 0;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(mbar_full));
    }
    constexpr unsigned TUSE = A1D ? (8 * JB + 32) : (16 * JB + 64);
    constexpr unsigned TCOLS = TUSE <= 32 ? 32u : TUSE <= 64 ? 64u : TUSE <= 128 ? 128u : TUSE <= 256 ? 256u : 512u;
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(tmem_slot)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const unsigned tmem_row = tmem_base + ((unsigned)(32 * wid) << 16);
    constexpr unsigned TD2 = A1D ? 0u : NH, TA = A1D ? NH : 2 * NH;  // column offsets of D2 and of the A operands
    const unsigned tmem_d1 = tmem_base, tmem_d2 = tmem_base + TD2;
    const unsigned tmem_a1 = tmem_base + TA, tmem_a2 = tmem_base + TA + (A1D ? 0u : 32u);
    {  // rows of A1, A2 -> lane r: 128 K bytes = 32 columns each
        unsigned w[32];
#pragma unroll 1
        for (int which = A1D ? 1 : 0; which < 2; which++) {
            const unsigned char* img = which ? a.a2_img : a.a1_img;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint4 t4 = reinterpret_cast<const uint4*>(img + (size_t)row * 128)[q];
                w[4 * q] = t4.x;
                w[4 * q + 1] = t4.y;
                w[4 * q + 2] = t4.z;
                w[4 * q + 3] = t4.w;
            }
            um_tmem_st32(tmem_row + TA + (A1D ? 0u : 32u * which), w);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // D s32, A u8 K-major (tensor memory), B u8 MN-major, N = 8 JB, M = 128
    const unsigned idesc = (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (1u << 16) | ((NH >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned sBx_addr = um_smem_u32(sBx), sBy_addr = um_smem_u32(sBy);
    const int KSTEPS = a.KS;
    unsigned char* const sBx_row = sBx + row * 16;
    unsigned char* const sBy_row = sBy + row * 16;

    const int degh1 = a.deg1[row] << 19, degh2 = a.deg2[row] << 19;  // deg 2^51 in units of the high word
    const double dg1 = (double)a.deg1[row], dg2 = (double)a.deg2[row];
    const int fpos = a.filter_pos[row];
    // A1D: the (at most 8) neighbours of this oscillator in A1, ascending, one byte each
    const uint2 nbw = A1D ? reinterpret_cast<const uint2*>(a.nbr1)[row] : make_uint2(0u, 0u);
    const int nb1 = A1D ? a.deg1[row] : 0, nb1max = A1D ? a.nb1max : 0;
    const double fxs = a.fx_scale, fxi = a.fx_inv;
    const int fx_hi_max = valid ? __double2hiint(a.fx_max) : 0x7fffffff;  // rows >= N never raise the flag
    unsigned phase = 0;

    // thread-private state of the JB elements (complex) of this oscillator
    double xs[JB], ys[JB], amx[JB], a2x[JB], amy[JB], a2y[JB];
#pragma unroll
    for (int j = 0; j < JB; j++) xs[j] = ys[j] = amx[j] = a2x[j] = amy[j] = a2y[j] = 0.;
    unsigned clampmask = 0u;  // bit j: a stage value of trajectory j was clamped during the current step

    auto begin_attempt = [&](SlSlotCtl& c) -> bool {
        if (++c.iter > o.maxiters) {
            c.ret = MCBB_RET_MAXITERS;
            return false;
        }
        c.dts = c.dt;
        c.clipped = 0;
        if (c.dts >= o.t1 - c.t) {
            c.dts = o.t1 - c.t;
            c.clipped = 1;
        }
        if (!c.clipped && c.dts <= o.dtmin) {
            c.ret = MCBB_RET_DTLESSTHANMIN;
            return false;
        }
        c.hstep = c.dts;
        return true;
    };

    for (;;) {
        // ---- per-trajectory control: PI controller, accept / reject, save window, finish, refill -----------------
        if (is_ctl) {
            SlSlotCtl& c = ctl[lane];
            c.accept = 0;
            c.fin = 0;
            c.fresh = 0;
            int finished = 0;
            if (c.mode == MODE_INIT) {
                c.mode = MODE_RUN;
                if (!begin_attempt(c)) finished = 1;
            } else if (c.mode == MODE_RUN) {
                c.natt++;
                double e2 = 0., nb = 0.;
                int clamped = 0;
                for (int w = 0; w < 4; w++) {
                    e2 += sm_red0[w * JB + lane];
                    nb += sm_red1[w * JB + lane];
                    clamped |= sm_flag[w * JB + lane];
                }
                // PI controller as in batched_umma.cu: dt * clamp(gamma exp(beta2 ln qold - beta1 ln EEst))
                const double m2 = e2 * a.inv_n;  // EEst^2 over N complex elements
                const double L = 0.5 * log(m2);
                if (!(m2 > 1.)) {
                    const double r = (m2 == 0.) ? o.qmax : fmin(o.qmax, fmax(o.qmin, o.gamma * exp(fma(o.beta2, c.lq, -o.beta1 * L))));
                    c.lq = fmax(L, a.ln_qoldinit);
                    const double tnew = c.clipped ? o.t1 : c.t + c.dts;
                    int k1 = c.isave;
                    while (k1 < n_t && sm_ts[k1] <= tnew) k1++;
                    c.s0 = c.isave;
                    c.s1 = k1;
                    c.isave = k1;
                    c.t_old = c.t;
                    c.dts_old = c.dts;
                    c.t = tnew;
                    c.dt = fmin(o.dtmax, fmax(o.dtmin, c.dts * r));
                    c.accept = 1;
                    c.ovf |= clamped;
                    if (nb > 0.) c.ret = MCBB_RET_UNSTABLE;  // NaN in the new state
                } else {
                    c.nrej++;
                    c.dt = c.dts * fmax(o.qmin, o.gamma * exp(-o.beta1 * L));
                }
                if (c.ret != 0 || (c.accept && !(c.t < o.t1)))
                    finished = 1;
                else if (!begin_attempt(c))
                    finished = 1;
            }
            if (finished) {
                c.fin = 1;
                c.fin_run = c.run;
                c.fin_ok = (c.ret == 0) && c.isave >= n_t;
                if (a.io.retcode) a.io.retcode[c.run] = c.ret;
                if (a.io.nsteps) {
                    a.io.nsteps[c.run] = c.natt;
                    a.io.nsteps[n_mc + c.run] = c.nrej;
                }
                if (c.ovf) atomicAdd(a.ovf, 1u);
                c.mode = MODE_FETCH;
            }
            if (c.mode == MODE_FETCH) {
                const long long run = (long long)atomicAdd(a.io.work_counter, 1ULL);
                if (run >= a.run_end) {
                    c.mode = MODE_DONE;
                } else {
                    double sc[MCBB_N_SCALARS];
                    for (int q = 0; q < MCBB_N_SCALARS; q++) sc[q] = a.scalars[q];
                    for (int p = 0; p < a.n_par; p++) sc[a.par_slot[p]] = a.io.par[(long long)p * n_mc + run];
                    c.eP1 = sc[0] / (2. * sc[3]);  // epsP1 = p.eps / (2 p.P_1), notebook cell 1
                    c.eP2 = sc[0] / (2. * sc[4]);
                    c.lam = sc[1];
                    c.om = sc[2];
                    c.run = run;
                    c.t = o.t0;
                    c.iter = 0;
                    c.nrej = 0;
                    c.natt = 0;
                    c.lq = a.ln_qoldinit;
                    c.hstep = 0.;
                    c.isave = 0;
                    c.ret = 0;
                    c.ovf = 0;
                    c.mode = MODE_INIT;
                    c.fresh = 1;
                }
            }
        }
        const bool all_done = __syncthreads_and(!is_ctl || ctl[lane].mode == MODE_DONE);

        // ---- oscillator owners: samples of the accepted step, z <- z_new, results of finished runs, new state ----
        bool any_init = false;
        unsigned runmask = 0u;
#pragma unroll
        for (int j = 0; j < JB; j++) {
            any_init |= (ctl[j].mode == MODE_INIT);
            runmask |= (ctl[j].mode == MODE_RUN ? 1u : 0u) << j;
        }
#pragma unroll
        for (int j = 0; j < JB; j++) {
            const SlSlotCtl& c = ctl[j];
            if (c.accept) {
                const int ks0 = c.s0 < 1 ? 1 : c.s0, ks1 = c.s1;  // sample 0 is dropped (eval_mc_prob.jl:152)
                const double dts = c.dts_old, t_old = c.t_old;
                double kx[8], ky[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2 t2 = reinterpret_cast<const double2*>(kb + (2 * j) * UM_NSLOT)[q];
                    const double2 t3 = reinterpret_cast<const double2*>(kb + (2 * j + 1) * UM_NSLOT)[q];
                    kx[2 * q] = t2.x;
                    kx[2 * q + 1] = t2.y;
                    ky[2 * q] = t3.x;
                    ky[2 * q + 1] = t3.y;
                }
                if (ks0 < ks1) {
                    const int lq = lane < 7 ? lane : 0;
                    const double pa = c_um_bth[lq][0], pb = c_um_bth[lq][1], pc = c_um_bth[lq][2];
                    for (int k = ks0; k < ks1; k++) {
                        // dense-output weights: lane q evaluates b_q(theta), a shuffle hands all seven to every row
                        const double th = um_div(sm_ts[k] - t_old, dts);
                        const double inner = fma(th, fma(th, pc, pb), pa);
                        const double bq = (lq == 0) ? th * fma(th, inner, 1.) : (th * th) * inner;
                        const double b0 = __shfl_sync(0xffffffffu, bq, 0);
                        double lx = kx[1] * b0, ly = ky[1] * b0;
#pragma unroll
                        for (int q = 1; q < 7; q++) {
                            const double bb = __shfl_sync(0xffffffffu, bq, q);
                            lx = fma(kx[1 + q], bb, lx);
                            ly = fma(ky[1 + q], bb, ly);
                        }
                        const double x = fma(dts, lx, kx[0]), y = fma(dts, ly, ky[0]);
                        if (SMP && valid)
                            // (a run that finished with this step has already handed its slot to the next run: fin_run)
                            *reinterpret_cast<double2*>(a.smp + (((size_t)((c.fin ? c.fin_run : c.run) - a.run_begin)) * n_t + k) * (2 * N) + 2 * row) =
                                make_double2(x, y);
                        const double rk = um_div(1., (double)k);
                        const double dx = x - amx[j], dy = y - amy[j];
                        if (fabs(dx) <= 1.7976931348623157e308) {
                            const double mnew = fma(dx, rk, amx[j]);
                            a2x[j] = fma(dx, x - mnew, a2x[j]);
                            amx[j] = mnew;
                        } else {
                            amx[j] = nonfinite_mean(amx[j], x);
                            a2x[j] = nan("");
                        }
                        if (fabs(dy) <= 1.7976931348623157e308) {
                            const double mnew = fma(dy, rk, amy[j]);
                            a2y[j] = fma(dy, y - mnew, a2y[j]);
                            amy[j] = mnew;
                        } else {
                            amy[j] = nonfinite_mean(amy[j], y);
                            a2y[j] = nan("");
                        }
                    }
                }
                // the accepted state: the expression the error estimate evaluated, recomputed
                double lux = TS_A(5, 0) * kx[1], luy = TS_A(5, 0) * ky[1];
#pragma unroll
                for (int q = 1; q < 6; q++) {
                    lux = fma(TS_A(5, q), kx[1 + q], lux);
                    luy = fma(TS_A(5, q), ky[1 + q], luy);
                }
                KX(0, j) = fma(dts, lux, kx[0]);
                KY(0, j) = fma(dts, luy, ky[0]);
                KX(1, j) = kx[7];  // FSAL
                KY(1, j) = ky[7];
            }
            if (c.fin) {  // uniform over the CTA
                if (SMP) {
                    // the KL histograms and the correlation need the whole series: the moments go to a scratch and
                    // k_sl_post reduces the saved samples of every run of the chunk afterwards, one CTA per run — inside
                    // this kernel the reduction stalled the three other trajectories of the CTA for its whole length
                    const size_t rl = (size_t)(c.fin_run - a.run_begin);
                    if (valid) *reinterpret_cast<double4*>(a.stat + (rl * N + row) * 4) = make_double4(amx[j], a2x[j], amy[j], a2y[j]);
                    if (tid == 0) a.okflag[rl] = c.fin_ok;
                } else {
                    sl_finish_run<false>(a.fin, nullptr, c.fin_ok != 0, c.fin_run, valid ? fpos : -1, amx[j], a2x[j], amy[j], a2y[j],
                                         smem, PANEL + XST, sm_mx, sm_my, sm_var, sm_hist);
                }
            }
            if (c.fresh) {
                if (valid) {
                    KX(0, j) = a.io.ic[(long long)(2 * row) * n_mc + c.run];
                    KY(0, j) = a.io.ic[(long long)(2 * row + 1) * n_mc + c.run];
#pragma unroll
                    for (int q = 1; q < UM_NSLOT; q++) KX(q, j) = KY(q, j) = 0.;
                }
                amx[j] = a2x[j] = amy[j] = a2y[j] = 0.;
            }
        }
        if (all_done) break;
        clampmask = 0u;  // (the digit panels and the Re staging, lent as scratch above, are rewritten by stage (a) of every round)

        // ---- one round: 6 right-hand-side evaluations of the whole batch (branch-free over the trajectories) --------
#pragma unroll 1
        for (int s = 0; s < 6; s++) {
            double cs[6];
#pragma unroll
            for (int q = 0; q < 6; q++) cs[q] = (q <= s) ? TS_A(s, q) : 0.;
            const double ci0 = (s == 1) ? 1. : 0.;  // initial-dt probe: u0 + dt0 f(u0) in the second evaluation
            // (a) stage input z -> fixed-point digits of Re and Im
#pragma unroll
            for (int j = 0; j < JB; j++) {
                const double2* qx = reinterpret_cast<const double2*>(kb + (2 * j) * UM_NSLOT);
                const double2* qy = reinterpret_cast<const double2*>(kb + (2 * j + 1) * UM_NSLOT);
                const double2 x01 = qx[0], x23 = qx[1], x45 = qx[2];
                const double x6 = qx[3].x;
                const double2 y01 = qy[0], y23 = qy[1], y45 = qy[2];
                const double y6 = qy[3].x;
                const double c0 = ((runmask >> j) & 1u) ? cs[0] : ci0;
                double lx = c0 * x01.y, ly = c0 * y01.y;
                lx = fma(cs[1], x23.x, lx);
                ly = fma(cs[1], y23.x, ly);
                lx = fma(cs[2], x23.y, lx);
                ly = fma(cs[2], y23.y, ly);
                lx = fma(cs[3], x45.x, lx);
                ly = fma(cs[3], y45.x, ly);
                lx = fma(cs[4], x45.y, lx);
                ly = fma(cs[4], y45.y, ly);
                lx = fma(cs[5], x6, lx);
                ly = fma(cs[5], y6, ly);
                const double h = ctl[j].hstep;
                const double x = fma(h, lx, x01.x), y = fma(h, ly, y01.x);
                xs[j] = x;
                ys[j] = y;
                // out of the fixed-point range (or NaN)?  One integer compare on the exponent words; the digits of such a
                // value are meaningless, which is harmless: a step that is ACCEPTED with the flag set invalidates the launch
                // (the ensemble is repeated on the lane kernels), a rejected one is simply retried with a smaller step
                const int hmx = A1D ? (UM_HI(y) & 0x7fffffff) : max(UM_HI(x) & 0x7fffffff, UM_HI(y) & 0x7fffffff);
                clampmask |= (hmx >= fx_hi_max ? 1u : 0u) << j;
                if (A1D) {
                    sm_xst[((s & 1) * JB + j) * 128 + row] = x;  // read by the neighbours after the stage barrier
                } else {
                    uint2 dx;
                    const double t = fma(x, fxs, UM_C(1));
                    dx.x = (unsigned)UM_LO(t);
                    dx.y = (unsigned)UM_HI(t) & 0xFFFFFu;
                    *reinterpret_cast<uint2*>(sBx_row + (j >> 1) * 2048 + (j & 1) * 8) = dx;  // rows >= N meet zero columns of A
                }
                uint2 dy;
                const double t2 = fma(y, fxs, UM_C(1));
                dy.x = (unsigned)UM_LO(t2);
                dy.y = (unsigned)UM_HI(t2) & 0xFFFFFu;
                *reinterpret_cast<uint2*>(sBy_row + (j >> 1) * 2048 + (j & 1) * 8) = dy;
            }
            // (b) D1 = A1 X, D2 = A2 Y on the tensor core; publish / issue / wait exactly as in batched_umma.cu
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            int issuer = 0;
            if (lane == 0) {
                unsigned long long st;
                unsigned pend;
                asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(mbar_full) : "memory");
                asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pend) : "l"(st));
                if (pend == 1u) {
                    issuer = 1;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t}" ::"r"(mbar_full),
                                 "r"(phase)
                                 : "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (!A1D)
                        for (int ks = 0; ks < KSTEPS; ks++)
                            um_mma_i8_ts(tmem_d1, tmem_a1 + 8u * ks, um_desc(sBx_addr + ks * 512, 128, 2048), idesc, ks > 0 ? 1u : 0u);
                    for (int ks = 0; ks < KSTEPS; ks++)
                        um_mma_i8_ts(tmem_d2, tmem_a2 + 8u * ks, um_desc(sBy_addr + ks * 512, 128, 2048), idesc, ks > 0 ? 1u : 0u);
                    um_commit(mbar);
                }
            }
            issuer = __shfl_sync(0xffffffffu, issuer, 0);
            if (issuer) {
                um_mbar_wait(mbar, phase);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // (c) row r of A1 X and A2 Y from tensor memory -> k = f(z) + coupling
            double kvx[JB], kvy[JB];
            double* const kst = kb + 2 + s;
            unsigned v1[A1D ? 1 : JB / 2][16], v2[JB / 2][16];  // all loads in flight, one wait
#pragma unroll
            for (int g = 0; g < JB / 2; g++) {
                if (!A1D) um_tmem_ld16_nowait(tmem_row + 16u * g, v1[g]);  // D1 columns of trajectories 2g, 2g + 1
                um_tmem_ld16_nowait(tmem_row + TD2 + 16u * g, v2[g]);      // D2
            }
            double at[JB];
            if (A1D) {  // sum_{A1} (x_j - x_i), neighbours in ascending order, while the tensor-memory loads are in flight
                const double* const xst = sm_xst + (s & 1) * JB * 128;
#pragma unroll
                for (int j = 0; j < JB; j++) at[j] = 0.;
                for (int q = 0; q < nb1max; q++) {
                    const unsigned nbq = ((q < 4 ? nbw.x : nbw.y) >> (8 * (q & 3))) & 0xffu;
                    if (q < nb1) {
#pragma unroll
                        for (int j = 0; j < JB; j++) at[j] += xst[j * 128 + nbq] - xs[j];
                    }
                }
            }
            um_tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < JB / 2; g++) {
                if (!A1D) um_tmem_landed(v1[g]);
                um_tmem_landed(v2[g]);
            }
#pragma unroll
            for (int j = 0; j < JB; j++) {
                const unsigned* w2 = v2[j >> 1] + 8 * (j & 1);
                const double sy = um_recombine(w2[0], w2[1], w2[2], w2[3], w2[4], w2[5], w2[6], degh2) * fxi;  // (A2 y)_r
                const SlSlotCtl& c = ctl[j];
                const double x = xs[j], y = ys[j];
                double attr;  // epsP1 sum_{A1} (x_j - x_i)
                if (A1D) {
                    attr = at[j] * c.eP1;
                } else {
                    const unsigned* w1 = v1[j >> 1] + 8 * (j & 1);
                    const double sx = um_recombine(w1[0], w1[1], w1[2], w1[3], w1[4], w1[5], w1[6], degh1) * fxi;  // (A1 x)_r
                    attr = fma(-dg1, x, sx) * c.eP1;
                }
                const double rep = fma(-dg2, y, sy) * c.eP2;
                const double cr = c.lam - fma(x, x, y * y);     // lambda - |z|^2
                kvx[j] = fma(cr, x, -c.om * y) + attr;
                kvy[j] = fma(cr, y, c.om * x) - rep;
                const bool run = (runmask >> j) & 1u;
                kst[(2 * j) * UM_NSLOT] = run ? kvx[j] : 0.;  // slots that do not integrate keep zeros
                kst[(2 * j + 1) * UM_NSLOT] = run ? kvy[j] : 0.;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (any_init && s < 2) {  // initial step size (OrdinaryDiffEq initdt.jl), complex element-wise norm
#pragma unroll
                for (int j = 0; j < JB; j++) {
                    if (ctl[j].mode == MODE_INIT) {
                        double p0 = 0., p1 = 0.;
                        if (valid) {
                            const double u0x = KX(0, j), u0y = KY(0, j);
                            const double sk = o.abstol + sqrt(fma(u0x, u0x, u0y * u0y)) * o.reltol;
                            if (s == 0) {
                                KX(1, j) = kvx[j];
                                KY(1, j) = kvy[j];
                                const double wx = u0x / sk, wy = u0y / sk, vx = kvx[j] / sk, vy = kvy[j] / sk;
                                p0 = fma(wx, wx, wy * wy);
                                p1 = fma(vx, vx, vy * vy);
                            } else {
                                const double vx = (kvx[j] - KX(1, j)) / sk, vy = (kvy[j] - KY(1, j)) / sk;
                                p0 = fma(vx, vx, vy * vy);
                            }
                        }
                        p0 = um_warp_sum(p0);
                        p1 = um_warp_sum(p1);
                        if (lane == 0) {
                            sm_red0[wid * JB + j] = p0;
                            sm_red1[wid * JB + j] = p1;
                        }
                    }
                }
                __syncthreads();
                if (is_ctl && ctl[lane].mode == MODE_INIT) {
                    SlSlotCtl& c = ctl[lane];
                    double r0 = 0., r1 = 0.;
                    for (int w = 0; w < 4; w++) {
                        r0 += sm_red0[w * JB + lane];
                        r1 += sm_red1[w * JB + lane];
                    }
                    if (o.dt_user > 0) {
                        c.dt = o.dt_user;
                        c.dt0 = 0.;
                    } else if (s == 0) {
                        c.d0 = sqrt(r0 / N);
                        c.d1 = sqrt(r1 / N);
                        const double dt0 = (c.d0 < 1e-5 || c.d1 < 1e-5) ? 1e-6 : 0.01 * (c.d0 / c.d1);
                        c.dt0 = fmin(dt0, o.dtmax);
                    } else {
                        const double d2 = sqrt(r0 / N) / c.dt0;
                        const double dm = fmax(c.d1, d2);
                        const double dt1 = (dm <= 1e-15) ? fmax(1e-6, c.dt0 * 1e-3) : pow(10., -(2. + log10(dm)) / 5.);
                        c.dt = fmin(fmin(100. * c.dt0, dt1), o.dtmax);
                    }
                    c.hstep = c.dt0;
                }
                __syncthreads();
            }
        }

        // ---- error estimate partials: one scale per complex element (calculate_residuals with the complex modulus) ----
        {
            double pe[JB];
            bool bad[JB];
#pragma unroll
            for (int j = 0; j < JB; j++) {
                const double dts = ctl[j].dts;
                double kx[8], ky[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2 t2 = reinterpret_cast<const double2*>(kb + (2 * j) * UM_NSLOT)[q];
                    const double2 t3 = reinterpret_cast<const double2*>(kb + (2 * j + 1) * UM_NSLOT)[q];
                    kx[2 * q] = t2.x;
                    kx[2 * q + 1] = t2.y;
                    ky[2 * q] = t3.x;
                    ky[2 * q + 1] = t3.y;
                }
                double ex = TS_BT(0) * kx[1], ey = TS_BT(0) * ky[1];
                double lux = TS_A(5, 0) * kx[1], luy = TS_A(5, 0) * ky[1];
#pragma unroll
                for (int q = 1; q < 7; q++) {
                    ex = fma(TS_BT(q), kx[1 + q], ex);
                    ey = fma(TS_BT(q), ky[1 + q], ey);
                    if (q < 6) {
                        lux = fma(TS_A(5, q), kx[1 + q], lux);
                        luy = fma(TS_A(5, q), ky[1 + q], luy);
                    }
                }
                const double u0x = kx[0], u0y = ky[0];
                const double u1x = fma(dts, lux, u0x), u1y = fma(dts, luy, u0y);
                const double m0 = sqrt(fma(u0x, u0x, u0y * u0y)), m1 = sqrt(fma(u1x, u1x, u1y * u1y));
                const double sk = o.abstol + fmax(m0, m1) * o.reltol;
                const double rx = um_div(dts * ex, sk), ry = um_div(dts * ey, sk);
                pe[j] = valid ? fma(rx, rx, ry * ry) : 0.;
                bad[j] = valid && ((u1x != u1x) || (u1y != u1y));
            }
#pragma unroll
            for (int off = 1; off < 32; off <<= 1)
#pragma unroll
                for (int j = 0; j < JB; j++) pe[j] += __shfl_xor_sync(0xffffffffu, pe[j], off);
#pragma unroll
            for (int j = 0; j < JB; j++) {
                const bool anybad = __any_sync(0xffffffffu, bad[j]);
                const bool anyclamp = __any_sync(0xffffffffu, (clampmask >> j) & 1u);
                if (lane == 0) {
                    sm_red0[wid * JB + j] = pe[j];
                    sm_red1[wid * JB + j] = anybad ? 1. : 0.;
                    sm_flag[wid * JB + j] = anyclamp ? 1 : 0;
                }
            }
        }
        __syncthreads();
    }
#undef KX
#undef KY
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
}

// KL histograms and the correlation ECDF / histogram of the runs of one chunk from their saved samples: one CTA per run,
// thread r = oscillator r — the reduction sl_finish_run<true> does, with the whole machine instead of one stalled CTA.
__global__ void __launch_bounds__(128) k_sl_post(const SlFin* __restrict__ fin, double* smp, const double* __restrict__ stat,
                                                 const int* __restrict__ okflag, const int* __restrict__ filter_pos,
                                                 const long long run_begin, const int scratch_bytes) {
    extern __shared__ __align__(16) unsigned char post_smem[];
    unsigned char* scratch = post_smem;
    double* sm_mx = reinterpret_cast<double*>(post_smem + scratch_bytes);
    double* sm_my = sm_mx + 128;
    double* sm_var = sm_my + 128;
    int* sm_hist = reinterpret_cast<int*>(sm_var + 128);
    const int tid = threadIdx.x, N = fin->N, n_t = fin->n_t;
    const size_t rl = blockIdx.x;
    const bool valid = tid < N;
    double4 st = make_double4(0., 0., 0., 0.);
    if (valid) st = *reinterpret_cast<const double4*>(stat + (rl * N + tid) * 4);
    sl_finish_run<true>(fin, smp + rl * (size_t)n_t * (2 * N), okflag[rl] != 0, run_begin + (long long)rl, valid ? filter_pos[tid] : -1,
                        st.x, st.y, st.z, st.w, scratch, scratch_bytes, sm_mx, sm_my, sm_var, sm_hist);
}

// The matrix measure of one run per CTA, tiled through shared memory: thread b owns the 4 x 4 tile number b of the upper
// triangle of the N x N complex correlation for the WHOLE series (32 accumulators), the centred samples arrive in chunks
// of SLC_K save times ([k][2N] doubles, centred while they are staged: no in-place pass over the scratch).  Same products
// in the same order as sl_corr_hist — which re-read every sample row from L2 once per tile (11.6 MB per run at N = 100) and
// ran at 13 % of the DFMA peak.
constexpr int SLC_K = 32;
// MAXT: threads per CTA >= tiles of the upper triangle (351 at N <= 104, 528 at N = 128)
template <int MAXT>
__global__ void __launch_bounds__(MAXT) k_sl_corr_post(const SlFin* __restrict__ fin, const double* __restrict__ smp,
                                                             const double* __restrict__ stat, const int* __restrict__ okflag,
                                                             const long long run_begin) {
    extern __shared__ __align__(16) unsigned char slc_smem[];
    const int tid = threadIdx.x, nth = blockDim.x, N = fin->N, n_t = fin->n_t, nb = fin->corr_nbins, n2 = 2 * N;
    // one complex sample = one 16-byte chunk; the chunk of oscillator i sits at i ^ ((i >> 3) & 7) of its row, so that the
    // lanes of a warp, which read the four consecutive columns of consecutive tiles (64-byte stride), spread over all banks
    const int rs = 2 * ((N + 7) & ~7);                        // row stride (doubles)
    double* s_x = reinterpret_cast<double*>(slc_smem);        // [SLC_K][rs]
    double* s_mu = s_x + (size_t)SLC_K * rs;                  // [2N] means (re, im interleaved)
    double* s_var = s_mu + n2;                                // [N]
    int* s_hist = reinterpret_cast<int*>(s_var + N);          // [nb]
    const size_t rl = blockIdx.x;
    const long long run = run_begin + (long long)rl, n_mc = fin->n_mc;
    if (!okflag[rl]) {
        for (int b = tid; b < nb; b += nth) fin->matrix[(long long)b * n_mc + run] = nan("");
        return;
    }
    for (int i = tid; i < N; i += nth) {
        const double4 st = *reinterpret_cast<const double4*>(stat + (rl * N + i) * 4);  // amx, a2x, amy, a2y
        s_mu[2 * i] = st.x;
        s_mu[2 * i + 1] = st.z;
        s_var[i] = st.y + st.w;
    }
    for (int b = tid; b < nb; b += nth) s_hist[b] = 0;
    const int NBk = (N + 3) >> 2;
    const int n_blocks = NBk * (NBk + 1) / 2;
    const bool work = tid < n_blocks;
    int bi = 0, bj = 0;
    if (work) {
        int rem = tid;  // tile id -> (bi, bj), bi <= bj, rows enumerated bi = 0, 1, ... with NBk - bi entries each
        while (rem >= NBk - bi) {
            rem -= NBk - bi;
            bi++;
        }
        bj = bi + rem;
    }
    // the last tile row / column may reach past N: clamp the loads to valid rows, the pairs are skipped below
    const int ia = min(4 * bi, N - 4 < 0 ? 0 : N - 4), ic = min(4 * bj, N - 4 < 0 ? 0 : N - 4);
    const int sa = 4 * bi - ia, sc = 4 * bj - ic;
    double re[4][4], im[4][4];
    int oa[4], oc[4];
#pragma unroll
    for (int p = 0; p < 4; p++) {
        oa[p] = 2 * ((ia + p) ^ (((ia + p) >> 3) & 7));
        oc[p] = 2 * ((ic + p) ^ (((ic + p) >> 3) & 7));
#pragma unroll
        for (int q = 0; q < 4; q++) re[p][q] = im[p][q] = 0.;
    }
    const double* const src = smp + rl * (size_t)n_t * n2;
    // the samples of the NEXT chunk travel to registers while the current one is multiplied (the scratch is HBM-cold: the
    // first version waited for every load — ncu: long scoreboard 3.4 stalls per issue, FP64 pipe 31 %)
    constexpr int PRE = 19;  // >= SLC_K * 2N / MAXT (blockDim is always MAXT)
    double pre[PRE];
    auto prefetch = [&](const int k0) {
        const int cnt = min(SLC_K, n_t - k0) * n2;
#pragma unroll
        for (int q = 0; q < PRE; q++) {
            const int e = tid + q * MAXT;
            pre[q] = (k0 < n_t && e < cnt) ? __ldcs(src + (size_t)k0 * n2 + e) : 0.;
        }
    };
    prefetch(1);
    for (int k0 = 1; k0 < n_t; k0 += SLC_K) {
        const int kc = min(SLC_K, n_t - k0);
        __syncthreads();  // the previous chunk is consumed (first pass: s_mu / s_var / s_hist are written)
#pragma unroll
        for (int q = 0; q < PRE; q++) {
            const int e = tid + q * MAXT;
            if (e < kc * n2) {
                const int kk = e / n2, d = e - kk * n2, i = d >> 1;
                s_x[kk * rs + 2 * (i ^ ((i >> 3) & 7)) + (d & 1)] = pre[q] - s_mu[d];
            }
        }
        __syncthreads();
        prefetch(k0 + SLC_K);
        if (work) {
#pragma unroll 4
            for (int kk = 0; kk < kc; kk++) {
                const double* row = s_x + (size_t)kk * rs;
                double2 za[4], zc[4];
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    za[p] = *reinterpret_cast<const double2*>(row + oa[p]);
                    zc[p] = *reinterpret_cast<const double2*>(row + oc[p]);
                }
#pragma unroll
                for (int p = 0; p < 4; p++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {  // (a + ib)(c - id)
                        re[p][q] = fma(za[p].x, zc[q].x, re[p][q]);
                        re[p][q] = fma(za[p].y, zc[q].y, re[p][q]);
                        im[p][q] = fma(za[p].y, zc[q].x, im[p][q]);
                        im[p][q] = fma(-za[p].x, zc[q].y, im[p][q]);
                    }
            }
        }
    }
    if (work) {
        const double step = 1. / (double)nb;
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = ia + p, j = ic + q;
                if (p < sa || q < sc) continue;  // rows of a clamped window that belong to the previous tile
                if (i >= j || j >= N) continue;
                const double den = sqrt(s_var[i]) * sqrt(s_var[j]);
                const double c = hypot(re[p][q] / den, im[p][q] / den);
                if (c != c) continue;  // zero-variance oscillator: dropped, as in the lane kernels and the oracle
                double f = c / step + 1.;  // Base.searchsortedlast on the range 0:1/nbins:1
                f = fmin(fmax(f, 1.), (double)(nb + 1));
                const int n0 = (int)rint(f);
                const int idx = (c < (double)(n0 - 1) / (double)nb) ? n0 - 1 : n0;
                if (idx >= 1 && idx <= nb) atomicAdd(&s_hist[idx - 1], 2);
            }
    }
    __syncthreads();
    if (tid == 0) {
        double tot = 0., cum = 0.;
        for (int b = 0; b < nb; b++) tot += (double)s_hist[b];
        for (int b = 0; b < nb; b++) {
            cum += (double)s_hist[b];
            fin->matrix[(long long)b * n_mc + run] = (fin->matrix_meas == MCBB_MAT_CORR_HIST) ? (double)s_hist[b] : cum / tot;
        }
    }
}

static size_t sl_smem_bytes(int JB, int N, bool a1d) {
    return (size_t)JB * (a1d ? 3072 : 2048) + sizeof(double) * ((size_t)(N + 1) * (2 * UM_NSLOT * JB + 2) + 8 * (size_t)JB + 3 * 128) +
           sizeof(SlSlotCtl) * (size_t)JB + 24 + sizeof(int) * (4 * (size_t)JB + 64) + 128;
}

template <int JB, int MINB, bool SMP, bool A1D>
static int launch_sl_inst(const SlArgs& a, size_t smem, int grid, cudaStream_t st) {
    MCBB_CUDA_OK(cudaFuncSetAttribute(k_solve_batched_umma_sl<JB, MINB, SMP, A1D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_solve_batched_umma_sl<JB, MINB, SMP, A1D><<<grid, 128, smem, st>>>(a);
    MCBB_LAUNCHED();
    return 0;
}

template <int JB, int MINB>
static int launch_sl_one(const SlArgs& a, size_t smem, int grid, bool smp, bool a1d, cudaStream_t st) {
    if (smp) return a1d ? launch_sl_inst<JB, MINB, true, true>(a, smem, grid, st) : launch_sl_inst<JB, MINB, true, false>(a, smem, grid, st);
    return a1d ? launch_sl_inst<JB, MINB, false, true>(a, smem, grid, st) : launch_sl_inst<JB, MINB, false, false>(a, smem, grid, st);
}

// Returns handled = true when the problem was run on the tcgen05 path.  `range_overflow` (may be null) receives the
// number of runs whose state left the fixed-point range; the caller repeats those ensembles on the lane kernels.
int launch_batched_umma_sl(const mcbb_system_desc* hs, const SysDev& S, const SolverDev& o, const FeatDev& F, const SolveIO& io,
                           cudaStream_t st, bool* handled) {
    *handled = false;
    const int N = S.n_osc;
    if (S.system != MCBB_SYS_STUART_LANDAU || hs == nullptr) return 0;
    if (o.alg != MCBB_ALG_TSIT5 || F.n_global > 0 || F.cyclic) return 0;
    if (N < 33 || N > 128) return 0;  // one 128-row MMA tile; small rings stay on the lane kernels
    if (io.n_t > UM_MAX_NT) return 0;
    if (tune_get("no_umma", 0) || tune_get("no_umma_sl", 0)) return 0;  // test hook
    if (F.need_kl && (F.kl_bins > UM_MAX_KL_BINS || F.kl_bins < 3)) return 0;
    if (F.matrix_meas != MCBB_MAT_NONE && (F.filter != nullptr || F.corr_nbins > 64 || F.corr_nbins < 1)) return 0;
    for (int m = 0; m < F.n_meas; m++)
        if (F.meas[m] != MCBB_MEAS_MEAN && F.meas[m] != MCBB_MEAS_STD && F.meas[m] != MCBB_MEAS_KL_HIST) return 0;
    for (int p = 0; p < S.n_par; p++)
        if (S.par_slot[p] < 0 || S.par_slot[p] > 4) return 0;
    if (!(hs->scalars[3] != 0.) || !(hs->scalars[4] != 0.)) return 0;

    SlArgs a = {};
    a.N = N;
    a.KS = (N + 31) / 32;
    a.inv_n = 1. / (double)N;
    a.ln_qoldinit = std::log(o.qoldinit);
    std::vector<unsigned char> img1(UM_A_BYTES, 0), img2(UM_A_BYTES, 0);
    std::vector<int> deg1(128, 0), deg2(128, 0);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            if (hs->mat_a[i + (size_t)N * j] != 0.) {
                img1[(size_t)i * 128 + j] = 1;
                deg1[i]++;
            }
            if (hs->mat_b[i + (size_t)N * j] != 0.) {
                img2[(size_t)i * 128 + j] = 1;
                deg2[i]++;
            }
        }
    // sparse A1 (every degree <= 8): summed directly, only A2 on the tensor core
    int nb1max = 0;
    for (int i = 0; i < N; i++) nb1max = std::max(nb1max, deg1[i]);
    const bool a1d = nb1max <= 8 && tune_get("umma_sl_a1d", 1) != 0;
    std::vector<unsigned char> nbr1(128 * 8, 0);
    if (a1d)
        for (int i = 0; i < N; i++) {
            int q = 0;
            for (int j = 0; j < N; j++)
                if (img1[(size_t)i * 128 + j]) nbr1[(size_t)i * 8 + q++] = (unsigned char)j;
        }
    std::vector<int> fpos(128, -1);
    if (F.filter) {
        std::vector<int> fh(F.n_filter);
        MCBB_CUDA_OK(cudaMemcpyAsync(fh.data(), F.filter, sizeof(int) * F.n_filter, cudaMemcpyDeviceToHost, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));
        for (int ii = 0; ii < F.n_filter; ii++) fpos[fh[ii] - 1] = ii;
    } else {
        for (int i = 0; i < N; i++) fpos[i] = i;
    }
    // fixed-point range: 4 x the larger of the limit-cycle radius sqrt(lambda) (over the swept range unknown here: the
    // base value, and any swept lambda is covered by the clamp counter) and the largest initial |component|
    double amp = 1.;
    {
        const double lam = hs->scalars[1];
        if (lam > 0.) amp = std::max(amp, std::sqrt(lam));
        // the ensemble's initial conditions live on the device: one reduction over |ic| would cost a pass; the clamp
        // counter below catches anything this bound misses, so use the documented IC range of the problem class (<= 2)
        amp = std::max(amp, 2.);
    }
    double R = 1.;
    while (R < 4. * amp) R *= 2.;
    a.fx_scale = 1125899906842624.0 / R;  // 2^50 / R
    a.fx_inv = R / 1125899906842624.0;
    a.fx_max = R * (1. - 1e-9);

    unsigned char* d_img1 = (unsigned char*)scratch_get(69, img1.size());
    unsigned char* d_img2 = (unsigned char*)scratch_get(70, img2.size());
    int* d_deg1 = (int*)scratch_get(71, sizeof(int) * 128);
    int* d_deg2 = (int*)scratch_get(72, sizeof(int) * 128);
    int* d_fpos = (int*)scratch_get(73, sizeof(int) * 128);
    unsigned int* d_ovf = (unsigned int*)scratch_get(74, 64);
    unsigned char* d_nbr1 = (unsigned char*)scratch_get(93, nbr1.size());
    if (!d_nbr1) MCBB_FAIL("solve: device allocation failed");
    MCBB_CUDA_OK(cudaMemcpyAsync(d_nbr1, nbr1.data(), nbr1.size(), cudaMemcpyHostToDevice, st));
    if (!d_img1 || !d_img2 || !d_deg1 || !d_deg2 || !d_fpos || !d_ovf) MCBB_FAIL("solve: device allocation failed");
    MCBB_CUDA_OK(cudaMemcpyAsync(d_img1, img1.data(), img1.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_img2, img2.data(), img2.size(), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_deg1, deg1.data(), sizeof(int) * 128, cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_deg2, deg2.data(), sizeof(int) * 128, cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemcpyAsync(d_fpos, fpos.data(), sizeof(int) * 128, cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaMemsetAsync(d_ovf, 0, 64, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    a.a1_img = d_img1;
    a.a2_img = d_img2;
    a.deg1 = d_deg1;
    a.deg2 = d_deg2;
    a.nbr1 = d_nbr1;
    a.nb1max = nb1max;
    a.filter_pos = d_fpos;
    a.ovf = d_ovf;
    for (int q = 0; q < MCBB_N_SCALARS; q++) a.scalars[q] = S.scalars[q];
    a.n_par = S.n_par;
    for (int p = 0; p < MCBB_MAX_PAR; p++) a.par_slot[p] = S.par_slot[p];
    a.o = o;
    a.n_meas = F.n_meas;
    for (int m = 0; m < F.n_meas; m++) {
        a.meas[m] = F.meas[m];
        a.comp[m] = F.comp[m];
    }
    a.n_filter = F.n_filter;
    a.io = io;
    a.need_kl = F.need_kl;
    a.kl_bins = F.kl_bins;
    a.kl_nstd = F.kl_nstd;
    a.kl_tol = F.kl_tol;
    a.matrix_meas = F.matrix_meas;
    a.corr_nbins = F.corr_nbins;

    // 2 trajectories x 5 CTAs per SM (20 warps, 96 registers) beats 4 x 3 (12 warps) by 6 % for mean / std and 3 % with the
    // KL-hist measures — the kernel is latency-bound per warp, so more resident warps pay (profiles/r2_sl_jb_variants.jsonl;
    // 2 x 6 at 80 registers: 132 k against 134 k for 2 x 5 — the gain from more warps ends there)
    struct Cfg { int jb, per_sm; };
    static const Cfg cfgs[] = {{2, 5}, {4, 3}, {4, 2}, {2, 2}, {2, 1}};
    const int want_jb = (int)tune_get("umma_sl_jb", 0), want_per = (int)tune_get("umma_sl_per_sm", 0);
    const Cfg* pick = nullptr;
    for (const Cfg& c : cfgs) {
        if (want_jb && c.jb != want_jb) continue;
        if (want_per && c.per_sm != want_per) continue;
        const size_t need = sl_smem_bytes(c.jb, N, a1d);
        if ((need + 1024) * c.per_sm <= (size_t)227 * 1024 && need <= (size_t)226 * 1024) {
            pick = &c;
            break;
        }
    }
    if (!pick) return 0;
    const size_t smem = sl_smem_bytes(pick->jb, N, a1d);
    int dev = 0, sms = 148;
    MCBB_CUDA_OK(cudaGetDevice(&dev));
    MCBB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    a.n_sms = sms;
    const bool smp = F.need_kl || F.matrix_meas != MCBB_MAT_NONE;
    SlFin hf = {};
    hf.N = N;
    hf.n_t = io.n_t;
    hf.n_meas = F.n_meas;
    for (int m = 0; m < F.n_meas; m++) {
        hf.meas[m] = F.meas[m];
        hf.comp[m] = F.comp[m];
    }
    hf.n_filter = F.n_filter;
    hf.kl_bins = F.kl_bins;
    hf.matrix_meas = F.matrix_meas;
    hf.corr_nbins = F.corr_nbins;
    hf.n_mc = io.n_mc;
    hf.kl_nstd = F.kl_nstd;
    hf.kl_tol = F.kl_tol;
    hf.feat = io.feat;
    hf.matrix = io.matrix;
    SlFin* d_fin = (SlFin*)scratch_get(92, 2 * sizeof(SlFin));
    if (!d_fin) MCBB_FAIL("solve: device allocation failed");
    SlFin hf2[2] = {hf, hf};
    hf2[1].matrix_meas = MCBB_MAT_NONE;  // k_sl_post: measures only — the matrix measure has its own kernel
    MCBB_CUDA_OK(cudaMemcpyAsync(d_fin, hf2, 2 * sizeof(SlFin), cudaMemcpyHostToDevice, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));  // hf2 is stack-owned
    a.fin = d_fin;
    MCBB_CUDA_OK(cudaMemcpyToSymbolAsync(c_um_ts, io.t_save, sizeof(double) * io.n_t, 0, cudaMemcpyDeviceToDevice, st));

    // Measures that need the whole series (KL-hist, correlation): the ensemble is integrated in chunks whose saved samples
    // fit a scratch (test hook "sl_post_mb"), each chunk followed by k_sl_post.  Other measure sets: one launch.
    const size_t run_bytes = sizeof(double) * (size_t)io.n_t * (size_t)(2 * N);
    long long chunk = io.n_mc;
    if (smp) {
        // at most 64 GB and at most half of what is free on the device right now (the scratch is kept for the next call)
        // (cudaMemGetInfo costs milliseconds: only ensembles whose samples exceed 2 GB ask how much memory is free)
        size_t cap = (size_t)2 << 30;
        if (run_bytes * (size_t)io.n_mc > cap) {
            size_t free_b = 0, total_b = 0;
            MCBB_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
            cap = std::max<size_t>(cap, std::min<size_t>((size_t)64 << 30, free_b / 2 + scratch_size(39)));
        }
        if (const long long mb = tune_get("sl_post_mb", 0)) cap = (size_t)mb << 20;
        chunk = std::max<long long>(1, std::min<long long>(io.n_mc, (long long)(cap / run_bytes)));
        a.smp = (double*)scratch_get(39, run_bytes * (size_t)chunk);
        a.stat = (double*)scratch_get(106, sizeof(double) * 4 * (size_t)N * (size_t)chunk);
        a.okflag = (int*)scratch_get(107, sizeof(int) * (size_t)chunk);
        if (!a.smp || !a.stat || !a.okflag) MCBB_FAIL("solve: device allocation failed");
    }
    const int post_scratch = smp ? (F.need_kl ? F.kl_bins * 256 : 0) : 0;
    const size_t post_smem = (size_t)post_scratch + sizeof(double) * 3 * 128 + sizeof(int) * 64;
    for (long long r0 = 0; r0 < io.n_mc; r0 += chunk) {
        const long long r1 = std::min<long long>(io.n_mc, r0 + chunk);
        a.run_begin = r0;
        a.run_end = r1;
        const unsigned long long c0 = (unsigned long long)r0;
        MCBB_CUDA_OK(cudaMemcpyAsync(io.work_counter, &c0, sizeof(c0), cudaMemcpyHostToDevice, st));
        MCBB_CUDA_OK(cudaStreamSynchronize(st));  // c0 is stack-owned
        const long long want = (r1 - r0 + pick->jb - 1) / pick->jb;
        const int grid = (int)std::min<long long>(want, (long long)sms * pick->per_sm);
        int rc = 1;
        if (pick->jb == 4 && pick->per_sm == 3) rc = launch_sl_one<4, 3>(a, smem, grid, smp, a1d, st);
        else if (pick->jb == 4 && pick->per_sm == 2) rc = launch_sl_one<4, 2>(a, smem, grid, smp, a1d, st);
        else if (pick->jb == 2 && pick->per_sm == 2) rc = launch_sl_one<2, 2>(a, smem, grid, smp, a1d, st);
        else if (pick->jb == 2 && pick->per_sm == 1) rc = launch_sl_one<2, 1>(a, smem, grid, smp, a1d, st);
        else if (pick->jb == 2 && pick->per_sm == 5) rc = launch_sl_one<2, 5>(a, smem, grid, smp, a1d, st);
        if (rc) return rc;
        if (smp) {
            MCBB_CUDA_OK(cudaFuncSetAttribute(k_sl_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem));
            k_sl_post<<<(unsigned)(r1 - r0), 128, post_smem, st>>>(d_fin + 1, a.smp, a.stat, a.okflag, d_fpos, r0, post_scratch);
            MCBB_LAUNCHED();
            if (F.matrix_meas != MCBB_MAT_NONE) {
                const int nbk = (N + 3) / 4, tiles = nbk * (nbk + 1) / 2;
                const int threads = ((tiles + 31) / 32) * 32 <= 352 ? 352 : 544;  // always the template's MAXT: idle threads help staging
                const size_t csm = sizeof(double) * ((size_t)SLC_K * 2 * ((N + 7) & ~7) + 2 * N + N) + sizeof(int) * 64;
                if (threads <= 352) {
                    MCBB_CUDA_OK(cudaFuncSetAttribute(k_sl_corr_post<352>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
                    k_sl_corr_post<352><<<(unsigned)(r1 - r0), threads, csm, st>>>(d_fin, a.smp, a.stat, a.okflag, r0);
                } else {
                    MCBB_CUDA_OK(cudaFuncSetAttribute(k_sl_corr_post<544>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
                    k_sl_corr_post<544><<<(unsigned)(r1 - r0), threads, csm, st>>>(d_fin, a.smp, a.stat, a.okflag, r0);
                }
                MCBB_LAUNCHED();
            }
        }
    }
    // a run that left the fixed-point range invalidates this launch: the caller falls through to the lane kernels
    unsigned int h_ovf = 0;
    MCBB_CUDA_OK(cudaMemcpyAsync(&h_ovf, d_ovf, sizeof(h_ovf), cudaMemcpyDeviceToHost, st));
    MCBB_CUDA_OK(cudaStreamSynchronize(st));
    if (h_ovf > 0) return 0;  // handled stays false
    *handled = true;
    return 0;
}

}  // namespace mcbb
