// batched_umma_lw.cuh — "light warp" variant of k_solve_batched_umma (included by batched_umma.cu) for networks whose last
// 32-row block holds only a few oscillators (N = 97 .. 96 + 32 / JB; the headline C5 has N = 100).
//
// In the plain mapping thread r owns oscillator row r for all JB trajectories of the CTA, so the fourth warp of a CTA works
// on N - 96 rows at the cost of 32: at N = 100 that is 22 % of all issue and FP64 slots spent on dummy rows.  Here the
// warp whose tensor-memory quadrant holds the short block carries ONE (row, trajectory) element per lane instead
// ((N - 96) * JB <= 32 lanes), i.e. one sixth of the per-stage work of a full warp, and also hosts the controller lanes.
// Which quadrant is the short one rotates from CTA to CTA (rho = resident-CTA index mod 4; the oscillator rows of block b
// live in quadrant (b + rho) mod 4, every thread simply loads ITS oscillator's row of A into ITS tensor-memory lane), so
// that each scheduler partition of an SM hosts three full warps and one light warp of the four resident CTAs.
//
// The tensor core computes D[lane, all trajectories]: lane l of the light warp holds row 96 + l % NR of A, so its
// accumulators are the sums of that oscillator for every trajectory; it keeps the 16 columns of its own trajectory
// (one tcgen05.ld per trajectory + a predicated copy).  Reductions over the rows of a trajectory (error norm, initial
// step) fetch the NR partials of a trajectory with shuffles and add them in the order the 32-lane butterfly of the plain
// mapping would, so results are bit-identical to k_solve_batched_umma (tools/ab_variants.py checks it).
#pragma once

namespace mcbb {

template <int V>
struct UmInt {
    static constexpr int value = V;
};

template <int JB, int MINB>
__global__ void __launch_bounds__(128, MINB) k_solve_batched_umma_lw(const UmmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int N = a.N;
    const int NR = N - 96;                           // rows of the short block (1 .. 32 / JB)
    // a.lw_mode (experiment switch): 1 = light warp + rotation, 2 = light warp, no rotation, 3 / 4 = every warp runs the full
    // body (the plain mapping through this code structure) with / without rotation
    const int rho = (a.lw_mode == 1 || a.lw_mode == 3) ? (blockIdx.x / a.n_sms) & 3 : 0;  // rotation of the row blocks over the quadrants
    const int blk = (wid - rho) & 3;                 // row block of this warp; block 3 is the short one
    const bool light = (blk == 3) && a.lw_mode <= 2;
    const int jl = light ? min(lane / NR, JB - 1) : 0;                  // trajectory of a light lane
    const bool valid = light ? (lane < NR * JB) : (32 * blk + lane < N);
    const int row = light ? (valid ? 96 + lane % NR : N) : (valid ? 32 * blk + lane : N);  // oscillator row (N = dummy, zero row of A)
    const SolverDev& o = a.o;
    const int n_t = a.io.n_t;
    const long long n_mc = a.io.n_mc;
    constexpr unsigned NCOL = 16 * JB;
    constexpr int RS = UM_NSLOT * JB + 2;

    unsigned char* sB = smem;                                   // [JB][128][16]
    double* sm_k = (double*)(sB + JB * 2048);                   // [N + 1][JB][8] (+2 pad)
    const double* const sm_ts = c_um_ts;
    double* sm_red0 = sm_k + (size_t)(N + 1) * RS;              // [4 row blocks][JB]
    double* sm_red1 = sm_red0 + 4 * JB;
    UmSlotCtl* ctl = (UmSlotCtl*)(sm_red1 + 4 * JB);
    unsigned long long* mbar_p = (unsigned long long*)(ctl + JB);
    unsigned* tmem_slot = (unsigned*)(mbar_p + 2);
    unsigned long long* full_p = mbar_p + 1;
    const unsigned mbar_full = um_smem_u32(full_p);
    const unsigned mbar = um_smem_u32(mbar_p);
    double* const kb = sm_k + (size_t)row * RS;

#define KP(slot, j) kb[(j)*UM_NSLOT + (slot)]

    for (int i = tid; i < JB * 2048 / 16; i += 128) ((uint4*)sB)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < (N + 1) * RS; i += 128) sm_k[i] = 0.;
    const bool is_ctl = (blk == 3) && lane < JB;  // the controller lanes live in the warp of the short block
    if (is_ctl) {
        UmSlotCtl& c = ctl[lane];
        c.mode = MODE_FETCH;
        c.accept = c.fresh = c.fin = 0;
        c.hstep = 0.;
        c.dts = 0.;
        c.al1s = 0.;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(mbar_full));
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(tmem_slot)),
                     "n"(UmCols<JB>::value));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const unsigned tmem_row = tmem_base + ((unsigned)(32 * wid) << 16);  // this warp's 32 TMEM lanes (quadrant = warp)
    const unsigned tmem_a = tmem_base + NCOL;
    {  // every thread: the row of A of ITS oscillator -> ITS tensor-memory lane (rows >= N of the image are zero)
        unsigned w[32];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint4 t4 = reinterpret_cast<const uint4*>(a.a_img + (size_t)row * 128)[q];
            w[4 * q] = t4.x;
            w[4 * q + 1] = t4.y;
            w[4 * q + 2] = t4.z;
            w[4 * q + 3] = t4.w;
        }
        um_tmem_st32(tmem_row + NCOL, w);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const unsigned idesc = (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (1u << 16) | ((NCOL >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned sB_addr = um_smem_u32(sB);
    const int KSTEPS = a.KS;
    // rows >= N publish nothing: the dummy lanes of the light warp would all write row N of the panels (a race between
    // lanes of different trajectories is harmless — row N meets a zero column of A — but keep the panel row zero)
    unsigned char* const sB_row = sB + row * 16;

    const double bias = a.bias[row];
    const int degh = a.deg[row] << 19;
    const int fpos = valid ? a.filter_pos[row] : -1;
    unsigned phase = 0;

    auto begin_attempt = [&](UmSlotCtl& c) -> bool {
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

    // sum over the rows of block 3 of one trajectory, in the order of the 32-lane xor butterfly of the plain mapping with
    // rows 96.. in lanes 0.. and zeros behind them: ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7)) + ...  (NR <= 16)
    auto light_group_sum = [&](const double p) -> double {
        const int base = jl * NR;
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const double t = __shfl_sync(0xffffffffu, p, (base + q) & 31);
            v[q] = (q < NR) ? t : 0.;
        }
#pragma unroll
        for (int w = 1; w < 16; w <<= 1)
#pragma unroll
            for (int q = 0; q < 16; q += 2 * w) v[q] += v[q + w];
        return v[0];
    };

    auto body = [&](auto JT, auto LT) {
        constexpr int JBT = decltype(JT)::value;     // elements (trajectories) per thread: JB (full warps) or 1 (light warp)
        constexpr bool LIGHT = decltype(LT)::value != 0;
        constexpr int G = JBT <= 4 ? JBT : (JBT % 3 == 0 ? 3 : 4);
        static_assert(JBT % G == 0, "interleave group");
        double sv[JBT], cv[JBT], un[JBT], am[JBT], a2[JBT], as[JBT];
#pragma unroll
        for (int j = 0; j < JBT; j++) sv[j] = cv[j] = un[j] = am[j] = a2[j] = as[j] = 0.;
        auto jof = [&](const int jj) -> int { return LIGHT ? jl : jj; };

        for (;;) {
            // ---- per-trajectory control (the light warp's first JB lanes) ------------------------------------------------
            if (is_ctl) {
                UmSlotCtl& c = ctl[lane];
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
                    for (int w = 0; w < 4; w++) {  // row blocks in ascending order, as the plain mapping sums its warps
                        e2 += sm_red0[w * JB + lane];
                        nb += sm_red1[w * JB + lane];
                    }
                    const double m2 = e2 * a.inv_n;  // EEst^2
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
                        if (nb > 0.) c.ret = MCBB_RET_UNSTABLE;
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
                    c.mode = MODE_FETCH;
                }
                if (c.mode == MODE_FETCH) {
                    const long long run = (long long)atomicAdd(a.io.work_counter, 1ULL);
                    if (run >= n_mc) {
                        c.mode = MODE_DONE;
                    } else {
                        double sc[MCBB_N_SCALARS];
                        for (int q = 0; q < MCBB_N_SCALARS; q++) sc[q] = a.scalars[q];
                        for (int p = 0; p < a.n_par; p++) sc[a.par_slot[p]] = a.io.par[(long long)p * n_mc + run];
                        c.al1s = ((sc[0] / (double)N) * a.unit_weight) * 8.881784197001252e-16;
                        c.run = run;
                        c.t = o.t0;
                        c.iter = 0;
                        c.nrej = 0;
                        c.natt = 0;
                        c.lq = a.ln_qoldinit;
                        c.hstep = 0.;
                        c.isave = 0;
                        c.ret = 0;
                        c.mode = MODE_INIT;
                        c.fresh = 1;
                    }
                }
            }
            const bool all_done = __syncthreads_and(!is_ctl || ctl[lane].mode == MODE_DONE);

            // ---- element owners: samples of the accepted step, u <- u_new, results of finished runs, new initial state ----
            bool any_init = false;
            unsigned runmask = 0u;
#pragma unroll
            for (int j = 0; j < JB; j++) {
                any_init |= (ctl[j].mode == MODE_INIT);
                runmask |= (ctl[j].mode == MODE_RUN ? 1u : 0u) << j;
            }
#pragma unroll
            for (int jj = 0; jj < JBT; jj++) {
                const int j = jof(jj);
                const UmSlotCtl& c = ctl[j];
                if (c.accept) {
                    const int ks0 = c.s0 < 1 ? 1 : c.s0, ks1 = c.s1;  // sample 0 is dropped (eval_mc_prob.jl:152)
                    if (ks0 < ks1) {
                        const double dts = c.dts_old, t_old = c.t_old;
                        double kk[8];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const double2 t2 = reinterpret_cast<const double2*>(kb + j * UM_NSLOT)[q];
                            kk[2 * q] = t2.x;
                            kk[2 * q + 1] = t2.y;
                        }
                        for (int k = ks0; k < ks1; k++) {
                            const double th = um_div(sm_ts[k] - t_old, dts);
                            double lin;
                            if constexpr (LIGHT) {
                                // lanes of the light warp belong to different trajectories (different theta): every lane
                                // evaluates its own seven dense-output weights, same expressions as the lane-parallel form
                                lin = 0.;
#pragma unroll
                                for (int q = 0; q < 7; q++) {
                                    const double inner = fma(th, fma(th, c_um_bth[q][2], c_um_bth[q][1]), c_um_bth[q][0]);
                                    const double bq = (q == 0) ? th * fma(th, inner, 1.) : (th * th) * inner;
                                    lin = (q == 0) ? kk[1] * bq : fma(kk[1 + q], bq, lin);
                                }
                            } else {
                                const int lq = lane < 7 ? lane : 0;
                                const double pa = c_um_bth[lq][0], pb = c_um_bth[lq][1], pc = c_um_bth[lq][2];
                                const double inner = fma(th, fma(th, pc, pb), pa);
                                const double bq = (lq == 0) ? th * fma(th, inner, 1.) : (th * th) * inner;
                                lin = kk[1] * __shfl_sync(0xffffffffu, bq, 0);
#pragma unroll
                                for (int q = 1; q < 7; q++) lin = fma(kk[1 + q], __shfl_sync(0xffffffffu, bq, q), lin);
                            }
                            double x = fma(dts, lin, kk[0]);
                            if (a.cyclic) {  // _cyclic_setback!, eval_mc_prob.jl:286-296
                                if (k == 1) {
                                    const double twopi = 6.283185307179586, pi = 3.141592653589793;
                                    double vv = julia_mod_pos(x, twopi);
                                    if (vv > pi) vv -= twopi;
                                    as[jj] = x - vv;
                                    x = vv;
                                } else {
                                    x = x - as[jj];
                                }
                            }
                            const double delta = x - am[jj];
                            if (fabs(delta) <= 1.7976931348623157e308) {
                                const double mnew = fma(delta, um_div(1., (double)k), am[jj]);
                                a2[jj] = fma(delta, x - mnew, a2[jj]);
                                am[jj] = mnew;
                            } else {
                                am[jj] = nonfinite_mean(am[jj], x);
                                a2[jj] = nan("");
                            }
                        }
                    }
                    KP(0, j) = un[jj];
                    KP(1, j) = KP(7, j);  // FSAL
                }
                if (c.fin && valid && fpos >= 0) {
                    const bool ok = c.fin_ok != 0;
                    const double mean = ok ? am[jj] : nan("");
                    const double sd = ok ? sqrt((a2[jj] < 0. ? 0. : a2[jj]) / (double)(n_t - 2)) : nan("");
                    for (int m = 0; m < a.n_meas; m++)
                        a.io.feat[((long long)m * a.n_filter + fpos) * n_mc + c.fin_run] = (a.meas[m] == MCBB_MEAS_MEAN) ? mean : sd;
                }
                if (c.fresh) {
                    if (valid) {
                        KP(0, j) = a.io.ic[(long long)row * n_mc + c.run];
#pragma unroll
                        for (int q = 1; q < UM_NSLOT; q++) KP(q, j) = 0.;
                    }
                    am[jj] = a2[jj] = as[jj] = 0.;
                }
            }
            if (all_done) break;

            // ---- one round: 6 right-hand-side evaluations of the whole batch -----------------------------------------
#pragma unroll 1
            for (int s = 0; s < 6; s++) {
                double cs[6];
#pragma unroll
                for (int q = 0; q < 6; q++) cs[q] = (q <= s) ? TS_A(s, q) : 0.;
                const double ci0 = (s == 1) ? 1. : 0.;
                // (a) stage input Y -> sin, cos -> the 16 digit bytes of this (row, trajectory) element
#pragma unroll
                for (int g0 = 0; g0 < JBT; g0 += G) {
                    double y[G];
#pragma unroll
                    for (int i = 0; i < G; i++) {
                        const int j = jof(g0 + i);
                        const double2* kq = reinterpret_cast<const double2*>(kb + j * UM_NSLOT);
                        const double2 k01 = kq[0], k23 = kq[1], k45 = kq[2];
                        const double k6 = kq[3].x;
                        const double c0 = ((runmask >> j) & 1u) ? cs[0] : ci0;
                        double lin = c0 * k01.y;
                        lin = fma(cs[1], k23.x, lin);
                        lin = fma(cs[2], k23.y, lin);
                        lin = fma(cs[3], k45.x, lin);
                        lin = fma(cs[4], k45.y, lin);
                        lin = fma(cs[5], k6, lin);
                        y[i] = fma(ctl[j].hstep, lin, k01.x);
                    }
                    um_sincos_vec<G>(y, &sv[g0], &cv[g0]);
#pragma unroll
                    for (int i = 0; i < G; i++) {
                        uint4 dg;
                        um_fixed52(sv[g0 + i], dg.x, dg.y);
                        um_fixed52(cv[g0 + i], dg.z, dg.w);
                        if (!LIGHT || valid) *reinterpret_cast<uint4*>(sB_row + jof(g0 + i) * 2048) = dg;
                    }
                }
                // (b) D = A * digits on the tensor core: publish, last arrival issues, issuer polls, the rest sleep on barrier 1
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
                        for (int ks = 0; ks < KSTEPS; ks++)
                            um_mma_i8_ts(tmem_base, tmem_a + 8u * ks, um_desc(sB_addr + ks * 512, 128, 2048), idesc, ks > 0 ? 1u : 0u);
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
                // (c) row of P, Q from tensor memory, int64 recombination, k = a1 (c P - s Q) + w
                double kv[JBT];
                double* const kst = kb + 2 + s;
                auto to_k = [&](const int jj, const unsigned* w) {
                    const int j = jof(jj);
                    const double P = um_recombine(w[0], w[1], w[2], w[3], w[4], w[5], w[6], degh);
                    const double Q = um_recombine(w[8], w[9], w[10], w[11], w[12], w[13], w[14], degh);
                    kv[jj] = fma(ctl[j].al1s, cv[jj] * P - sv[jj] * Q, bias);
                    kst[j * UM_NSLOT] = ((runmask >> j) & 1u) ? kv[jj] : 0.;
                };
                if constexpr (LIGHT) {
                    // the accumulators of every trajectory pass by (all loads in flight, one wait); keep those of this lane's
                    unsigned v[JB][16];
#pragma unroll
                    for (int j = 0; j < JB; j++) um_tmem_ld16_nowait(tmem_row + 16u * (unsigned)j, v[j]);
                    um_tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < JB; j++) um_tmem_landed(v[j]);
                    unsigned mine[16];
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        mine[q] = v[0][q];
#pragma unroll
                        for (int j = 1; j < JB; j++) mine[q] = (j == jl) ? v[j][q] : mine[q];
                    }
                    to_k(0, mine);
                } else {
#pragma unroll
                    for (int g0 = 0; g0 + 1 < JBT; g0 += 2) {
                        unsigned v[32];
                        um_tmem_ld32(tmem_row + 16u * (unsigned)g0, v);
                        to_k(g0, v);
                        to_k(g0 + 1, v + 16);
                    }
                    if (JBT & 1) {
                        unsigned v[16];
                        um_tmem_ld16(tmem_row + 16u * (unsigned)(JBT - 1), v);
                        to_k(JBT - 1, v);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (any_init && s < 2) {  // initial step size (OrdinaryDiffEq initdt.jl), only in rounds that start a run
                    if constexpr (LIGHT) {
                        const int j = jl;
                        double p0 = 0., p1 = 0.;
                        const bool ini = ctl[j].mode == MODE_INIT;
                        if (ini && valid) {
                            const double u0 = KP(0, j);
                            const double sk = o.abstol + fabs(u0) * o.reltol;
                            if (s == 0) {
                                KP(1, j) = kv[0];
                                const double w0 = u0 / sk, w1 = kv[0] / sk;
                                p0 = w0 * w0;
                                p1 = w1 * w1;
                            } else {
                                const double vv = (kv[0] - KP(1, j)) / sk;
                                p0 = vv * vv;
                            }
                        }
                        p0 = light_group_sum(p0);
                        p1 = light_group_sum(p1);
                        if (ini && valid && lane == j * NR) {
                            sm_red0[3 * JB + j] = p0;
                            sm_red1[3 * JB + j] = p1;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < JB; j++) {
                            if (ctl[j].mode == MODE_INIT) {
                                double p0 = 0., p1 = 0.;
                                if (valid) {
                                    const double u0 = KP(0, j);
                                    const double sk = o.abstol + fabs(u0) * o.reltol;
                                    if (s == 0) {
                                        KP(1, j) = kv[j];
                                        const double w0 = u0 / sk, w1 = kv[j] / sk;
                                        p0 = w0 * w0;
                                        p1 = w1 * w1;
                                    } else {
                                        const double vv = (kv[j] - KP(1, j)) / sk;
                                        p0 = vv * vv;
                                    }
                                }
                                p0 = um_warp_sum(p0);
                                p1 = um_warp_sum(p1);
                                if (lane == 0) {
                                    sm_red0[blk * JB + j] = p0;
                                    sm_red1[blk * JB + j] = p1;
                                }
                            }
                        }
                    }
                    __syncthreads();
                    if (is_ctl && ctl[lane].mode == MODE_INIT) {
                        UmSlotCtl& c = ctl[lane];
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

            // ---- error estimate partials and the candidate new state ------------------------------------------------------
            {
                double pe[JBT];
                bool bad[JBT];
#pragma unroll
                for (int jj = 0; jj < JBT; jj++) {
                    const int j = jof(jj);
                    const double dts = ctl[j].dts;
                    double kk[8];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const double2 t2 = reinterpret_cast<const double2*>(kb + j * UM_NSLOT)[q];
                        kk[2 * q] = t2.x;
                        kk[2 * q + 1] = t2.y;
                    }
                    double lin = TS_BT(0) * kk[1];
                    double lu = TS_A(5, 0) * kk[1];
#pragma unroll
                    for (int q = 1; q < 7; q++) {
                        lin = fma(TS_BT(q), kk[1 + q], lin);
                        if (q < 6) lu = fma(TS_A(5, q), kk[1 + q], lu);
                    }
                    const double u0 = kk[0];
                    const double u1 = fma(dts, lu, u0);
                    const double r = um_div(dts * lin, o.abstol + fmax(fabs(u0), fabs(u1)) * o.reltol);
                    pe[jj] = valid ? r * r : 0.;
                    un[jj] = u1;
                    bad[jj] = valid && (u1 != u1);
                }
                if constexpr (LIGHT) {
                    const double s0 = light_group_sum(pe[0]);
                    const unsigned bm = __ballot_sync(0xffffffffu, bad[0]);
                    if (valid && lane == jl * NR) {
                        sm_red0[3 * JB + jl] = s0;
                        sm_red1[3 * JB + jl] = ((bm >> (jl * NR)) & ((1u << NR) - 1u)) ? 1. : 0.;
                    }
                } else {
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1)
#pragma unroll
                        for (int jj = 0; jj < JBT; jj++) pe[jj] += __shfl_xor_sync(0xffffffffu, pe[jj], off);
#pragma unroll
                    for (int jj = 0; jj < JBT; jj++) {
                        const bool anybad = __any_sync(0xffffffffu, bad[jj]);
                        if (lane == 0) {
                            sm_red0[blk * JB + jj] = pe[jj];
                            sm_red1[blk * JB + jj] = anybad ? 1. : 0.;
                        }
                    }
                }
            }
            __syncthreads();
        }
    };
    if (light)
        body(UmInt<1>{}, UmInt<1>{});
    else
        body(UmInt<JB>{}, UmInt<0>{});
#undef KP
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(UmCols<JB>::value));
}

}  // namespace mcbb
