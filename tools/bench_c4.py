"""bench.py --config C4 — BASELINE.json configs[3]: nonlocal Stuart-Landau ring (paper/stuartlandau/stuart_landau_sathiyadevi-
standard-config.ipynb cells 1, 4), N = 100 complex oscillators, 2*10^5 runs, tspan (0, 200), tail 0.7, eps ~ U(1.8, 2.1),
measures [mean, std, KL] x [Re, Im] + correlation_ecdf, then normalize -> histogram-mode distance (k_bin = 0.25, weights
[0.25, 1, 0, 0.25, 1, 0, 0, 1]) -> on-the-fly (row-sharded for N > 1) DBSCAN(eps 0.008, minpts 10) -> cluster_membership.

Same JSON contract as bench.py: `value` = trajectories/s of integrate+featurize resident in HBM (K timed steps of
`--c4-runs`/N runs per GPU), `e2e` through the host C-ABI call with host buffers, `pipeline` = the whole chain once as wall
clock.  The host-side stages of the reference that stay host code here (normalize, common histogram edges, per-run
ECDF rows: O(N_mc) numpy) are timed and reported separately inside `pipeline`."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 100
WORKLOAD = ("C4 stuart_landau N=100 (lambda 1, omega 2, P1 1, P2 22, Watts-Strogatz p_r 0.04 seed 7), eps~U(1.8,2.1), "
            "z0~U(-1,1)+iU(-1,1), tspan (0,200), tail 0.7, N_t=400, Tsit5 abstol 1e-6 reltol 1e-3, measures "
            "[mean,std,KL]x[re,im] + correlation_ecdf")
WEIGHTS = [0.25, 1.0, 0.0, 0.25, 1.0, 0.0, 0.0, 1.0]
DB_EPS, DB_MINPTS = 0.008, 10


def algorithmic_flops(nnz1, nnz2, nsteps_total):
    """FP64 flops of the minimal algorithm per Tsit5 step attempt of one trajectory: 6 right-hand-side
    evaluations x (2 flops per coupling term [difference, accumulate] + 2N scalings + N x 12 for (lambda + i omega -
    |z|^2) z) + stage combinations 2*21 per real component + error estimate and new state 32 per real component."""
    per_step = 6 * (2 * (nnz1 + nnz2) + 2 * N + 12 * N) + 42 * (2 * N) + 32 * (2 * N)
    return per_step * nsteps_total


def main(args):
    import torch
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    from tests import cases
    import bench
    m = load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    m.check(m.lib().mcbb_init(local))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device("cuda", local)
    R = max(1, args.c4_runs // world)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    prob, _ = cases.c4_problem(m, 4)  # descriptors only; the ensemble arrays are generated below
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    opts = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    A1, A2 = prob.pars.A_1, prob.pars.A_2
    nnz1, nnz2 = int(np.count_nonzero(A1)), int(np.count_nonzero(A2))
    rng = np.random.default_rng(2000 + rank)
    ic = np.empty((2 * N, R))
    ic[0::2] = rng.uniform(-1, 1, (N, R))
    ic[1::2] = rng.uniform(-1, 1, (N, R))
    par = rng.uniform(1.8, 2.1, (1, R))
    ic_h = torch.empty((2 * N, R), dtype=torch.float64).pin_memory()
    ic_h.copy_(torch.from_numpy(ic))
    par_h = torch.empty((1, R), dtype=torch.float64).pin_memory()
    par_h.copy_(torch.from_numpy(par))
    ic_d, par_d = ic_h.to(dev), par_h.to(dev)
    ts_d = torch.from_numpy(ts).to(dev)
    nf, nm, nb = N, pfeat.n_meas, 30
    feat_d = torch.empty((nm * nf, R), dtype=torch.float64, device=dev)
    mat_d = torch.empty((nb, R), dtype=torch.float64, device=dev)
    ret_d = torch.empty((R,), dtype=torch.int32, device=dev)
    ns_d = torch.empty((2, R), dtype=torch.int64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    stream_ptr = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def solve_dev(icx, parx):
        m.check(m.lib().mcbb_solve_features_dev(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), R, p(icx), p(parx),
                                                p(ts_d), ts_d.numel(), p(feat_d), p(mat_d), C.c_void_p(0), p(ret_d),
                                                p(ns_d), stream_ptr()))

    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        solve_dev(ic_d, par_d)
    barrier()
    launches0 = m.kernel_launch_count()
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        l2_flush.fill_(k)
        ev[k][0].record()
        solve_dev(ic_d, par_d)
        ev[k][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = m.kernel_launch_count() - launches0
    ms_total = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    nst = ns_d[0].sum().to(torch.float64).reshape(1)
    failed = (ret_d != 0).sum().to(torch.float64).reshape(1)
    if dist is not None:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(nst, op=dist.ReduceOp.SUM)
        dist.all_reduce(failed, op=dist.ReduceOp.SUM)
    ms_total = float(ms_total.item())
    value = world * R * args.steps / (ms_total * 1e-3)

    # ---- e2e through the host entry (pinned host buffers in, features + matrix out) -------------------------
    dp, ip, lp = m.abi.c_double_p, m.abi.c_int32_p, m.abi.c_int64_p
    feat_h = torch.empty((nm * nf, R), dtype=torch.float64).pin_memory()
    mat_h = torch.empty((nb, R), dtype=torch.float64).pin_memory()
    ret_h = torch.empty((R,), dtype=torch.int32).pin_memory()
    ns_h = torch.empty((2, R), dtype=torch.int64).pin_memory()

    def call_host():
        m.check(m.lib().mcbb_solve_features(
            C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), R, C.cast(ic_h.data_ptr(), dp),
            C.cast(par_h.data_ptr(), dp), ts.ctypes.data_as(dp), len(ts), C.cast(feat_h.data_ptr(), dp),
            C.cast(mat_h.data_ptr(), dp), dp(), C.cast(ret_h.data_ptr(), ip), C.cast(ns_h.data_ptr(), lp)))
    call_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 2
    for _ in range(e2e_steps):
        call_host()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * R * e2e_steps / float(t_e2e.item())
    same = bool(torch.equal(torch.nan_to_num(feat_h), torch.nan_to_num(feat_d.cpu())))
    h2d = (ic_h.numel() + par_h.numel() + len(ts)) * 8
    d2h = (feat_h.numel() + mat_h.numel()) * 8 + ret_h.numel() * 4 + ns_h.numel() * 8

    peak = C.c_double(0.0)
    mhz = C.c_double(0.0)
    m.check(m.lib().mcbb_measure_fp64_fma_peak(C.byref(peak), C.byref(mhz)))
    fl = algorithmic_flops(nnz1, nnz2, float(nst.item()) * args.steps)
    # the N x N complex correlation of every run (eval_mc_prob.jl:462): 8 flops per (pair, sample), upper triangle
    fl_corr = 8.0 * (N * (N - 1) / 2) * (len(ts) - 1) * world * R * args.steps
    achieved = (fl + fl_corr) / (ms_total * 1e-3) / world / 1e12
    roofline = {"bound": "fp64_fma", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved / peak.value if peak.value > 0 else None, "traffic": None,
                "kernel": "k_solve_batched_umma_sl (tcgen05.mma kind::i8: A1 Re z and A2 Im z as exact fixed-point GEMMs)",
                "integrator_flops": fl, "correlation_flops": fl_corr,
                "note": "algorithmic FP64 flops (integrator per step attempt x the kernel's count, + the per-run N x N "
                        "complex correlation) / CUDA-event time; 44 of every 60 coupling flops run on the tensor pipe; the "
                        "kernel is bound by the per-stage synchronisation chain (publish -> MMA -> TMEM), not by a pipe"}

    # ---- the whole C4 chain once more as one wall-clock region ---------------------------------------------
    pipe = None
    cluster = None
    if not args.no_cluster:
        def pipeline():
            t = {}
            barrier()
            t0 = time.perf_counter()
            icx, parx = ic_h.to(dev, non_blocking=True), par_h.to(dev, non_blocking=True)
            solve_dev(icx, parx)
            torch.cuda.synchronize()
            t["h2d_solve_s"] = time.perf_counter() - t0
            # the measures stay on the device: all-gather -> sortperm by the parameter (sort!(sol, prob),
            # setup_mc_prob.jl:520-536) -> normalize + histogram edges + Float32 ECDF rows (mcbb_order_stats_dev,
            # mcbb_hist_rows_dev; bit-identical to the host mirror, tests/test_gpu_parity.py) -> fused DBSCAN.  The
            # correlation ECDF has weight 0 in the notebook's distance and is not gathered.
            Xl = torch.cat([feat_d, parx], dim=0)
            if dist is not None:
                parts = [torch.empty_like(Xl) for _ in range(world)]
                dist.all_gather(parts, Xl)
                Xa = torch.cat(parts, dim=1)
            else:
                Xa = Xl
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            t["allgather_s"] = t1 - t0 - t["h2d_solve_s"]
            n_all = Xa.shape[1]
            order = torch.sort(Xa[-1], stable=True).indices
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                X_d, gl, gw, _, _ = m.dist.histogram_features_device(Xa[:nm * nf], nm, nf, WEIGHTS[:nm] + WEIGHTS[-1:], Xa[-1:],
                                                                     order, normalize=True, k_bin=0.25)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            t["sort_normalize_hist_s"] = t2 - t1
            pr = cases.c4_problem(m, 4)[0]
            pr.par = np.asfortranarray(Xa[-1][order].cpu().numpy().reshape(-1, 1))
            pr.N_mc = n_all
            if world > 1:
                a_, s_, c_ = m.dist.dbscan_features_sharded(X_d, gl, gw, DB_EPS, DB_MINPTS, rank, world, dist)
            else:
                a_, s_, c_ = m.dist.dbscan_features_device(X_d, gl, gw, DB_EPS, DB_MINPTS)
            labels = a_.cpu().numpy()
            torch.cuda.synchronize()
            t["dbscan_labels_s"] = time.perf_counter() - t2
            res = m.api.DbscanResult(s_.cpu().numpy(), labels, c_.cpu().numpy())
            cm = m.cluster_membership(pr, res, 0.025, 0.01, min_members=50)  # notebook cell 7
            barrier()
            t["pipeline_s"] = time.perf_counter() - t0
            return t, X_d, gl, gw, a_, s_, c_, cm

        pipeline()
        m.check(m.lib().mcbb_pair_feature_counter(1, None, stream_ptr()))
        t, X_d, gl, gw, a_, s_, c_, cm = pipeline()
        pf = C.c_int64(0)
        m.check(m.lib().mcbb_pair_feature_counter(0, C.byref(pf), stream_ptr()))
        keys = ("h2d_solve_s", "allgather_s", "sort_normalize_hist_s", "dbscan_labels_s", "pipeline_s")
        tt = torch.tensor([t[k] for k in keys] + [float(pf.value)], dtype=torch.float64, device=dev)
        tsum = tt.clone()
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        pipe = {k: float(v) for k, v in zip(keys, tt.tolist())}
        pipe["n_runs"] = int(X_d.shape[1])
        pipe["what"] = ("wall clock, max over ranks: pinned H2D -> integrate+featurize (all measures) -> NCCL all-gather of "
                        "the measures -> on the device: sortperm by eps, normalize, common histogram edges (k_bin 0.25; "
                        "quartiles by a device sort), Float32 ECDF rows -> fused (row-sharded) DBSCAN(0.008, 10) -> labels "
                        "D2H -> cluster_membership")
        n_all, Ftot = int(X_d.shape[1]), int(X_d.shape[0])
        t_cl = pipe["dbscan_labels_s"]
        executed = float(tsum[5].item())
        cluster = {"n_runs": n_all, "time_s": t_cl, "eps": DB_EPS, "minpts": DB_MINPTS, "n_clusters": int(s_.numel()),
                   "n_noise": int((a_ == 0).sum().item()), "n_features": Ftot,
                   "feature_groups": [int(x) for x in gl], "largest_clusters": sorted([int(x) for x in c_.tolist()], reverse=True)[:6],
                   "membership_windows": int(cm.data.shape[0]), "labels_sha1": bench.labels_hash(a_),
                   "pairs_per_s": 0.5 * n_all * (n_all - 1) / t_cl,
                   "roofline": {"bound": "fp64_add", "unit": "TFLOP/s (FP64 add/sub)",
                                "achieved": 2.0 * executed / t_cl / world / 1e12, "peak": peak.value / 2.0,
                                "frac": 2.0 * executed / t_cl / world / 1e12 / (peak.value / 2.0),
                                "executed_pair_features": executed,
                                "brute_force_pair_features": 2.0 * 0.5 * n_all * (n_all - 1) * Ftot}}

    if rank == 0:
        cpu = None
        steps_ratio = None
        if not args.no_cpu and world == 1:
            from oracle import oracle as orc
            ns_ = args.cpu_sample
            srng = np.random.default_rng(4242)
            ics = np.empty((ns_, 2 * N))
            ics[:, 0::2] = srng.uniform(-1, 1, (ns_, N))
            ics[:, 1::2] = srng.uniform(-1, 1, (ns_, N))
            pars = srng.uniform(1.8, 2.1, (ns_, 1))
            threads = os.cpu_count() or 1
            t0 = time.perf_counter()
            ores = orc.solve_features(psys, opts, pfeat, ics, pars, ts, n_threads=threads)
            dt = time.perf_counter() - t0
            cpu = {"value": ns_ / dt, "unit": "trajectories/s", "cores": threads, "kind": "port",
                   "sample": "%d trajectories of the same C4 workload in %.1f s (oracle restatement of the reference's CPU "
                             "path; Julia itself is not installable here)" % (ns_, dt)}
            Rs = R
            # the same sample through the kernel: step attempts against the oracle's
            ic2 = torch.from_numpy(np.ascontiguousarray(ics.T)).to(dev)
            pa2 = torch.from_numpy(np.ascontiguousarray(pars.T)).to(dev)
            f2 = torch.empty((nm * nf, ns_), dtype=torch.float64, device=dev)
            m2 = torch.empty((nb, ns_), dtype=torch.float64, device=dev)
            r2 = torch.empty((ns_,), dtype=torch.int32, device=dev)
            n2 = torch.empty((2, ns_), dtype=torch.int64, device=dev)
            m.check(m.lib().mcbb_solve_features_dev(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), ns_, p(ic2), p(pa2),
                                                    p(ts_d), ts_d.numel(), p(f2), p(m2), C.c_void_p(0), p(r2), p(n2),
                                                    stream_ptr()))
            g = n2[0].double().cpu().numpy()
            o = ores["nsteps"][:, 0].astype(float)
            steps_ratio = {"gpu_over_oracle": float(g.sum() / o.sum()), "gpu_mean": float(g.mean()),
                           "oracle_mean": float(o.mean()), "runs": int(ns_)}
            del Rs
        line = {
            "metric": bench.METRIC, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "runs_per_gpu_per_step": R, "runs_per_step": world * R,
                       "l2": "flushed between timed steps"},
            "mean_step_attempts_per_run": float(nst.item()) / (world * R), "failed_runs": int(failed.item()),
            "steps_ratio": steps_ratio,
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "bitwise_equal_to_resident": same,
                    "host_memory": "pinned"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "pipeline": pipe, "pipeline_s": pipe["pipeline_s"] if pipe else None, "cluster": cluster,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
