"""bench.py --config C1 | C2 | C3 — the three small BASELINE configs with SURVEY §8(d)'s inputs, same JSON contract as the
headline bench: `value` = trajectories/s of integrate+featurize with the inputs resident in HBM (K timed steps, CUDA
events), `e2e` through the host C-ABI call with host buffers, `pipeline` = solve -> distance -> DBSCAN -> cluster_membership
once as wall clock with §8(d)'s eps / minpts, `cpu_baseline` = the oracle on a bounded sample of the same runs.

  C1  kuramoto_network N=6, A=ones, w~N(0.5,0.01), 500 ICs~U(-pi,pi)^6 x K in {1,1.5,2,2.5,3} = 2500 runs, tspan (0,40),
      tail 0.9, Tsit5 defaults, measures [mean,std,KL-hist], weights [1,.5,.5,1], eps = median(KNN_dist_relative), minpts 4
      (test/ode_prob_example.jl:36-41, 75-76)
  C2  logistic map, 10^4 runs, r~U(2.5,4), x0~U(0.1,0.99), tspan (0,1000), tail 0.8, measures [mean,std,KL], weights
      [1,.75,.5,1], eps 0.04, minpts 4 (docs/src/basic_usage.md:26-78)
  C3  second_order_kuramoto N=10 on a random 3-regular graph (seed 5), drive +-1, damping 0.1, coupling = 10 i / N_mc,
      10^5 runs, tspan (0,1000), tail 0.9, state_filter N+1:2N, measures [mean,std], weights [1,0,1], eps 5,
      minpts max(5, N_mc/200), DBSCAN on the fly (paper/kuramoto/1-simulate.jl:22-27, 61; 2-analyze.jl:70-75)"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKLOADS = {
    "C1": "C1 kuramoto_network N=6 A=ones, 500 ICs x K in {1,1.5,2,2.5,3} = 2500 runs, tspan (0,40), tail 0.9, Tsit5 "
          "abstol 1e-6 reltol 1e-3, measures [mean,std,KL-hist]",
    "C2": "C2 logistic map, 10^4 runs r~U(2.5,4), tspan (0,1000), tail 0.8 (200 saved), FunctionMap, measures [mean,std,KL-hist]",
    "C3": "C3 second_order_kuramoto N=10 random 3-regular (seed 5), coupling 10 i/N_mc, 10^5 runs, tspan (0,1000), tail 0.9, "
          "Tsit5 abstol 1e-6 reltol 1e-3, measures [mean,std] of the frequencies",
}


def build(m, cases, cfg):
    if cfg == "C1":
        prob, weights = cases.c1_problem(m, n_ics=500, seed=1)
        return prob, weights, None, 4, True
    if cfg == "C2":
        prob, weights = cases.c2_problem(m, n_mc=10000, seed=3)
        return prob, weights, 0.04, 4, True
    n = 100000
    prob, weights = cases.c3_problem(m, n_mc=8, seed=5, tspan=(0.0, 1000.0))
    rng = np.random.default_rng(6)
    prob.ic = np.asfortranarray(rng.uniform(-np.pi, np.pi, size=(n, 20)))
    prob.par = np.asfortranarray((10.0 * np.arange(1, n + 1) / n)[:, None])
    prob.N_mc = n
    return prob, weights, 5.0, max(5, n // 200), False


def flops_per_step(cfg, prob):
    """Algorithmic FP64 flops per Tsit5 step attempt of one trajectory (sin = 20, sincos = 40 flops)."""
    if cfg == "C1":
        N, nnz = 6, 36
        return 6 * (4 * nnz + N * 50) + 74 * N
    if cfg == "C3":
        N, E = 10, 15
        return 6 * (E * 24 + 4 * N) + 74 * 2 * N
    return 3.0  # logistic: r x (1 - x), per iteration


def main(args):
    import torch
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    from tests import cases
    import bench
    m = load_package()
    cfg = args.config
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--config %s is a single-GPU configuration" % cfg)
    m.check(m.lib().mcbb_init(0))
    dev = torch.device("cuda:0")
    prob, weights, eps_fixed, minpts, materialize = build(m, cases, cfg)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    n = prob.N_mc
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    alg = m.abi.ALG_FUNCTIONMAP if prob.p.discrete else m.abi.ALG_TSIT5
    opts = m.abi.solver_opts(alg, prob.p.tspan)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    nf = pfeat.n_filter(psys)
    ic_d = torch.from_numpy(np.ascontiguousarray(prob.ic.T)).to(dev)
    par_d = torch.from_numpy(np.ascontiguousarray(prob.par.T)).to(dev)
    ts_d = torch.from_numpy(ts).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        out = m.dist.solve_features_device(psys, opts, pfeat, ic_d, par_d, ts_d)
    torch.cuda.synchronize()
    launches0 = m.kernel_launch_count()
    sampler = bench.ClockSampler(0)
    sampler.start()
    ms = []
    for k in range(args.steps):
        flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        feat_d, _, ret_d, ns_d = m.dist.solve_features_device(psys, opts, pfeat, ic_d, par_d, ts_d)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    clocks = sampler.stop()
    launches = m.kernel_launch_count() - launches0
    ms_total = sum(ms)
    value = n * args.steps / (ms_total * 1e-3)
    nst = float(ns_d[0].sum().item())

    # ---- e2e through the host entry -------------------------------------------------------------------------
    dp, ip, lp = m.abi.c_double_p, m.abi.c_int32_p, m.abi.c_int64_p
    ic_h = torch.from_numpy(np.ascontiguousarray(prob.ic.T)).pin_memory()
    par_h = torch.from_numpy(np.ascontiguousarray(prob.par.T)).pin_memory()
    feat_h = torch.empty((pfeat.n_meas * nf, n), dtype=torch.float64).pin_memory()
    ret_h = torch.empty((n,), dtype=torch.int32).pin_memory()
    ns_h = torch.empty((2, n), dtype=torch.int64).pin_memory()

    def call_host():
        m.check(m.lib().mcbb_solve_features(
            C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n, C.cast(ic_h.data_ptr(), dp), C.cast(par_h.data_ptr(), dp),
            ts.ctypes.data_as(dp), len(ts), C.cast(feat_h.data_ptr(), dp), dp(), dp(), C.cast(ret_h.data_ptr(), ip),
            C.cast(ns_h.data_ptr(), lp)))
    call_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        call_host()
    torch.cuda.synchronize()
    e2e_value = n * e2e_steps / (time.perf_counter() - t0)
    same = bool(torch.equal(torch.nan_to_num(feat_h), torch.nan_to_num(feat_d.cpu())))

    peak = C.c_double(0.0)
    mhz = C.c_double(0.0)
    m.check(m.lib().mcbb_measure_fp64_fma_peak(C.byref(peak), C.byref(mhz)))
    fl = flops_per_step(cfg, prob) * nst * args.steps
    achieved = fl / (ms_total * 1e-3) / 1e12
    roofline = {"bound": "fp64_fma" if cfg != "C2" else "latency", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved / peak.value if peak.value > 0 else None, "traffic": None,
                "kernel": ("k_solve_lane (one trajectory per lane, state in shared memory)" if cfg == "C2" else
                           "k_solve_group (a group of 8 / 16 lanes per trajectory, one oscillator per lane, stage vectors in "
                           "registers; bitwise equal to the lane kernel)"),
                "note": "algorithmic FP64 flops per step attempt (tools/bench_small.py::flops_per_step) x the kernel's step "
                        "count / CUDA-event time.  C1 and C2 are far too small to fill the machine for long: C1 is 2500 "
                        "trajectories for 40 time units (31 step attempts each), C2 is 10^4 x 1000 three-flop iterations — "
                        "launch and tail latency, not a pipe, bound them"}

    # ---- the pipeline once as wall clock (through the host mirror, like a user of the reference would) -----------
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.solve(prob, flag_check_inf_nan=False)  # warm (sorts prob in place; runs are re-sorted identically below)
        prob.ic, prob.par = ic0.copy(), par0.copy()
        t0 = time.perf_counter()
        sol = m.solve(prob, flag_check_inf_nan=False)
        t_solve = time.perf_counter() - t0
        t1 = time.perf_counter()
        if materialize:
            D = m.distance_matrix(sol, prob, weights)
            t_dist = time.perf_counter() - t1
            eps = eps_fixed if eps_fixed is not None else float(np.median(m.KNN_dist_relative(D)))
            t2 = time.perf_counter()
            db = m.dbscan(D, eps, minpts)
        else:
            D = m.distance_matrix(sol, prob, weights, materialize=False)
            t_dist = time.perf_counter() - t1
            eps = eps_fixed
            t2 = time.perf_counter()
            db = m.dbscan(D, eps, minpts)
        t_db = time.perf_counter() - t2
        window = {"C1": (0.5, 0.25), "C2": (0.005, 0.001), "C3": (0.5, 0.1)}[cfg]
        cm = m.cluster_membership(prob, db, *window)
        t_pipe = time.perf_counter() - t0
    pipe = {"pipeline_s": t_pipe, "solve_s": t_solve, "distance_s": t_dist, "dbscan_s": t_db, "materialised": materialize,
            "what": "host mirror: solve (H2D, integrate+featurize, D2H, sort) -> distance_matrix -> dbscan -> cluster_membership"}
    cluster = {"n_runs": n, "eps": eps, "minpts": minpts, "n_clusters": int(len(db.seeds)),
               "n_noise": int(m.cluster_n_noise(db)), "largest_clusters": sorted([int(x) for x in db.counts], reverse=True)[:5],
               "membership_windows": int(cm.data.shape[0]), "time_s": t_dist + t_db}

    cpu = None
    steps_ratio = None
    if not args.no_cpu:
        from oracle import oracle as orc
        ns_ = min(n, {"C1": 2500, "C2": 10000, "C3": 1024}[cfg])
        sel = np.linspace(0, n - 1, ns_).astype(np.int64)
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        ores = orc.solve_features(psys, opts, pfeat, ic0[sel], par0[sel], ts, n_threads=threads)
        dt = time.perf_counter() - t0
        cpu = {"value": ns_ / dt, "unit": "trajectories/s", "cores": threads, "kind": "port",
               "sample": "%d of the %d runs in %.2f s (oracle restatement of the reference's CPU path; Julia itself is not "
                         "installable here)" % (ns_, n, dt)}
        feat0, _, ret0, ns0 = m.dist.solve_features_device(
            psys, opts, pfeat, torch.from_numpy(np.ascontiguousarray(ic0[sel].T)).to(dev),
            torch.from_numpy(np.ascontiguousarray(par0[sel].T)).to(dev), ts_d)
        g, o = ns0[0].double().cpu().numpy(), ores["nsteps"][:, 0].astype(float)
        steps_ratio = {"gpu_over_oracle": float(g.sum() / o.sum()), "identical_runs_frac": float((g == o).mean()), "runs": int(ns_)}
    line = {
        "metric": bench.METRIC, "value": value, "unit": "trajectories/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg], "runs_per_gpu_per_step": n, "runs_per_step": n, "l2": "flushed between timed steps"},
        "mean_step_attempts_per_run": nst / n, "failed_runs": int((ret_d != 0).sum().item()), "steps_ratio": steps_ratio,
        "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": int((ic_h.numel() + par_h.numel() + len(ts)) * 8),
                "d2h_bytes_per_step": int(feat_h.numel() * 8 + ret_h.numel() * 4 + ns_h.numel() * 8), "steps": e2e_steps,
                "bitwise_equal_to_resident": same, "host_memory": "pinned"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "pipeline": pipe,
        "pipeline_s": t_pipe, "cluster": cluster,
    }
    print(json.dumps(line))
