#!/usr/bin/env python
"""bench.py — headline benchmark of the MCBB hot path on B200 (contract: see the task statement / DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference ...                      the CPU arm (the oracle restatement; Julia is absent)
    python bench.py --config C4 ...                           the Stuart-Landau pipeline (BASELINE.json configs[3])
    python bench.py --config C1|C2|C3 ...                     the small configs with SURVEY §8(d)'s inputs (single GPU)

Workload C5 (BASELINE.json configs[4]): Kuramoto network N=100, Erdos-Renyi p=0.2 adjacency, w~N(0.5,0.1),
ICs~U(-pi,pi)^100, K~U(0,8), tspan (0,1000), tail 0.9 (N_t=400), Tsit5 abstol 1e-6 / reltol 1e-3, measures
[mean, std] with cyclic_setback.  One STEP = integrate + featurize one batch of `--runs-per-gpu` trajectories per
GPU (default 125 000 = 10^6 / 8, so that --gpus 8 is the 10^6-run configuration; weak scaling).
`value` = trajectories/s with inputs resident in HBM; `e2e` = the same through the host C-ABI call
(mcbb_solve_features: pinned host buffers, H2D of ic/par and D2H of the features inside the timed region; also
measured with pageable numpy arrays, what a Julia `ccall` passes).  After the timed steps the whole pipeline is run
once more as ONE wall-clock region (`pipeline_s`: H2D -> solve -> feature all-gather -> sort -> eps from the
reference's KNN heuristic -> fused DBSCAN -> labels D2H), and the cluster stage gets its own roofline.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OSC = 100
TSPAN = (0.0, 1000.0)
TAIL = 0.9
N_T = 400
SINCOS_FLOPS = 40  # sincos_fast: 20 FMA-class FP64 instructions
METRIC = "trajectories/sec (integrate+featurize)"
WORKLOAD = ("C5 kuramoto_network N=100 ER p=0.2 (seed 10), w~N(0.5,0.1), ic~U(-pi,pi)^100, K~U(0,8), tspan (0,1000), "
            "tail 0.9, N_t=400, Tsit5 abstol 1e-6 reltol 1e-3, measures [mean,std] cyclic_setback")


def config_dict(world, runs_per_gpu):
    """Identical for both arms (the reference arm times a bounded sample of this configuration per step)."""
    return {"workload": WORKLOAD, "runs_per_gpu_per_step": runs_per_gpu, "runs_per_step": world * runs_per_gpu,
            "l2": "flushed between timed steps"}


def load_abi():
    """mcbb.jl_b200/_abi.py alone: ctypes declarations, loads no library (shared by both arms)."""
    import importlib.util
    name = "mcbb_jl_b200._abi"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "mcbb.jl_b200", "_abi.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def erdos_renyi(N, p, seed):
    rng = np.random.default_rng(seed)
    U = np.triu(rng.random((N, N)) < p, 1)
    return (U | U.T).astype(np.float64)


def make_system(abi):
    """The C5 system / featurizer / solver descriptors from the ABI declarations alone."""
    A = erdos_renyi(N_OSC, 0.2, 10)
    w = np.random.default_rng(11).normal(0.5, 0.1, N_OSC)
    psys = abi.PackedSystem(abi.SYS_KURAMOTO_NETWORK, N_OSC, N_OSC, [0.5], [0], mat_a=A, vec_a=w)
    pfeat = abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD], cyclic_setback=True)
    opts = abi.solver_opts(abi.ALG_TSIT5, TSPAN)
    return A, psys, pfeat, opts


def make_runs(n_runs, seed_offset):
    """Synthetic C5 initial conditions and parameters (numpy PCG64, seeded per rank / per sample)."""
    rng = np.random.default_rng(1000 + seed_offset)
    ic = np.asfortranarray(rng.uniform(-np.pi, np.pi, size=(n_runs, N_OSC)))
    par = np.asfortranarray(rng.uniform(0.0, 8.0, size=(n_runs, 1)))
    return ic, par


def make_inputs(m, n_runs, seed_offset):
    """Our arm: descriptors + runs + the product's tsave_array."""
    A, psys, pfeat, opts = make_system(m.abi)
    ic, par = make_runs(n_runs, seed_offset)
    n = C.c_int32(0)
    m.check(m.lib().mcbb_tsave_array(C.byref(opts), N_T, TAIL, m.abi.c_double_p(), C.byref(n)))
    ts = np.zeros(n.value)
    m.check(m.lib().mcbb_tsave_array(C.byref(opts), N_T, TAIL, ts.ctypes.data_as(m.abi.c_double_p), C.byref(n)))
    return A, ic, par, psys, pfeat, opts, ts


def algorithmic_flops(A, nsteps_total):
    """FP64 flops of the minimal (factored, zero-skipping) algorithm, DESIGN.md §5.  Per step attempt:
    6 RHS evaluations x (4 nnz [two mat-vecs, FMA = 2 flops] + N (sincos + 10)) + stage combinations 2*21 N
    + error estimate and new state 2*13 N + 6 N."""
    N = A.shape[0]
    nnz = int(np.count_nonzero(A))
    per_step = 6 * (4 * nnz + N * (SINCOS_FLOPS + 10)) + 42 * N + 32 * N
    per_step_dense = 6 * (4 * N * N + N * (SINCOS_FLOPS + 10)) + 42 * N + 32 * N
    return per_step * nsteps_total, per_step_dense * nsteps_total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(orc, psys, pfeat, opts, ts, n_sample, seed_offset=0):
    """The oracle restatement (literal N^2-sin formulas, OpenMP over trajectories) on a bounded sample.
    Returns (traj/s, seconds, threads, oracle result, ic, par)."""
    ic, par = make_runs(n_sample, seed_offset)
    threads = os.cpu_count() or 1  # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    t0 = time.perf_counter()
    res = orc.solve_features(psys, opts, pfeat, ic, par, ts, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, threads, res, ic, par


def run_reference(args):
    """--impl reference: Julia is not installed here or on the GPU box, so the reference arm is the CPU oracle
    (oracle/mcbb_oracle.c, a restatement of the reference's DifferentialEquations.jl + StatsBase path) on all
    host threads, on a bounded sample of the same workload per step.  Loads oracle/_ref/libmcbb_oracle.so and the
    ctypes declarations only: the product library is never touched on this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    abi = load_abi()
    from oracle import oracle as orc
    orc.build()
    n_sample = args.cpu_sample
    _, psys, pfeat, opts = make_system(abi)
    ts = orc.tsave_array(opts, N_T, TAIL)
    for w in range(args.warmup):
        cpu_baseline(orc, psys, pfeat, opts, ts, max(8, n_sample // 8), seed_offset=100 + w)
    t_tot, n_tot, threads = 0.0, 0, 1
    for k in range(args.steps):
        _, dt, threads, _, _, _ = cpu_baseline(orc, psys, pfeat, opts, ts, n_sample, seed_offset=k)
        t_tot += dt
        n_tot += n_sample
    value = n_tot / t_tot
    assert "mcbb_jl_b200._lib" not in sys.modules, "the reference arm must not load the product library"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, args.runs_per_gpu),
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": threads, "kind": "port",
                         "sample": "%d trajectories per step of the same C5 workload (oracle restatement of the "
                                   "reference's CPU path, all host threads; Julia itself is not installable here)"
                                   % n_sample},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def labels_hash(t):
    return hashlib.sha1(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C5", choices=["C5", "C4", "C3", "C2", "C1"])
    ap.add_argument("--runs-per-gpu", type=int, default=125000)
    ap.add_argument("--cpu-sample", type=int, default=384)
    ap.add_argument("--c4-runs", type=int, default=200000, help="--config C4: runs of the whole ensemble (all GPUs)")
    ap.add_argument("--no-cluster", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.config == "C4":
        from tools import bench_c4
        return bench_c4.main(args)
    if args.config in ("C1", "C2", "C3"):
        from tools import bench_small
        return bench_small.main(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from __graft_entry__ import load_package
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
        # NCCL prints its version banner on stdout at communicator creation: keep stdout for the single JSON line
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
    R = args.runs_per_gpu
    A, ic, par, psys, pfeat, opts, ts = make_inputs(m, R, rank)
    ic_d = torch.from_numpy(np.ascontiguousarray(ic.T)).to(dev)  # [n_dim, R]: run index contiguous
    par_d = torch.from_numpy(np.ascontiguousarray(par.T)).to(dev)
    ts_d = torch.from_numpy(ts).to(dev)
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream_ptr = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return m.dist.solve_features_device(psys, opts, pfeat, ic_d, par_d, ts_d)

    for _ in range(args.warmup):
        feat_d, _, ret_d, nsteps_d = step_resident()
    barrier()
    launches0 = m.kernel_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        l2_flush.fill_(k)  # flush L2 between timed iterations (inputs are 100 MB, the L2 is 126 MB)
        ev[k][0].record()
        feat_d, _, ret_d, nsteps_d = step_resident()
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = m.kernel_launch_count() - launches0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_total = torch.tensor([sum(ms_steps)], dtype=torch.float64, device=dev)
    nst = nsteps_d[0].sum().to(torch.float64).reshape(1)
    if dist is not None:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(nst, op=dist.ReduceOp.SUM)
    ms_total = float(ms_total.item())
    value = world * R * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host C-ABI (host buffers, H2D + D2H inside the timed region) --------------------
    nf = pfeat.n_filter(psys)
    dp, ip, lp = m.abi.c_double_p, m.abi.c_int32_p, m.abi.c_int64_p
    ic_h = torch.empty((N_OSC, R), dtype=torch.float64).pin_memory()
    ic_h.copy_(torch.from_numpy(np.ascontiguousarray(ic.T)))
    par_h = torch.empty((1, R), dtype=torch.float64).pin_memory()
    par_h.copy_(torch.from_numpy(np.ascontiguousarray(par.T)))
    feat_h = torch.empty((pfeat.n_meas * nf, R), dtype=torch.float64).pin_memory()
    ret_h = torch.empty((R,), dtype=torch.int32).pin_memory()
    ns_h = torch.empty((2, R), dtype=torch.int64).pin_memory()

    def call_host(ic_p, par_p, feat_p, ret_p, ns_p):
        m.check(m.lib().mcbb_solve_features(
            C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), R, C.cast(ic_p, dp), C.cast(par_p, dp),
            ts.ctypes.data_as(dp), len(ts), C.cast(feat_p, dp), dp(), dp(), C.cast(ret_p, ip), C.cast(ns_p, lp)))

    def time_host(fn, steps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * R * steps / float(t.item())

    e2e_steps = max(2, min(args.steps, 3))
    e2e_value = time_host(lambda: call_host(ic_h.data_ptr(), par_h.data_ptr(), feat_h.data_ptr(), ret_h.data_ptr(),
                                            ns_h.data_ptr()), e2e_steps)
    same = bool(torch.equal(torch.nan_to_num(feat_h), torch.nan_to_num(feat_d.cpu())))
    # pageable numpy arrays, exactly what a Julia ccall hands over (Array{Float64} is ordinary malloc'ed memory)
    feat_p = np.empty((pfeat.n_meas * nf, R))
    ret_p = np.empty(R, dtype=np.int32)
    ns_p = np.empty((2, R), dtype=np.int64)
    e2e_pageable = time_host(lambda: call_host(ic.ctypes.data, par.ctypes.data, feat_p.ctypes.data, ret_p.ctypes.data,
                                               ns_p.ctypes.data), e2e_steps)
    h2d = ic_h.numel() * 8 + par_h.numel() * 8 + len(ts) * 8
    d2h = feat_h.numel() * 8 + ret_h.numel() * 4 + ns_h.numel() * 8

    # ---- roofline of the dominant kernel (FP64 FMA pipe) -----------------------------------------------------
    peak = C.c_double(0.0)
    mhz = C.c_double(0.0)
    m.check(m.lib().mcbb_measure_fp64_fma_peak(C.byref(peak), C.byref(mhz)))
    fl_alg, fl_dense = algorithmic_flops(A, float(nst.item()) * args.steps)
    traffic_ncu = None
    traffic = None
    try:  # DRAM traffic of the same kernel from the committed ncu --set full capture (one launch of 125 000 runs)
        import csv
        src = next(p for p in ("r2_ncu_k_solve_batched_umma.csv", "r1_ncu_k_solve_batched_umma.csv")
                   if os.path.exists(os.path.join(ROOT, "profiles", p)))
        rows = {r[0]: r for r in csv.reader(open(os.path.join(ROOT, "profiles", src)))}
        unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
        ncu_runs = 125000
        traffic_ncu = {"runs_per_launch": ncu_runs, "source": "profiles/" + src,
                       "dram_bytes": float(rd[2]) * unit[rd[1]] + float(wr[2]) * unit[wr[1]],
                       "algorithmic_bytes": ncu_runs * (N_OSC * 8 + 8 + 2 * N_OSC * 8 + 4 + 16)}
        if R == ncu_runs:
            traffic = traffic_ncu["dram_bytes"]
    except Exception:
        pass
    achieved = fl_alg / (ms_total * 1e-3) / world / 1e12
    roofline = {
        "bound": "fp64_fma", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
        "frac": achieved / peak.value if peak.value > 0 else None, "traffic": traffic,
        "kernel": "k_solve_batched_umma (tcgen05.mma kind::i8, A and accumulators in TMEM)", "traffic_ncu": traffic_ncu,
        "peak_source": "measured in this run by the library's DFMA micro-benchmark (MEASURED_PEAKS.json has no FP64 "
                       "entry; nominal B200 ~37 TFLOP/s)",
        "achieved_dense_equiv": fl_dense / (ms_total * 1e-3) / world / 1e12,
        "note": "achieved = algorithmic FP64 flops of the zero-skipping factored algorithm (DESIGN.md §5) x the kernel's "
                "step attempts (checked against the oracle's: steps_ratio) / CUDA-event time of the kernel launches; "
                "the HBM roofline named by the contract does not bound this kernel (arithmetic intensity > 1e4 flop/B)",
    }
    evals = 6.0 * float(nst.item()) * args.steps / world  # per GPU in the timed region
    int8_ops = evals * (2.0 * 128 * 128 * 16)             # 16 accumulator columns per trajectory
    roofline["tensor"] = {
        "bound": "tensor", "achieved": int8_ops / (ms_total * 1e-3) / 1e12, "peak": 8192 * 2 * 148 * 1.965e9 / 1e12,
        "unit": "TOP/s (int8)", "frac": int8_ops / (ms_total * 1e-3) / (8192 * 2 * 148 * 1.965e9),
        "peak_source": "tcgen05.mma kind::i8 M=128: 8192 MAC/clk/SM measured by profiles/microbench/umma_probe.cu, x 148 SMs x 1.965 GHz",
        "note": "the tensor cores replace 54 % of the algorithmic FP64 flops; they are not the bound of the kernel",
    }

    # ---- the whole pipeline as one wall-clock region + the cluster stage's own numbers ------------------------
    gl = np.array([nf, nf, 1], dtype=np.int32)
    gw = np.array([1.0, 0.5, 1.0])
    minpts = 20

    def knn_eps(X):
        """median(KNN_dist_relative(D, 0.005)) (src/eval_clustering.jl:1526-1553, docs/src/basic_usage.md:72) on a
        10^4-run subsample, identical on every rank."""
        n_all = X.shape[1]
        ms_ = min(10000, n_all)
        sub = torch.from_numpy(np.sort(np.random.default_rng(7).choice(n_all, size=ms_, replace=False))).to(dev)
        Xs = X[:, sub].contiguous()
        K = max(1, int(round(0.005 * ms_)))
        cum = torch.empty((ms_,), dtype=torch.float64, device=dev)
        kth = torch.empty((ms_,), dtype=torch.float64, device=dev)
        # k = minpts + 1: the distance to the minpts-th OTHER point (the zero self-distance is rank 1)
        m.check(m.lib().mcbb_knn_dist_features_dev(C.c_void_p(Xs.data_ptr()), ms_, 3, gl.ctypes.data_as(ip),
                                                   gw.ctypes.data_as(dp), minpts + 1, K, 0, C.c_void_p(kth.data_ptr()),
                                                   C.c_void_p(cum.data_ptr()), stream_ptr()))
        knn_eps.kdist = float(kth.median().item())  # the docs' other heuristic: k-dist graph (basic_usage.md:64-66)
        return float(cum.median().item()), K, ms_

    def pipeline(trace=None):
        """H2D -> solve -> all-gather -> sort by parameter -> eps -> fused (row-sharded) DBSCAN -> labels D2H."""
        t = {}
        barrier()
        t0 = time.perf_counter()
        icd = ic_h.to(dev, non_blocking=True)
        pard = par_h.to(dev, non_blocking=True)
        f_d, _, _, _ = m.dist.solve_features_device(psys, opts, pfeat, icd, pard, ts_d)
        Xl = torch.cat([f_d, pard], dim=0)
        torch.cuda.synchronize()
        t["h2d_solve_s"] = time.perf_counter() - t0
        if dist is not None:
            parts = [torch.empty_like(Xl) for _ in range(world)]
            dist.all_gather(parts, Xl)
            X = torch.cat(parts, dim=1)
            torch.cuda.synchronize()
        else:
            X = Xl
        t["allgather_s"] = time.perf_counter() - t0 - t["h2d_solve_s"]
        X = torch.nan_to_num(X, nan=1e30)
        # MCBB.solve sorts the runs by parameter before any clustering (sort!, setup_mc_prob.jl:520-536):
        # stable sort on the parameter row, identical on every rank
        order = torch.sort(X[-1], stable=True).indices
        X = X[:, order].contiguous()
        eps, K, n_sub = knn_eps(X)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        t["sort_eps_s"] = t1 - t0 - t["h2d_solve_s"] - t["allgather_s"]
        if world > 1:
            a_, s_, c_ = m.dist.dbscan_features_sharded(X, gl, gw, eps, minpts, rank, world, dist)
        else:
            a_, s_, c_ = m.dist.dbscan_features_device(X, gl, gw, eps, minpts)
        labels = a_.cpu()
        torch.cuda.synchronize()
        t["dbscan_labels_s"] = time.perf_counter() - t1
        barrier()
        t["pipeline_s"] = time.perf_counter() - t0
        return t, X, eps, K, n_sub, a_, s_, c_, labels

    cluster = None
    pipe = None
    if not args.no_cluster:
        pipeline()  # first pass allocates the DBSCAN scratch (reported separately as the cold run would mislead)
        m.check(m.lib().mcbb_pair_feature_counter(1, None, stream_ptr()))
        t, X, eps, K, n_sub, a_, s_, c_, labels = pipeline()
        pf = C.c_int64(0)
        m.check(m.lib().mcbb_pair_feature_counter(0, C.byref(pf), stream_ptr()))
        # the counter also saw the 10^4-run eps subsample (materialised blocks): n_sub^2/2-ish pairs x F, negligible
        tt = torch.tensor([t[k] for k in ("h2d_solve_s", "allgather_s", "sort_eps_s", "dbscan_labels_s", "pipeline_s")]
                          + [float(pf.value)], dtype=torch.float64, device=dev)
        tsum = tt.clone()
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        n_all = int(X.shape[1])
        Ftot = int(X.shape[0])
        t_cl = float(tt[3].item())
        executed = float(tsum[5].item())  # pair x feature terms over all ranks
        dadd_peak = peak.value / 2.0  # T instr/s: a DADD occupies the pipe like a DFMA and counts one flop
        pipe = {"pipeline_s": float(tt[4].item()), "h2d_solve_s": float(tt[0].item()),
                "allgather_s": float(tt[1].item()), "sort_eps_s": float(tt[2].item()),
                "dbscan_labels_s": t_cl, "n_runs": n_all,
                "what": "wall clock, max over ranks: pinned H2D of ic/par -> integrate+featurize -> NCCL all-gather -> "
                        "stable sort by parameter -> eps = median(KNN_dist_relative) on a 10^4 subsample -> fused "
                        "(row-sharded) DBSCAN -> labels D2H"}
        cluster = {"n_runs": n_all, "time_s": t_cl, "eps": eps, "eps_rule": "median(KNN_dist_relative(D,0.005)) on "
                   "%d runs: cumulative distance to the K=%d nearest" % (n_sub, K), "minpts": minpts,
                   "n_clusters": int(s_.numel()), "n_noise": int((a_ == 0).sum().item()), "n_features": Ftot,
                   "largest_clusters": sorted([int(x) for x in c_.tolist()], reverse=True)[:5],
                   "pairs_per_s": 0.5 * n_all * (n_all - 1) / t_cl, "labels_sha1": labels_hash(a_),
                   "roofline": {"bound": "fp64_add", "unit": "TFLOP/s (FP64 add/sub)",
                                "achieved": 2.0 * executed / t_cl / world / 1e12, "peak": dadd_peak,
                                "frac": 2.0 * executed / t_cl / world / 1e12 / dadd_peak,
                                "executed_pair_features": executed,
                                "brute_force_pair_features": 2.0 * 0.5 * n_all * (n_all - 1) * Ftot,
                                "note": "2 FP64 adds (x_i - x_j, accumulate |.|) per executed pair x feature term, counted "
                                        "by the tiles themselves (early exits and saturation skips excluded), over the "
                                        "wall time of the whole DBSCAN incl. compaction, sorts and the labels D2H; "
                                        "peak = DFMA peak / 2"}}
        # The cumulative-KNN eps is as large as the integration noise of this workload: everything is one
        # cluster and the saturation skips remove most of the pair space.  Second clustering on the record with the docs'
        # k-dist heuristic (median distance to the minpts-th neighbour on the same subsample): a tight eps, most points
        # noise or small clusters, NO saturation — the expensive case for the tiles.
        eps2 = knn_eps.kdist
        m.check(m.lib().mcbb_pair_feature_counter(1, None, stream_ptr()))
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            a3, s3, c3 = m.dist.dbscan_features_sharded(X, gl, gw, eps2, minpts, rank, world, dist)
        else:
            a3, s3, c3 = m.dist.dbscan_features_device(X, gl, gw, eps2, minpts)
        barrier()
        t3 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        pf3 = C.c_int64(0)
        m.check(m.lib().mcbb_pair_feature_counter(0, C.byref(pf3), stream_ptr()))
        ex3 = torch.tensor([float(pf3.value)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            dist.all_reduce(ex3, op=dist.ReduceOp.SUM)
        cluster["kdist_eps"] = {"eps": eps2, "eps_rule": "median distance to the minpts-th neighbour (k_dist) on the same subsample",
                                "time_s": float(t3.item()), "n_clusters": int(s3.numel()), "n_noise": int((a3 == 0).sum().item()),
                                "largest_clusters": sorted([int(x) for x in c3.tolist()], reverse=True)[:5],
                                "pairs_per_s": 0.5 * n_all * (n_all - 1) / float(t3.item()), "labels_sha1": labels_hash(a3),
                                "roofline_frac_fp64_add": 2.0 * float(ex3.item()) / float(t3.item()) / world / 1e12 / dadd_peak,
                                "executed_pair_features": float(ex3.item())}
        if world == 1:
            # the same k-dist heuristic on ALL runs (no subsample) through the fused per-row K-smallest kernel (csrc/topk.cuh)
            kth_all = torch.empty((n_all,), dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.check(m.lib().mcbb_knn_dist_features_dev(C.c_void_p(X.data_ptr()), n_all, 3, gl.ctypes.data_as(ip),
                                                       gw.ctypes.data_as(dp), minpts + 1, 1, 0, C.c_void_p(kth_all.data_ptr()),
                                                       C.c_void_p(0), stream_ptr()))
            torch.cuda.synchronize()
            cluster["kdist_all_runs"] = {"time_s": time.perf_counter() - t0, "k": minpts + 1,
                                         "median": float(kth_all.median().item()), "n_runs": n_all,
                                         "pairs_per_s": float(n_all) * n_all / (time.perf_counter() - t0)}
            del kth_all
            # worst case for the early exits: the same points in random order (no locality along the index)
            perm = torch.from_numpy(np.random.default_rng(3).permutation(n_all)).to(dev)
            Xs = X[:, perm].contiguous()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a2, s2, c2 = m.dist.dbscan_features_device(Xs, gl, gw, eps, minpts)
            torch.cuda.synchronize()
            cluster["shuffled_order_time_s"] = time.perf_counter() - t0
            cluster["shuffled_same_partition"] = bool(int(s2.numel()) == int(s_.numel())
                                                      and int((a2 == 0).sum()) == int((a_ == 0).sum()))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a4, s4, c4 = m.dist.dbscan_features_device(Xs, gl, gw, eps2, minpts)
            torch.cuda.synchronize()
            cluster["kdist_eps"]["shuffled_order_time_s"] = time.perf_counter() - t0
            cluster["kdist_eps"]["shuffled_same_partition"] = bool(int(s4.numel()) == int(s3.numel())
                                                                   and int((a4 == 0).sum()) == int((a3 == 0).sum()))
            # materialised distance matrix of 20 000 points at the two HBM-bound feature counts of the BASELINE configs
            # (C1: F = 19 in groups [6,6,6,1]; C2: F = 4) — synthetic features, the kernel does not care — and at this
            # workload's own F = 201 (FP64-add bound: 2 F adds per 16 bytes written)
            nm = min(20000, n_all)
            Dm = torch.empty((nm, nm), dtype=torch.float64, device=dev)
            hbm_peak = 6551.7
            try:
                hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except Exception:
                pass
            cluster["materialize"] = {"bound": "hbm", "n": nm, "peak": hbm_peak, "unit": "GB/s",
                                      "note": "8 n^2 bytes written / best of 3 launches (mcbb_distance_matrix_dev); peak = "
                                              "measured HBM copy bandwidth (MEASURED_PEAKS.json)"}
            for name, Xm, glm, gwm in (("C1_F19", torch.randn((19, nm), dtype=torch.float64, device=dev), [6, 6, 6, 1], [1, .5, .5, 1]),
                                       ("C2_F4", torch.randn((4, nm), dtype=torch.float64, device=dev), [1, 1, 1, 1], [1, .75, .5, 1]),
                                       ("C5_F201", X[:, :nm].contiguous(), list(gl), list(gw))):
                glm_, gwm_ = np.asarray(glm, dtype=np.int32), np.asarray(gwm, dtype=np.float64)
                best = None
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    m.check(m.lib().mcbb_distance_matrix_dev(C.c_void_p(Xm.data_ptr()), nm, len(glm), glm_.ctypes.data_as(ip),
                                                             gwm_.ctypes.data_as(dp), C.c_void_p(Dm.data_ptr()), stream_ptr()))
                    e1.record()
                    torch.cuda.synchronize()
                    ms_ = e0.elapsed_time(e1)
                    best = ms_ if best is None else min(best, ms_)
                gbs = 8.0 * nm * nm / (best * 1e-3) / 1e9
                cluster["materialize"][name] = {"ms": best, "achieved": gbs, "frac": gbs / hbm_peak}
            del Dm
        else:
            # sharded labels == single-GPU fused labels on a 2*10^5-run subset (every rank runs the sharded stages,
            # rank 0 alone the single-GPU DBSCAN)
            ns_ = min(200000, n_all)
            Xsub = X[:, :ns_].contiguous()
            a_sh, s_sh, c_sh = m.dist.dbscan_features_sharded(Xsub, gl, gw, eps, minpts, rank, world, dist)
            if rank == 0:
                a_1, s_1, c_1 = m.dist.dbscan_features_device(Xsub, gl, gw, eps, minpts)
                same_lab = bool(torch.equal(a_sh, a_1) and torch.equal(s_sh, s_1) and torch.equal(c_sh, c_1))
                cluster["sharded_equals_single_gpu"] = {"n_runs": ns_, "equal": same_lab,
                                                        "sha1_sharded": labels_hash(a_sh), "sha1_single": labels_hash(a_1)}
                assert same_lab, "row-sharded DBSCAN labels differ from the single-GPU fused DBSCAN"
            barrier()

    if rank == 0:
        cpu = None
        steps_ratio = None
        if not args.no_cpu and world == 1:  # reported at N=1 only (rank 0)
            from oracle import oracle as orc
            v, dt, threads, ores, ic_s, par_s = cpu_baseline(orc, psys, pfeat, opts, ts, args.cpu_sample)
            cpu = {"value": v, "unit": "trajectories/s", "cores": threads, "kind": "port",
                   "sample": "%d trajectories of the same C5 workload in %.1f s (oracle restatement of the reference's "
                             "CPU path; Julia itself is not installable here)" % (args.cpu_sample, dt)}
            # the same sample through the GPU kernel: its step attempts must be the oracle's (roofline.achieved
            # multiplies algorithmic flops per step by the KERNEL's count)
            ics_d = torch.from_numpy(np.ascontiguousarray(ic_s.T)).to(dev)
            pars_d = torch.from_numpy(np.ascontiguousarray(par_s.T)).to(dev)
            _, _, _, ns_s = m.dist.solve_features_device(psys, opts, pfeat, ics_d, pars_d, ts_d)
            g = ns_s[0].double().cpu().numpy()
            o = ores["nsteps"][:, 0].astype(float)
            steps_ratio = {"gpu_over_oracle": float(g.sum() / o.sum()), "gpu_mean": float(g.mean()),
                           "oracle_mean": float(o.mean()), "runs": int(len(o)),
                           "identical_runs_frac": float((g == o).mean())}
        line = {
            "metric": METRIC, "value": value, "unit": "trajectories/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world, R),
            "mean_step_attempts_per_run": float(nst.item()) / (world * R), "steps_ratio": steps_ratio,
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "bitwise_equal_to_resident": same,
                    "host_memory": "pinned", "pageable_value": e2e_pageable},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "pipeline": pipe, "pipeline_s": pipe["pipeline_s"] if pipe else None,
            "cluster": cluster, "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
