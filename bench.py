#!/usr/bin/env python
"""bench.py — headline benchmark of the MCBB hot path on B200 (contract: see the task statement / DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference ...                      the CPU arm (the oracle restatement; Julia is absent)

Workload (BASELINE.json configs[4], "C5"): Kuramoto network N=100, Erdos-Renyi p=0.2 adjacency, w~N(0.5,0.1),
ICs~U(-pi,pi)^100, K~U(0,8), tspan (0,1000), tail 0.9 (N_t=400), Tsit5 abstol 1e-6 / reltol 1e-3, measures
[mean, std] with cyclic_setback.  One STEP = integrate + featurize one batch of `--runs-per-gpu` trajectories per
GPU (default 125 000 = 10^6 / 8, so that --gpus 8 is the 10^6-run configuration; weak scaling).
`value` = trajectories/s with inputs resident in HBM; `e2e` = the same through the host C-ABI call
(mcbb_solve_features: pinned host buffers, H2D of ic/par and D2H of the features inside the timed region).
After the timed steps the features are all-gathered and DBSCAN-clustered once (fused, row-sharded): `cluster`.
"""
import argparse
import ctypes as C
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


def make_inputs(m, n_runs, seed_offset):
    """Synthetic C5 inputs for one rank (numpy PCG64, seeded per rank)."""
    from tests import cases
    A = cases.erdos_renyi(N_OSC, 0.2, 10)
    w = np.random.default_rng(11).normal(0.5, 0.1, N_OSC)
    pars = m.kuramoto_network_parameters(0.5, w, N_OSC, A)
    rng = np.random.default_rng(1000 + seed_offset)
    ic = np.asfortranarray(rng.uniform(-np.pi, np.pi, size=(n_runs, N_OSC)))
    par = np.asfortranarray(rng.uniform(0.0, 8.0, size=(n_runs, 1)))
    psys = pars.pack(["K"])
    pfeat = m.EvalOdeRun(eval_funcs=(m.mean, m.std), cyclic_setback=True).pack()
    opts = m.abi.solver_opts(m.abi.ALG_TSIT5, TSPAN)
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


def cpu_baseline(m, A, psys, pfeat, opts, ts, n_sample, seed_offset=0):
    """The oracle restatement (literal N^2-sin formulas, OpenMP over trajectories) on a bounded sample."""
    from oracle import oracle as orc
    rng = np.random.default_rng(1000 + seed_offset)
    ic = np.asfortranarray(rng.uniform(-np.pi, np.pi, size=(n_sample, N_OSC)))
    par = np.asfortranarray(rng.uniform(0.0, 8.0, size=(n_sample, 1)))
    threads = os.cpu_count() or 1  # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    t0 = time.perf_counter()
    orc.solve_features(psys, opts, pfeat, ic, par, ts, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, threads


def run_reference(args):
    """--impl reference: Julia is not installed here or on the GPU box, so the reference arm is the CPU oracle
    (oracle/mcbb_oracle.c, a restatement of the reference's DifferentialEquations.jl + StatsBase path) on all
    host threads, on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from __graft_entry__ import load_package
    m = load_package()
    from oracle import oracle as orc
    orc.build()
    n_sample = args.cpu_sample
    A, _, _, psys, pfeat, opts, ts = make_inputs(m, 8, 0)
    for w in range(args.warmup):
        cpu_baseline(m, A, psys, pfeat, opts, ts, max(8, n_sample // 8), seed_offset=100 + w)
    t_tot, n_tot, threads = 0.0, 0, 1
    for k in range(args.steps):
        v, dt, threads = cpu_baseline(m, A, psys, pfeat, opts, ts, n_sample, seed_offset=k)
        t_tot += dt
        n_tot += n_sample
    value = n_tot / t_tot
    line = {
        "impl": "reference", "metric": "trajectories/sec (integrate+featurize)", "value": value,
        "unit": "trajectories/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5 kuramoto_network N=100 ER p=0.2, tspan (0,1000), tail 0.9, Tsit5 1e-6/1e-3, "
                               "measures [mean,std] cyclic_setback", "runs_per_step": n_sample},
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": threads, "kind": "port",
                         "sample": "%d trajectories per step of the same C5 workload (oracle restatement, not Julia)"
                                   % n_sample},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--runs-per-gpu", type=int, default=125000)
    ap.add_argument("--cpu-sample", type=int, default=384)
    ap.add_argument("--no-cluster", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
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

    # ---- end to end through the host C-ABI (pinned host buffers, H2D + D2H inside the timed region) ----------
    ic_h = torch.from_numpy(ic).pin_memory() if False else torch.empty((N_OSC, R), dtype=torch.float64).pin_memory()
    ic_h.copy_(torch.from_numpy(np.ascontiguousarray(ic.T)))
    par_h = torch.empty((1, R), dtype=torch.float64).pin_memory()
    par_h.copy_(torch.from_numpy(np.ascontiguousarray(par.T)))
    nf = pfeat.n_filter(psys)
    feat_h = torch.empty((pfeat.n_meas * nf, R), dtype=torch.float64).pin_memory()
    ret_h = torch.empty((R,), dtype=torch.int32).pin_memory()
    ns_h = torch.empty((2, R), dtype=torch.int64).pin_memory()
    dp, ip, lp = m.abi.c_double_p, m.abi.c_int32_p, m.abi.c_int64_p

    def step_e2e():
        m.check(m.lib().mcbb_solve_features(
            C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), R, C.cast(ic_h.data_ptr(), dp),
            C.cast(par_h.data_ptr(), dp), ts.ctypes.data_as(dp), len(ts), C.cast(feat_h.data_ptr(), dp), dp(), dp(),
            C.cast(ret_h.data_ptr(), ip), C.cast(ns_h.data_ptr(), lp)))

    step_e2e()
    barrier()
    e2e_steps = max(2, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * R * e2e_steps / float(t_e2e.item())
    same = bool(torch.equal(torch.nan_to_num(feat_h), torch.nan_to_num(feat_d.cpu())))
    h2d = ic_h.numel() * 8 + par_h.numel() * 8 + len(ts) * 8
    d2h = feat_h.numel() * 8 + ret_h.numel() * 4 + ns_h.numel() * 8

    # ---- roofline of the dominant kernel (k_solve_batched: FP64 FMA pipe) ------------------------------------
    peak = C.c_double(0.0)
    mhz = C.c_double(0.0)
    m.check(m.lib().mcbb_measure_fp64_fma_peak(C.byref(peak), C.byref(mhz)))
    fl_alg, fl_dense = algorithmic_flops(A, float(nst.item()) * args.steps)
    traffic_ncu = None
    traffic = None
    try:  # DRAM traffic of the same kernel from the committed ncu --set full capture (one launch of 125 000 runs)
        import csv
        rows = {r[0]: r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r1_ncu_k_solve_batched_umma.csv")))}
        unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
        ncu_runs = 125000
        traffic_ncu = {"runs_per_launch": ncu_runs, "dram_bytes": float(rd[2]) * unit[rd[1]] + float(wr[2]) * unit[wr[1]],
                       "algorithmic_bytes": ncu_runs * (N_OSC * 8 + 8 + 2 * N_OSC * 8 + 4 + 16)}
        if R == ncu_runs:
            traffic = traffic_ncu["dram_bytes"]
    except Exception:
        pass
    achieved = fl_alg / (ms_total * 1e-3) / world / 1e12
    roofline = {
        "bound": "fp64_fma", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
        "frac": achieved / peak.value if peak.value > 0 else None, "traffic": traffic,
        "kernel": "k_solve_batched_umma<6,4> (tcgen05.mma kind::i8, A and accumulators in TMEM)", "traffic_ncu": traffic_ncu,
        "peak_source": "measured in this run by the library's DFMA micro-benchmark (MEASURED_PEAKS.json has no FP64 "
                       "entry; nominal B200 ~37 TFLOP/s)",
        "achieved_dense_equiv": fl_dense / (ms_total * 1e-3) / world / 1e12,
        "note": "achieved = algorithmic FP64 flops of the zero-skipping factored algorithm (DESIGN.md §5) / CUDA-event "
                "time of the kernel launches; the HBM roofline named by the contract does not bound this kernel "
                "(arithmetic intensity > 1e4 flop/B)",
    }

    # tensor-pipe view of the same kernel: every right-hand-side evaluation of a CTA (6 trajectories) is one
    # D[128 x 96] += A[128 x 128] B[128 x 96] int8 GEMM (zero padding of rows / K / the two pad columns included)
    evals = 6.0 * float(nst.item()) * args.steps / world  # per GPU in the timed region
    int8_ops = evals * (2.0 * 128 * 128 * 16)             # 16 accumulator columns per trajectory
    roofline["tensor"] = {
        "bound": "tensor", "achieved": int8_ops / (ms_total * 1e-3) / 1e12, "peak": 8192 * 2 * 148 * 1.965e9 / 1e12,
        "unit": "TOP/s (int8)", "frac": int8_ops / (ms_total * 1e-3) / (8192 * 2 * 148 * 1.965e9),
        "peak_source": "tcgen05.mma kind::i8 M=128: 8192 MAC/clk/SM measured by profiles/microbench/umma_probe.cu, x 148 SMs x 1.965 GHz",
        "note": "the tensor cores replace 54 % of the algorithmic FP64 flops; they are not the bound of the kernel",
    }

    # ---- cluster time at N runs (fused row-sharded DBSCAN on the all-gathered features) -----------------------
    cluster = None
    if not args.no_cluster:
        Xl = torch.cat([feat_d, par_d], dim=0).contiguous()  # [F, R] features + parameter column
        if dist is not None:
            parts = [torch.empty_like(Xl) for _ in range(world)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dist.all_gather(parts, Xl)
            torch.cuda.synchronize()
            t_gather = time.perf_counter() - t0
            X = torch.cat(parts, dim=1).contiguous()
        else:
            X, t_gather = Xl, 0.0
        X = torch.nan_to_num(X, nan=1e30)
        # MCBB.solve sorts the runs by parameter before any clustering (sort!, setup_mc_prob.jl:1124-1126):
        # stable sort on the parameter row, identical on every rank
        order = torch.sort(X[-1], stable=True).indices
        X = X[:, order].contiguous()
        n_all = X.shape[1]
        gl = np.array([nf, nf, 1], dtype=np.int32)
        gw = np.array([1.0, 0.5, 1.0])
        # eps: median distance to the 0.5 %-th nearest neighbour on a 4096-run subsample (KNN_dist_relative's
        # rel_K = 0.005, eval_clustering.jl:1526-1553), identical on every rank
        sub = torch.from_numpy(np.random.default_rng(7).choice(n_all, size=min(4096, n_all), replace=False)).to(dev)
        Xs = X[:, sub].contiguous()
        ms_ = Xs.shape[1]
        Ds = torch.empty((ms_, ms_), dtype=torch.float64, device=dev)
        m.check(m.lib().mcbb_distance_matrix_dev(C.c_void_p(Xs.data_ptr()), ms_, 3, gl.ctypes.data_as(ip),
                                                 gw.ctypes.data_as(dp), C.c_void_p(Ds.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        kth = max(2, int(round(0.005 * ms_)))
        eps = float(torch.sort(Ds, dim=0).values[kth].median().item())
        minpts = 20
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            a_, s_, c_ = m.dist.dbscan_features_sharded(X, gl, gw, eps, minpts, rank, world, dist)
        else:
            a_, s_, c_ = m.dist.dbscan_features_device(X, gl, gw, eps, minpts)
        barrier()
        t_cl = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_cl, op=dist.ReduceOp.MAX)
        cluster = {"n_runs": int(n_all), "time_s": float(t_cl.item()), "allgather_s": t_gather, "eps": eps,
                   "minpts": minpts, "n_clusters": int(s_.numel()), "n_noise": int((a_ == 0).sum().item()),
                   "n_features": int(X.shape[0]),
                   "pairs_per_s": 0.5 * n_all * (n_all - 1) / float(t_cl.item())}

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:  # reported at N=1 only (rank 0)
            v, dt, threads = cpu_baseline(m, A, psys, pfeat, opts, ts, args.cpu_sample)
            cpu = {"value": v, "unit": "trajectories/s", "cores": threads, "kind": "port",
                   "sample": "%d trajectories of the same C5 workload in %.1f s (oracle restatement of the reference's "
                             "CPU path; Julia itself is not installable here)" % (args.cpu_sample, dt)}
        line = {
            "metric": "trajectories/sec (integrate+featurize)", "value": value, "unit": "trajectories/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5 kuramoto_network N=100 ER p=0.2, tspan (0,1000), tail 0.9, N_t=400, Tsit5 "
                                   "abstol 1e-6 reltol 1e-3, measures [mean,std] cyclic_setback",
                       "runs_per_gpu_per_step": R, "runs_per_step": world * R, "l2": "flushed between timed steps",
                       "mean_step_attempts_per_run": float(nst.item()) / (world * R)},
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "bitwise_equal_to_resident": same},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "cluster": cluster, "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
