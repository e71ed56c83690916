#!/usr/bin/env python
"""Micro-benchmark of the clustering back end on synthetic featurized runs (device resident, CUDA events):
materialising distance_matrix (HBM-write roofline) and the fused DBSCAN passes (FP64-add roofline).
Usage: python tools/bench_cluster.py [--n 20000] [--f 201] [--n-fused 125000]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--f", type=int, default=19)
    ap.add_argument("--n-fused", type=int, default=100000)
    ap.add_argument("--f-fused", type=int, default=201)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from __graft_entry__ import load_package
    m = load_package()
    m.check(m.lib().mcbb_init(0))
    dev = torch.device("cuda", 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ip, dp = m.abi.c_int32_p, m.abi.c_double_p
    out = {}
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
    fp = C.c_double(0)
    m.check(m.lib().mcbb_measure_fp64_fma_peak(C.byref(fp), None))

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best * 1e-3

    # ---- materialising distance matrix (C1-like feature count): bound by the 8 n^2 byte write ----------------
    g = torch.Generator(device=dev).manual_seed(1)
    n, F = args.n, args.f
    X = torch.randn((F, n), dtype=torch.float64, device=dev, generator=g)
    D = torch.empty((n, n), dtype=torch.float64, device=dev)
    gl = np.array([F - 1, 1], dtype=np.int32)
    gw = np.array([1.0, 1.0])
    t = timed(lambda: m.check(m.lib().mcbb_distance_matrix_dev(C.c_void_p(X.data_ptr()), n, 2, gl.ctypes.data_as(ip),
                                                               gw.ctypes.data_as(dp), C.c_void_p(D.data_ptr()), st)))
    bytes_alg = 8.0 * n * n + 8.0 * n * F
    out["distance_matrix"] = {"n": n, "F": F, "time_s": t, "achieved_gbs": bytes_alg / t / 1e9,
                              "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_frac": bytes_alg / t / 1e9 / peaks["hbm_gbs"],
                              "fp64_add_tflops": n * (n - 1) / 2 * 2 * F / t / 1e12,
                              "fp64_add_frac_of_fma_peak": n * (n - 1) / 2 * 2 * F / t / 1e12 / (fp.value / 2)}
    # ---- fused DBSCAN on clustered synthetic features -------------------------------------------------------------
    n, F = args.n_fused, args.f_fused
    centers = torch.randn((F, 8), dtype=torch.float64, device=dev, generator=g) * 3
    lab = torch.randint(0, 8, (n,), device=dev, generator=g)
    X = (centers[:, lab] + 0.3 * torch.randn((F, n), dtype=torch.float64, device=dev, generator=g)).contiguous()
    gl = np.array([F // 2, F - F // 2 - 1, 1], dtype=np.int32)
    gw = np.array([1.0, 0.5, 1.0])
    eps = 0.42 * F  # E|a-b| for sigma 0.3 is 0.34 per feature (weights 1 / 0.5): most same-blob pairs fall inside
    a_ = torch.empty((n,), dtype=torch.int64, device=dev)
    s_ = torch.empty((n,), dtype=torch.int64, device=dev)
    c_ = torch.empty((n,), dtype=torch.int64, device=dev)
    k = np.zeros(1, dtype=np.int64)
    t = timed(lambda: m.check(m.lib().mcbb_dbscan_features_dev(C.c_void_p(X.data_ptr()), n, 3, gl.ctypes.data_as(ip),
                                                               gw.ctypes.data_as(dp), float(eps), 20, C.c_void_p(a_.data_ptr()),
                                                               C.c_void_p(s_.data_ptr()), C.c_void_p(c_.data_ptr()),
                                                               k.ctypes.data_as(m.abi.c_int64_p), st)))
    out["dbscan_fused"] = {"n": n, "F": F, "time_s": t, "pairs_per_s": n * (n - 1) / 2 / t, "n_clusters": int(k[0]),
                           "n_noise": int((a_ == 0).sum().item()), "eps": eps}
    out["fp64_fma_peak_tflops"] = fp.value
    print(json.dumps(out))


if __name__ == "__main__":
    main()
