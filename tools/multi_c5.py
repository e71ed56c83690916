"""north_star's end-to-end target through the in-library multi-device entry point: 10^6 C5 runs (Kuramoto network N = 100)
integrated, featurized, sorted and DBSCAN-clustered by ONE call of mcbb_solve_cluster_multi from host buffers — one host
process, ctypes + numpy only, NCCL inside the library.

    python tools/multi_c5.py [n_devices] [runs_per_device] [eps]     prints one JSON object"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from __graft_entry__ import load_package
    m = load_package()
    L = m.lib()
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
    eps = float(sys.argv[3]) if len(sys.argv) > 3 else 9454.0  # median(KNN_dist_relative) of this workload (bench.py)
    n = G * R
    m.check(L.mcbb_init(0))
    m.check(L.mcbb_multi_init(G, None))
    A, psys, pfeat, opts = bench.make_system(m.abi)
    nt = C.c_int32(0)
    m.check(L.mcbb_tsave_array(C.byref(opts), 400, 0.9, m.abi.c_double_p(), C.byref(nt)))
    ts = np.zeros(nt.value)
    m.check(L.mcbb_tsave_array(C.byref(opts), 400, 0.9, ts.ctypes.data_as(m.abi.c_double_p), C.byref(nt)))
    t0 = time.perf_counter()
    ic, par = bench.make_runs(n, 0)
    t_gen = time.perf_counter() - t0
    dp, ip, lp = m.abi.c_double_p, m.abi.c_int32_p, m.abi.c_int64_p
    nf = 100
    feat = np.empty((n, nf, 2), order="F")
    ret = np.zeros(n, dtype=np.int32)
    ns = np.zeros((n, 2), dtype=np.int64, order="F")
    order = np.zeros(n, dtype=np.int64)
    a, s, c = (np.zeros(n, dtype=np.int64) for _ in range(3))
    k = np.zeros(1, dtype=np.int64)
    w = np.array([1.0, 0.5, 1.0])
    out = {"n_devices": G, "n_runs": n, "eps": eps, "minpts": 20, "input_generation_s": t_gen}
    for rep in ("first_call_s", "second_call_s"):
        t0 = time.perf_counter()
        m.check(L.mcbb_solve_cluster_multi(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n, ic.ctypes.data_as(dp),
                                           par.ctypes.data_as(dp), ts.ctypes.data_as(dp), len(ts), w.ctypes.data_as(dp), eps, 20,
                                           feat.ctypes.data_as(dp), ret.ctypes.data_as(ip), ns.ctypes.data_as(lp),
                                           order.ctypes.data_as(lp), a.ctypes.data_as(lp), s.ctypes.data_as(lp),
                                           c.ctypes.data_as(lp), k.ctypes.data_as(lp)))
        out[rep] = time.perf_counter() - t0
    t0 = time.perf_counter()
    f2 = np.empty((n, nf, 2), order="F")
    m.check(L.mcbb_solve_features_multi(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n, ic.ctypes.data_as(dp),
                                        par.ctypes.data_as(dp), ts.ctypes.data_as(dp), len(ts), f2.ctypes.data_as(dp), dp(), dp(),
                                        ret.ctypes.data_as(ip), ns.ctypes.data_as(lp)))
    out["solve_features_multi_s"] = time.perf_counter() - t0
    out["traj_per_s_solve_multi_host_buffers"] = n / out["solve_features_multi_s"]
    out["features_equal_between_entries"] = bool(np.array_equal(feat, f2, equal_nan=True))
    out["n_clusters"] = int(k[0])
    out["n_noise"] = int((a == 0).sum())
    out["failed_runs"] = int((ret != 0).sum())
    out["order_is_stable_sortperm"] = bool(np.array_equal(order, np.argsort(par[:, 0], kind="stable")))
    out["mean_step_attempts"] = float(ns[:, 0].mean())
    m.check(L.mcbb_multi_finalize())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
