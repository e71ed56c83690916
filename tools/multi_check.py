"""Drives the in-library multi-device entry points (mcbb_multi_init / mcbb_solve_features_multi / mcbb_dbscan_features_multi
/ mcbb_solve_cluster_multi) from ONE process with ctypes and numpy only — no torch, no torch.distributed — the way a Julia
host would through ccall, and checks them against the single-device entry points.

    python tools/multi_check.py [n_devices]        (default: as many as mcbb_multi_init accepts, at most 8)
Prints one JSON object; exits non-zero on any mismatch."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n_devices=None, n_solve=3000, n_cluster=40000, verbose=True, forbid_torch=False):
    from __graft_entry__ import load_package
    from tests import cases
    m = load_package()
    L = m.lib()
    if forbid_torch:
        assert "torch" not in sys.modules, "this check must not need torch"
    m.check(L.mcbb_init(0))
    G = None
    for g in ([n_devices] if n_devices else [8, 4, 2, 1]):
        if L.mcbb_multi_init(g, None) == 0:
            G = g
            break
    if G is None:
        raise SystemExit("mcbb_multi_init failed: " + m.last_error())
    nd, ver = C.c_int32(0), C.c_int32(0)
    m.check(L.mcbb_multi_info(C.byref(nd), C.byref(ver)))
    out = {"n_devices": int(nd.value), "nccl_version": int(ver.value)}
    dp, ip, lp = m.abi.c_double_p, m.abi.c_int32_p, m.abi.c_int64_p
    try:
        # ---- 1. integrate + featurize: every device a block of the runs, results bitwise equal to one device ------
        prob, weights = cases.c5_problem(m, n_solve, N=48, seed=61, tspan=(0.0, 60.0))
        psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
        opts = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
        ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
        ic = np.asfortranarray(prob.ic)
        par = np.asfortranarray(prob.par)
        nf = pfeat.n_filter(psys)

        def solve(fn):
            feat = np.full((n_solve, nf, pfeat.n_meas), np.nan, order="F")
            ret = np.zeros(n_solve, dtype=np.int32)
            ns = np.zeros((n_solve, 2), dtype=np.int64, order="F")
            t0 = time.perf_counter()
            m.check(fn(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n_solve, ic.ctypes.data_as(dp),
                       par.ctypes.data_as(dp), ts.ctypes.data_as(dp), len(ts), feat.ctypes.data_as(dp), dp(), dp(),
                       ret.ctypes.data_as(ip), ns.ctypes.data_as(lp)))
            return feat, ret, ns, time.perf_counter() - t0
        f1, r1, n1, _ = solve(L.mcbb_solve_features)
        fG, rG, nG, tG = solve(L.mcbb_solve_features_multi)
        out["solve_equal"] = bool(np.array_equal(f1, fG, equal_nan=True) and np.array_equal(r1, rG) and np.array_equal(n1, nG))
        out["solve_multi_s"] = tG
        # ---- 2. DBSCAN over all devices == single-device fused DBSCAN ------------------------------------------
        rng = np.random.default_rng(4)
        n, F = n_cluster, 12
        centers = rng.normal(0, 4, size=(6, F))
        X = centers[rng.integers(0, 6, n)] + rng.normal(0, 0.4, size=(n, F))
        X = np.asfortranarray(X[np.argsort(X[:, -1], kind="stable")])  # sorted along one feature, like runs by parameter
        gl = np.array([5, 6, 1], dtype=np.int32)
        gw = np.array([1.0, 0.5, 1.0])

        def cluster(fn):
            a, s, c = (np.zeros(n, dtype=np.int64) for _ in range(3))
            k = np.zeros(1, dtype=np.int64)
            t0 = time.perf_counter()
            m.check(fn(X.ctypes.data_as(dp), n, 3, gl.ctypes.data_as(ip), gw.ctypes.data_as(dp), 2.2, 10, a.ctypes.data_as(lp),
                       s.ctypes.data_as(lp), c.ctypes.data_as(lp), k.ctypes.data_as(lp)))
            return a, s[:k[0]].copy(), c[:k[0]].copy(), time.perf_counter() - t0
        a1, s1, c1, _ = cluster(L.mcbb_dbscan_features)
        aG, sG, cG, tC = cluster(L.mcbb_dbscan_features_multi)
        out["dbscan_equal"] = bool(np.array_equal(a1, aG) and np.array_equal(s1, sG) and np.array_equal(c1, cG))
        out["dbscan_clusters"] = int(len(s1))
        out["dbscan_multi_s"] = tC
        # ---- 3. solve -> sort -> dbscan in one call == the same steps through the single-device entries ----------
        w = np.array([1.0, 0.5, 1.0])
        feat = np.full((n_solve, nf, pfeat.n_meas), np.nan, order="F")
        ret = np.zeros(n_solve, dtype=np.int32)
        ns = np.zeros((n_solve, 2), dtype=np.int64, order="F")
        order = np.zeros(n_solve, dtype=np.int64)
        a, s, c = (np.zeros(n_solve, dtype=np.int64) for _ in range(3))
        k = np.zeros(1, dtype=np.int64)
        eps, minpts = 30.0, 5
        t0 = time.perf_counter()
        m.check(L.mcbb_solve_cluster_multi(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n_solve,
                                           ic.ctypes.data_as(dp), par.ctypes.data_as(dp), ts.ctypes.data_as(dp), len(ts),
                                           w.ctypes.data_as(dp), eps, minpts, feat.ctypes.data_as(dp), ret.ctypes.data_as(ip),
                                           ns.ctypes.data_as(lp), order.ctypes.data_as(lp), a.ctypes.data_as(lp),
                                           s.ctypes.data_as(lp), c.ctypes.data_as(lp), k.ctypes.data_as(lp)))
        out["solve_cluster_multi_s"] = time.perf_counter() - t0
        want_order = np.argsort(par[:, 0], kind="stable")
        Xs = np.asfortranarray(np.concatenate([f1[:, :, 0], f1[:, :, 1], par], axis=1)[want_order])
        Xs = np.nan_to_num(Xs, nan=1e30)
        gl2 = np.array([nf, nf, 1], dtype=np.int32)
        a2, s2, c2 = (np.zeros(n_solve, dtype=np.int64) for _ in range(3))
        k2 = np.zeros(1, dtype=np.int64)
        m.check(L.mcbb_dbscan_features(Xs.ctypes.data_as(dp), n_solve, 3, gl2.ctypes.data_as(ip), w.ctypes.data_as(dp), eps,
                                       minpts, a2.ctypes.data_as(lp), s2.ctypes.data_as(lp), c2.ctypes.data_as(lp),
                                       k2.ctypes.data_as(lp)))
        out["pipeline_equal"] = bool(np.array_equal(order, want_order) and np.array_equal(feat, f1, equal_nan=True)
                                     and np.array_equal(a, a2) and k[0] == k2[0]
                                     and np.array_equal(s[:k[0]], s2[:k2[0]]) and np.array_equal(c[:k[0]], c2[:k2[0]]))
        out["pipeline_clusters"] = int(k[0])
    finally:
        L.mcbb_multi_finalize()
    if verbose:
        print(json.dumps(out))
    return out


if __name__ == "__main__":
    res = run(int(sys.argv[1]) if len(sys.argv) > 1 else None, forbid_torch=True)
    sys.exit(0 if res["solve_equal"] and res["dbscan_equal"] and res["pipeline_equal"] else 1)
