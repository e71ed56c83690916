#!/usr/bin/env python
"""Times the non-headline BASELINE configs at full size through the host API (solve -> distance_matrix/dbscan):
C1 Kuramoto N=6 (2500 runs), C2 logistic (10^4 runs), C3 second-order Kuramoto N=10 (10^5 runs, T=1000)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from __graft_entry__ import load_package
    from tests import cases
    m = load_package()
    m.check(m.lib().mcbb_init(0))
    out = {}

    def run(name, prob, weights, eps_k, minpts, materialize=True, **kw):
        m.solve(prob, **kw)  # warm-up (allocations, module load)
        t0 = time.perf_counter()
        sol = m.solve(prob, flag_check_inf_nan=False, **kw)
        t_solve = time.perf_counter() - t0
        r = {"n_mc": prob.N_mc, "solve_s": t_solve, "traj_per_s": prob.N_mc / t_solve,
             "mean_steps": float(sol.nsteps[:, 0].mean()), "failed": int((sol.retcode != 0).sum())}
        t0 = time.perf_counter()
        if materialize:
            D = m.distance_matrix(sol, prob, weights)
            r["distance_s"] = time.perf_counter() - t0
            eps = float(np.median(np.sort(D.data[:, ::max(1, prob.N_mc // 500)], axis=0)[eps_k]))
            t0 = time.perf_counter()
            db = m.dbscan(D, eps, minpts)
        else:
            fm = m.distance_matrix(sol, prob, weights, materialize=False)
            sub = np.random.default_rng(0).choice(prob.N_mc, 2000, replace=False)
            Xs = fm.X[sub]
            Ds = np.abs(Xs[:, None, :] - Xs[None, :, :])
            w = np.repeat(fm.group_weight, fm.group_len)
            eps = float(np.median(np.sort((Ds * w).sum(-1), axis=0)[eps_k]))
            t0 = time.perf_counter()
            db = m.dbscan(fm, eps, minpts)
        r["dbscan_s"] = time.perf_counter() - t0
        r["n_clusters"] = int(len(db.seeds))
        r["n_noise"] = int(m.cluster_n_noise(db))
        out[name] = r

    prob, w = cases.c1_problem(m, n_ics=500)
    run("C1_kuramoto_N6", prob, w, 12, 4)
    prob, w = cases.c2_problem(m, n_mc=10000)
    run("C2_logistic", prob, w, 50, 4)
    prob, w = cases.c3_problem(m, n_mc=100000, tspan=(0.0, 1000.0), lam_max=10.0)
    run("C3_second_order_kuramoto_N10", prob, w, 10, 500, materialize=False)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
