"""Tiny end-to-end invocations of every kernel family, for compute-sanitizer (one tool per gpurun call)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from tests import cases  # noqa: E402

m = load_package()
m.check(m.lib().mcbb_init(0))
# lane kernel: mean/std/KL + correlation + sum
prob, w = cases.c1_problem(m, n_ics=20, seed=1, N=5)
prob.eval_ode_func = m.EvalOdeRun(eval_funcs=(m.mean, m.std, m.empirical_1D_KL_divergence_hist),
                                  matrix_eval_funcs=[m.correlation_ecdf], global_eval_funcs=["sum"])
sol = m.solve(prob, flag_check_inf_nan=False)
D = m.distance_matrix(sol, prob, [1, 0.5, 0.5, 0.25, 0.1, 1])
m.dbscan(D, float(np.nanmedian(D.data)), 4)
m.dbscan(m.distance_matrix(sol, prob, [1, 0.5, 0.5, 0.25, 0.1, 1], materialize=False), float(np.nanmedian(D.data)), 4)
m.k_dist(np.nan_to_num(D.data), 4)
# logistic map
prob, w = cases.c2_problem(m, n_mc=200, tspan=(0.0, 300.0))
m.solve(prob)
# IMMA batched kernel (N=100 unit adjacency) and FP64 batched kernel (ring, weighted)
prob, w = cases.c5_problem(m, 40, tspan=(0.0, 30.0))
sol = m.solve(prob)
fm = m.distance_matrix(sol, prob, w, materialize=False)
m.dbscan(fm, 50.0, 3)
N = 40
pars = m.non_local_kuramoto_ring_parameters(N, 0.2, 0.3, lambda x: np.exp(-2 * abs(x)) / (2 * np.pi))
rp = m.ODEProblem(m.non_local_kuramoto_ring, np.zeros(N), (0.0, 10.0), pars)
rng = np.random.default_rng(0)
prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-3, 3), 30, pars, ("phase_delay", lambda: rng.uniform(0, 0.5)),
                       m.EvalOdeRun(eval_funcs=(m.mean, m.std)), 0.8)
m.solve(prob)
# global-state lane kernel (Stuart-Landau N=100)
N = 100
A1 = np.roll(np.eye(N), 1, 0) + np.roll(np.eye(N), -1, 0)
pars = m.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.9, N, 1, 1, A1, A1)
rp = m.ODEProblem(m.stuart_landau_sathiyadevi, np.zeros(N, dtype=complex), (0.0, 3.0), pars)
prob = m.DEMCBBProblem(rp, lambda: complex(rng.uniform(-1, 1), rng.uniform(-1, 1)), 8, pars,
                       ("eps", lambda: rng.uniform(1.8, 2.1)),
                       m.EvalOdeRun(eval_funcs=("mean_real", "std_real", "mean_imag", "std_imag")), 0.7)
m.solve(prob)
print("sanitize_small: all kernels ran")
