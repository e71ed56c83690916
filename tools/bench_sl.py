"""Times C4's system (Stuart-Landau ring N=100, tspan (0,200), tail 0.7, mean/std/KL of re and im) on a sample."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
m = load_package()
m.check(m.lib().mcbb_init(0))
N = 100
rng = np.random.default_rng(7)
def ring_adj(P):
    A = np.zeros((N, N))
    for d in range(1, P + 1):
        A += np.roll(np.eye(N), d, 0) + np.roll(np.eye(N), -d, 0)
    return A
pars = m.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.925, N, 1, 22, ring_adj(1), ring_adj(22))
ev = m.EvalOdeRun(eval_funcs=("mean_real", "std_real", "kl_real", "mean_imag", "std_imag", "kl_imag"))
rp = m.ODEProblem(m.stuart_landau_sathiyadevi, np.zeros(N, dtype=complex), (0.0, 200.0), pars)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ic = (rng.uniform(-1, 1, (n, N)) + 1j * rng.uniform(-1, 1, (n, N)))
it = iter(range(10**9))
prob = m.DEMCBBProblem(rp, lambda: ic[next(it) % n], n, pars, ("eps", lambda: rng.uniform(1.8, 2.1)), ev, 0.7)
m.solve(prob, flag_check_inf_nan=False)
t0 = time.perf_counter()
sol = m.solve(prob, flag_check_inf_nan=False)
dt = time.perf_counter() - t0
print({"n": n, "solve_s": dt, "traj_per_s": n / dt, "mean_steps": float(sol.nsteps[:, 0].mean()), "failed": int((sol.retcode != 0).sum())})
