#!/usr/bin/env python
"""One-off confidence check (not part of pytest: the oracle needs ~10 s per 1000 runs at N = 100): a larger converged
C5 sweep through the tensor-core integrator against the oracle, every run compared at 1e-6."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from __graft_entry__ import load_package
    m = load_package()  # before the oracle wrapper: both must share one ctypes ABI module
    from oracle import oracle as orc
    from tests import cases
    m.check(m.lib().mcbb_init(0))
    n = int(os.environ.get("RUNS", 1536))
    prob, _ = cases.c5_problem(m, n, N=100, seed=77, tspan=(0.0, 60.0))
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-9, abstol=1e-9, maxiters=10**7)
    sol = m.solve(prob, **kw)
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    o = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan, **kw)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    idx = np.argsort(par0[:, 0], kind="stable")
    t0 = time.perf_counter()
    ref = orc.solve_features(psys, o, pfeat, ic0[idx], par0[idx], ts, n_threads=os.cpu_count())
    print("oracle %.1f s" % (time.perf_counter() - t0))
    assert np.array_equal(sol.retcode, ref["retcode"])
    stats = cases.compare_features(sol.feat, ref, orc, psys, o, pfeat, ic0[idx], par0[idx], ts, min_checked=0.9)
    print("parity sweep:", n, "runs", stats)


if __name__ == "__main__":
    main()
