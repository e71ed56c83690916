"""Times C4's integrate+featurize stage (Stuart-Landau N=100, Watts-Strogatz A_1/A_2, tspan (0,200), tail 0.7) resident on
the device, for three measure sets: mean/std, + KL-hist, + correlation_ecdf (the notebook's full set).

    python tools/bench_sl.py [n_runs] [hook=value ...]      prints one JSON object (hooks: mcbb_test_hook_set)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
from tests import cases  # noqa: E402


def main():
    m = load_package()
    m.check(m.lib().mcbb_init(0))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    hooks = dict(kv.split("=") for kv in sys.argv[2:])
    for k, v in hooks.items():
        m.check(m.lib().mcbb_test_hook_set(k.encode(), int(v)))
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(7)
    ic = np.empty((200, n))
    ic[0::2] = rng.uniform(-1, 1, (100, n))
    ic[1::2] = rng.uniform(-1, 1, (100, n))
    par = rng.uniform(1.8, 2.1, (1, n))
    ic_d, par_d = torch.from_numpy(ic).to(dev), torch.from_numpy(par).to(dev)
    out = {"n_runs": n, "hooks": hooks}
    sets = {"mean_std": (("mean_real", "std_real", "mean_imag", "std_imag"), False),
            "mean_std_kl": (("mean_real", "std_real", "kl_real", "mean_imag", "std_imag", "kl_imag"), False),
            "mean_std_kl_corr_ecdf": (("mean_real", "std_real", "kl_real", "mean_imag", "std_imag", "kl_imag"), True)}
    for name, (funcs, corr) in sets.items():
        prob, _ = cases.c4_problem(m, 4, corr=corr)
        prob.eval_ode_func = m.EvalOdeRun(eval_funcs=funcs, matrix_eval_funcs=[m.correlation_ecdf] if corr else [])
        psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
        opts = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
        ts_d = torch.from_numpy(m.tsave_array(prob.p, 400, prob.rel_transient_time)).to(dev)
        nf = pfeat.n_filter(psys)
        feat = torch.empty((pfeat.n_meas * nf, n), dtype=torch.float64, device=dev)
        mat = torch.empty((30, n), dtype=torch.float64, device=dev)
        ret = torch.empty((n,), dtype=torch.int32, device=dev)
        ns = torch.empty((2, n), dtype=torch.int64, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())

        def run():
            m.check(m.lib().mcbb_solve_features_dev(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n, p(ic_d), p(par_d),
                                                    p(ts_d), ts_d.numel(), p(feat), p(mat) if corr else C.c_void_p(0),
                                                    C.c_void_p(0), p(ret), p(ns),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        run()
        torch.cuda.synchronize()
        before = m.kernel_launch_count()
        ms = []
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        out[name] = {"traj_per_s": n / (min(ms) * 1e-3), "ms": [round(x, 2) for x in ms],
                     "launches_per_solve": (m.kernel_launch_count() - before) / 2,
                     "mean_step_attempts": float(ns[0].double().mean().item()), "failed": int((ret != 0).sum().item())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
