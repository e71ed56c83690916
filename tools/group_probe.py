"""Group kernel (csrc/group.cu) against the lane kernel on the same ensembles: C1-type problems of several sizes and C3.
Prints one JSON object: case -> {group_ms, lane_ms, equal}."""
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
    dev = torch.device("cuda:0")
    out = {}
    for name, make in (("C1_2500", lambda: cases.c1_problem(m, n_ics=500)[0]),
                       ("C1_100k", lambda: cases.c1_problem(m, n_ics=20000)[0]),
                       ("C1_1M_mean_std", lambda: cases.c1_problem(m, n_ics=200000, eval_funcs=(m.mean, m.std))[0]),
                       ("C3_100k", lambda: cases.c3_problem(m, n_mc=100000, tspan=(0.0, 1000.0))[0])):
        prob = make()
        psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
        opts = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
        ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
        n = prob.N_mc
        ic_d = torch.from_numpy(np.ascontiguousarray(prob.ic.T)).to(dev)
        par_d = torch.from_numpy(np.ascontiguousarray(prob.par.T)).to(dev)
        ts_d = torch.from_numpy(ts).to(dev)
        nf = pfeat.n_filter(psys)
        p = lambda t: C.c_void_p(t.data_ptr())
        res = {}
        for variant, hook in (("group", -2 ** 63), ("lane", 1)):
            m.lib().mcbb_test_hook_set(b"no_group", hook)
            feat = torch.full((pfeat.n_meas * nf, n), float("nan"), dtype=torch.float64, device=dev)
            ret = torch.empty((n,), dtype=torch.int32, device=dev)
            ns = torch.empty((2, n), dtype=torch.int64, device=dev)

            def run():
                m.check(m.lib().mcbb_solve_features_dev(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n, p(ic_d), p(par_d),
                                                        p(ts_d), ts_d.numel(), p(feat), C.c_void_p(0), C.c_void_p(0), p(ret), p(ns),
                                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            run()
            torch.cuda.synchronize()
            ms = []
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            res[variant] = (min(ms), feat.clone(), ns.clone())
        m.lib().mcbb_test_hook_set(b"no_group", -2 ** 63)
        eq = bool(torch.equal(torch.nan_to_num(res["group"][1]), torch.nan_to_num(res["lane"][1])) and
                  torch.equal(res["group"][2], res["lane"][2]))
        out[name] = {"n_runs": n, "group_ms": round(res["group"][0], 3), "lane_ms": round(res["lane"][0], 3),
                     "group_traj_per_s": n / (res["group"][0] * 1e-3), "lane_traj_per_s": n / (res["lane"][0] * 1e-3),
                     "bitwise_equal": eq}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
