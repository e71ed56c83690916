"""Where do C1's 13 ms go?  Times the C1 solve (2500 runs, N = 6) resident on the device for a few measure sets / sizes."""
import ctypes as C
import json
import os
import sys
import time

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
    for name, funcs, n_ics in (("mean_std", (m.mean, m.std), 500), ("mean_std_kl", None, 500), ("mean_std_kl_x8", None, 4000)):
        prob, _ = cases.c1_problem(m, n_ics=n_ics, eval_funcs=funcs)
        psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
        opts = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
        ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
        n = prob.N_mc
        ic_d = torch.from_numpy(np.ascontiguousarray(prob.ic.T)).to(dev)
        par_d = torch.from_numpy(np.ascontiguousarray(prob.par.T)).to(dev)
        ts_d = torch.from_numpy(ts).to(dev)
        nf = pfeat.n_filter(psys)
        feat = torch.empty((pfeat.n_meas * nf, n), dtype=torch.float64, device=dev)
        ret = torch.empty((n,), dtype=torch.int32, device=dev)
        ns = torch.empty((2, n), dtype=torch.int64, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())

        def run():
            m.check(m.lib().mcbb_solve_features_dev(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n, p(ic_d), p(par_d),
                                                    p(ts_d), ts_d.numel(), p(feat), C.c_void_p(0), C.c_void_p(0), p(ret), p(ns),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        run()
        torch.cuda.synchronize()
        before = m.kernel_launch_count()
        ms, wall = [], []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(e0.elapsed_time(e1))
        out[name] = {"n": n, "ms_events": [round(x, 3) for x in ms], "ms_wall": [round(x, 3) for x in wall],
                     "launches": (m.kernel_launch_count() - before) / 3, "steps": float(ns[0].double().mean().item()),
                     "n_t": len(ts)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
