#!/usr/bin/env python
"""C5 workload (Kuramoto N=100) with the reference's DEFAULT measures (mean, std, KL-hist): trajectories/s of the
tensor-core integrator with the sample scratch + binning, next to the (mean, std) headline configuration."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    from __graft_entry__ import load_package
    m = load_package()
    m.check(m.lib().mcbb_init(0))
    dev = torch.device("cuda:0")
    R = int(os.environ.get("RUNS", 59200))
    A, ic, par, psys, pfeat, opts, ts = bench.make_inputs(m, R, 0)
    ic_d = torch.from_numpy(np.ascontiguousarray(ic.T)).to(dev)
    par_d = torch.from_numpy(np.ascontiguousarray(par.T)).to(dev)
    ts_d = torch.from_numpy(ts).to(dev)
    out = {}
    for name, funcs in (("mean_std", (m.mean, m.std)), ("mean_std_kl", (m.mean, m.std, m.empirical_1D_KL_divergence_hist))):
        pf = m.EvalOdeRun(eval_funcs=funcs, cyclic_setback=True).pack()
        for _ in range(2):
            feat, _, ret, nst = m.dist.solve_features_device(psys, opts, pf, ic_d, par_d, ts_d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            feat, _, ret, nst = m.dist.solve_features_device(psys, opts, pf, ic_d, par_d, ts_d)
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"traj_per_s": 3 * R / (e0.elapsed_time(e1) * 1e-3), "finite_frac": float(torch.isfinite(feat).float().mean())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
