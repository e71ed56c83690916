#!/usr/bin/env python
"""A/B harness for kernel experiments: runs the C5 headline workload (resident inputs) under several test-hook
variants (mcbb_test_hook_set keys) IN ONE PROCESS — one gpurun call instead of one per variant — times each with CUDA events and checks that every
variant returns bitwise the same features and step counts as the first one.

    python tools/ab_variants.py "" "umma_jb=4,umma_per_sm=4" "no_umma=1"
    RUNS=125000 REPS=3 python tools/ab_variants.py ...

The switches are read by the library at launch time (tune_get in csrc/), so setting them between launches is enough.  Prints one JSON object: variant -> {traj_per_s, ms, launches, equal_to_first}."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def parse(spec):
    env = {}
    for item in filter(None, (s.strip() for s in spec.split(","))):
        k, _, v = item.partition("=")
        env[k] = int(v) if v != "" else 1
    return env


def main():
    from __graft_entry__ import load_package
    m = load_package()
    m.check(m.lib().mcbb_init(0))
    dev = torch.device("cuda:0")
    R = int(os.environ.get("RUNS", 125000))
    reps = int(os.environ.get("REPS", 3))
    variants = sys.argv[1:] or [""]
    A, ic, par, psys, pfeat, opts, ts = bench.make_inputs(m, R, 0)
    ic_d = torch.from_numpy(np.ascontiguousarray(ic.T)).to(dev)
    par_d = torch.from_numpy(np.ascontiguousarray(par.T)).to(dev)
    ts_d = torch.from_numpy(ts).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2
    touched = set()
    out, first = {}, None
    for spec in variants:
        env = parse(spec)
        for k in touched:
            m.lib().mcbb_test_hook_set(k.encode(), -2**63)  # INT64_MIN erases the key
        for k, v in env.items():
            m.lib().mcbb_test_hook_set(k.encode(), v)
        touched |= set(env)
        for _ in range(2):  # warm-up: module load, scratch allocation
            feat, _, ret, nst = m.dist.solve_features_device(psys, opts, pfeat, ic_d, par_d, ts_d)
        torch.cuda.synchronize()
        before = m.kernel_launch_count()
        ms = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            feat, _, ret, nst = m.dist.solve_features_device(psys, opts, pfeat, ic_d, par_d, ts_d)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        launches = (m.kernel_launch_count() - before) / reps
        cur = (feat.clone(), nst.clone())
        if first is None:
            first = cur
        same = bool(torch.equal(cur[0].view(torch.int64), first[0].view(torch.int64)) and torch.equal(cur[1], first[1]))
        out[spec or "default"] = {"traj_per_s": R / (min(ms) * 1e-3), "ms": [round(x, 3) for x in ms],
                                  "launches_per_solve": launches, "equal_to_first": same,
                                  "failed": int((ret != 0).sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
