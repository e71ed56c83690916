"""Generates tests/golden/oracle_golden.json from the CPU oracle (python -m tests.golden.make_golden).

The reference (Julia) cannot run in this image and its tests contain no expected values, so these vectors do NOT
come from the reference: they freeze the oracle restatement (which is itself checked against analytic answers,
scipy and sklearn in tests/test_oracle_cpu.py) so that later edits cannot drift silently."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def compute_all(orc):
    abi = orc.abi
    out = {}
    rng = np.random.default_rng(2024)
    # C1 shape, 16 runs
    N = 6
    w = rng.normal(0.5, 0.01, N)
    ps = abi.PackedSystem(abi.SYS_KURAMOTO_NETWORK, N, N, [0.5], [0], mat_a=np.ones((N, N)), vec_a=w)
    o = abi.solver_opts(abi.ALG_TSIT5, (0.0, 40.0))
    ts = orc.tsave_array(o, 400, 0.9)
    ic = rng.uniform(-np.pi, np.pi, (16, N))
    K = np.repeat([1.0, 1.5, 2.0, 3.0], 4)[:, None]
    pf = abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD, abi.MEAS_KL_HIST])
    r = orc.solve_features(ps, o, pf, ic, K, ts)
    out["c1_feat"] = r["feat"].tolist()
    out["c1_nsteps"] = r["nsteps"].tolist()
    # C2 shape, 32 runs
    pl = abi.PackedSystem(abi.SYS_LOGISTIC, 1, 1, [2.5], [0])
    ol = abi.solver_opts(abi.ALG_FUNCTIONMAP, (0.0, 1000.0))
    tl = orc.tsave_array(ol, 400, 0.8)
    r = orc.solve_features(pl, ol, pf, rng.uniform(0.1, 0.99, (32, 1)), rng.uniform(2.5, 4.0, (32, 1)), tl)
    out["c2_feat"] = r["feat"].tolist()
    # DBSCAN on a tie-rich lattice
    g = np.array([(i, j) for i in range(8) for j in range(8)], dtype=float)[rng.permutation(64)][:50]
    D = orc.distance_matrix(g, np.array([2], dtype=np.int32), np.array([1.0]))
    for eps, mp in [(1.0, 3), (2.0, 4), (1.5, 1)]:
        db = orc.dbscan(D, eps, mp)
        out["db_%s_%d" % (eps, mp)] = db["assignments"].tolist()
        out["db_seeds_%s_%d" % (eps, mp)] = db["seeds"].tolist()
    out["D_lattice_row0"] = D[0].tolist()
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    json.dump(compute_all(orc), open(os.path.join(HERE, "oracle_golden.json"), "w"))
    print("wrote oracle_golden.json")
