"""Generates tests/golden/oracle_golden.json from the CPU oracle (python -m tests.golden.make_golden).

The reference (Julia) cannot run in this image and its tests contain no expected values, so these vectors do NOT
come from the reference: they freeze the oracle restatement (which is itself checked against analytic answers,
scipy and sklearn in tests/test_oracle_cpu.py) so that later edits cannot drift silently.  `build_inputs` rebuilds
the seeded inputs alone, so the CUDA path (tests/test_gpu_parity.py) and the host build of the lane source
(tests/test_lane_source_on_host.py) can be compared with the committed vectors without running the oracle."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build_inputs(abi):
    """The seeded inputs of every golden case (one generator, read in a fixed order)."""
    I = {}
    rng = np.random.default_rng(2024)
    # C1 shape, 16 runs
    N = 6
    w = rng.normal(0.5, 0.01, N)
    I["c1_sys"] = abi.PackedSystem(abi.SYS_KURAMOTO_NETWORK, N, N, [0.5], [0], mat_a=np.ones((N, N)), vec_a=w)
    I["c1_opts"] = abi.solver_opts(abi.ALG_TSIT5, (0.0, 40.0))
    I["c1_ic"] = rng.uniform(-np.pi, np.pi, (16, N))
    I["c1_par"] = np.repeat([1.0, 1.5, 2.0, 3.0], 4)[:, None]
    I["feat_default"] = abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD, abi.MEAS_KL_HIST])
    # C2 shape, 32 runs
    I["c2_sys"] = abi.PackedSystem(abi.SYS_LOGISTIC, 1, 1, [2.5], [0])
    I["c2_opts"] = abi.solver_opts(abi.ALG_FUNCTIONMAP, (0.0, 1000.0))
    I["c2_ic"] = rng.uniform(0.1, 0.99, (32, 1))
    I["c2_par"] = rng.uniform(2.5, 4.0, (32, 1))
    # DBSCAN on a tie-rich lattice
    I["lattice"] = np.array([(i, j) for i in range(8) for j in range(8)], dtype=float)[rng.permutation(64)][:50]
    I["lattice_cases"] = [(1.0, 3), (2.0, 4), (1.5, 1)]
    # power-grid frequencies (C3 shape, 8 runs)
    Np = 6
    E = np.zeros((Np, Np + 1))
    for e, (a, b) in enumerate([(i, (i + 1) % Np) for i in range(Np)] + [(0, 3)]):
        E[a, e], E[b, e] = -1.0, 1.0
    I["c3_sys"] = abi.PackedSystem(abi.SYS_SECOND_ORDER_KURAMOTO, Np, 2 * Np, [0.1, 1.0], [1], mat_a=E,
                                   vec_a=np.array([1.0, -1.0] * (Np // 2)), n_edges=Np + 1)
    I["c3_opts"] = abi.solver_opts(abi.ALG_TSIT5, (0.0, 60.0), reltol=1e-9, abstol=1e-9)
    I["c3_ic"] = rng.uniform(-np.pi, np.pi, (8, 2 * Np))
    I["c3_par"] = rng.uniform(0.5, 4.0, (8, 1))
    filt = list(range(Np + 1, 2 * Np + 1))
    I["c3_feat"] = {name: abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD], state_filter=filt, matrix_meas=kind,
                                         corr_nbins=30)
                    for name, kind in (("hist", abi.MAT_CORR_HIST), ("ecdf", abi.MAT_CORR_ECDF))}
    # complex states: a synthetic complex series for featurize_saved
    Nc = 5
    ring = np.roll(np.eye(Nc), 1, 0) + np.roll(np.eye(Nc), -1, 0)
    I["sl_sys"] = abi.PackedSystem(abi.SYS_STUART_LANDAU, Nc, 2 * Nc, [1.9, 1.0, 2.0, 1.0, 1.0], [0], mat_a=ring,
                                   mat_b=ring)
    z = (rng.normal(size=(3, Nc, Nc)) + 1j * rng.normal(size=(3, Nc, Nc))) @ (
        rng.normal(size=(3, Nc, 60)) + 1j * rng.normal(size=(3, Nc, 60)))
    I["sl_saved"] = np.stack([z.real, z.imag], axis=2).reshape(3, 2 * Nc, 60)
    I["sl_feat"] = abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD], meas_comp=[0, 1], matrix_meas=abi.MAT_CORR_HIST,
                                  corr_nbins=30)
    return I


def compute_all(orc):
    abi = orc.abi
    I = build_inputs(abi)
    out = {}
    ts = orc.tsave_array(I["c1_opts"], 400, 0.9)
    r = orc.solve_features(I["c1_sys"], I["c1_opts"], I["feat_default"], I["c1_ic"], I["c1_par"], ts)
    out["c1_feat"] = r["feat"].tolist()
    out["c1_nsteps"] = r["nsteps"].tolist()
    tl = orc.tsave_array(I["c2_opts"], 400, 0.8)
    r = orc.solve_features(I["c2_sys"], I["c2_opts"], I["feat_default"], I["c2_ic"], I["c2_par"], tl)
    out["c2_feat"] = r["feat"].tolist()
    D = orc.distance_matrix(I["lattice"], np.array([2], dtype=np.int32), np.array([1.0]))
    for eps, mp in I["lattice_cases"]:
        db = orc.dbscan(D, eps, mp)
        out["db_%s_%d" % (eps, mp)] = db["assignments"].tolist()
        out["db_seeds_%s_%d" % (eps, mp)] = db["seeds"].tolist()
    out["D_lattice_row0"] = D[0].tolist()
    # --- added with the matrix measures / sample export
    tg = orc.tsave_array(I["c3_opts"], 400, 0.5)
    for name in ("hist", "ecdf"):  # correlation_hist / correlation_ecdf of the grid frequencies over the transient
        out["c3_corr_" + name] = orc.solve_features(I["c3_sys"], I["c3_opts"], I["c3_feat"][name], I["c3_ic"],
                                                    I["c3_par"], tg)["matrix"].tolist()
    # one saved trajectory of the same system (what get_trajectory returns), every 40th sample
    sv = orc.trajectory(I["c3_sys"], I["c3_opts"], I["c3_ic"][0], I["c3_par"][0], tg)[0]
    out["c3_trajectory_run0"] = sv[:, ::40].tolist()
    # |cor| histogram of the complex series
    out["complex_corr_hist"] = orc.featurize_saved(I["sl_sys"], I["sl_feat"], I["sl_saved"])["matrix"].tolist()
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    json.dump(compute_all(orc), open(os.path.join(HERE, "oracle_golden.json"), "w"))
    print("wrote oracle_golden.json")
