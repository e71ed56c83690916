"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on identical
seeded inputs.  Bit-exact for integer work (DBSCAN labels / seeds / counts), 1e-6 for FP64 features
(north_star), 1e-12 for distances."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _oracle_ref(m, orc, prob, ic0, par0, alg=None, n_t=400, **kw):
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    a = alg.alg if alg is not None else (m.abi.ALG_FUNCTIONMAP if prob.p.discrete else m.abi.ALG_TSIT5)
    o = m.abi.solver_opts(a, prob.p.tspan, **kw)
    ts = m.tsave_array(prob.p, n_t, prob.rel_transient_time)
    idx = np.argsort(par0[:, 0], kind="stable")
    ic, par = ic0[idx], par0[idx]
    return orc.solve_features(psys, o, pfeat, ic, par, ts), (psys, o, pfeat, ic, par, ts)


def test_loaded_library_is_the_cuda_one(gpu):
    sm, maj, mnr, mem = (C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64())
    gpu.check(gpu.lib().mcbb_device_info(C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)))
    assert maj.value == 10, "expected a Blackwell (sm_100) device, got cc %d.%d" % (maj.value, mnr.value)
    assert sm.value >= 100


def test_c1_kuramoto_solve_features(gpu, orc):
    """C1: Kuramoto N=6, 500 ICs x 5 K, Tsit5, default eval_ode_run (mean, std, KL)."""
    m = gpu
    prob, _ = cases.c1_problem(m, n_ics=500)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    before = m.kernel_launch_count()
    sol = m.solve(prob)
    assert m.kernel_launch_count() > before, "no CUDA kernel was launched"
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0)
    assert np.array_equal(prob.par[:, 0], np.sort(par0[:, 0]))  # sort!(sol, prob) permuted the problem
    assert np.array_equal(sol.retcode, ref["retcode"])
    assert np.array_equal(sol.nsteps, ref["nsteps"]), "step sequences differ from the oracle"
    stats = cases.compare_features(sol.feat, ref, orc, *ctx)
    assert stats["checked"] >= 0.95 * stats["n"]


def test_c1_variants_cyclic_filter_global(gpu, orc):
    m = gpu
    prob, _ = cases.c1_problem(m, n_ics=100, seed=7, eval_funcs=(m.mean, m.std), cyclic=True)
    prob.eval_ode_func = m.EvalOdeRun(state_filter=[2, 3, 4], eval_funcs=(m.mean, m.std),
                                      global_eval_funcs=["sum"], cyclic_setback=True)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    sol = m.solve(prob)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0)
    assert sol.feat.shape == (500, 3, 2)
    cases.compare_features(sol.feat, ref, orc, *ctx)
    assert np.allclose(sol.glob, ref["glob"], rtol=1e-6)


def test_c1_rk4_fixed_step(gpu, orc):
    m = gpu
    prob, _ = cases.c1_problem(m, n_ics=40, seed=11)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    sol = m.solve(prob, m.RK4(), dt=0.05)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, alg=m.RK4(), dt=0.05)
    assert np.array_equal(sol.nsteps, ref["nsteps"])
    cases.compare_features(sol.feat, ref, orc, *ctx)


def test_c2_logistic_map(gpu, orc):
    """C2: logistic map (DiscreteProblem / FunctionMap), 10^4 runs."""
    m = gpu
    prob, _ = cases.c2_problem(m, n_mc=10000)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    sol = m.solve(prob)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0)
    assert np.array_equal(sol.nsteps[:, 0], ref["nsteps"][:, 0])
    # the map is evaluated with the reference's literal expression: even chaotic orbits agree
    cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.3)
    # known answers: fixed point x* = 1 - 1/r for r < 3 => mean = x*, std = 0, KL = 0
    r = prob.par[:, 0]
    fp = r < 2.95
    assert np.allclose(sol.feat[fp, 0, 0], 1 - 1 / r[fp], atol=1e-9)
    assert np.all(sol.feat[fp, 0, 1] < 1e-6) and np.all(sol.feat[fp, 0, 2] == 0.0)


def test_c3_second_order_kuramoto(gpu, orc):
    m = gpu
    prob, _ = cases.c3_problem(m, n_mc=600)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    sol = m.solve(prob)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0)
    assert np.array_equal(sol.retcode, ref["retcode"])
    cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.3)


def test_other_systems_small(gpu, orc):
    """kuramoto (all-to-all), non_local_kuramoto_ring, roessler_network, Stuart-Landau at small N."""
    m = gpu
    rng = np.random.default_rng(21)
    N = 8
    checks = []
    # all-to-all Kuramoto
    pars = m.kuramoto_parameters(1.0, rng.normal(0.5, 0.05, N), N)
    checks.append((pars, m.kuramoto, N, "K", (0.5, 3.0), (0.0, 60.0), m.EvalOdeRun(eval_funcs=(m.mean, m.std)), False))
    # non-local ring with exponential kernel
    pars = m.non_local_kuramoto_ring_parameters(N, 0.3, 0.2, lambda x: np.exp(-abs(x)) / (2 * np.pi))
    checks.append((pars, m.non_local_kuramoto_ring, N, "phase_delay", (0.0, 1.4), (0.0, 60.0),
                   m.EvalOdeRun(eval_funcs=(m.mean, m.std), cyclic_setback=True), False))
    # Roessler network on a ring Laplacian
    Nr = 4
    Lp = 2 * np.eye(Nr) - np.roll(np.eye(Nr), 1, 0) - np.roll(np.eye(Nr), -1, 0)
    pars = m.roessler_parameters(0.2 * np.ones(Nr), 0.2 * np.ones(Nr), 5.7 * np.ones(Nr), 0.05, Lp, Nr)
    checks.append((pars, m.roessler_network, 3 * Nr, "K", (0.0, 0.1), (0.0, 30.0),
                   m.EvalOdeRun(eval_funcs=(m.mean, m.std)), False))
    # Stuart-Landau, complex state
    Ns = 8
    A1 = np.roll(np.eye(Ns), 1, 0) + np.roll(np.eye(Ns), -1, 0)
    A2 = A1 + np.roll(np.eye(Ns), 2, 0) + np.roll(np.eye(Ns), -2, 0)
    pars = m.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.9, Ns, 1, 2, A1, A2)
    ev = m.EvalOdeRun(eval_funcs=("mean_real", "std_real", "kl_real", "mean_imag", "std_imag", "kl_imag"))
    checks.append((pars, m.stuart_landau_sathiyadevi, Ns, "eps", (1.8, 2.1), (0.0, 50.0), ev, True))
    for pars, f, ndim, pname, prange, tspan, ev, cplx in checks:
        prng = np.random.default_rng(5)
        if cplx:
            gen = lambda: complex(rng.uniform(-1, 1), rng.uniform(-1, 1))
            u0 = np.zeros(ndim, dtype=complex)
        else:
            gen = lambda: rng.uniform(-1.0, 1.0)
            u0 = np.zeros(ndim)
        rp = m.ODEProblem(f, u0, tspan, pars)
        prob = m.DEMCBBProblem(rp, gen, 96, pars, (pname, lambda: prng.uniform(*prange)), ev, 0.8)
        ic0, par0 = prob.ic.copy(), prob.par.copy()
        sol = m.solve(prob)
        if cplx:
            ic0 = np.stack([ic0.real, ic0.imag], axis=2).reshape(ic0.shape[0], -1)
        ref, ctx = _oracle_ref(m, orc, prob, ic0, par0)
        assert np.array_equal(sol.retcode, ref["retcode"]), f.__name__
        cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.3)


def test_retcodes_maxiters_and_unstable(gpu, orc):
    m = gpu
    prob, _ = cases.c1_problem(m, n_ics=8, seed=2)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    sol = m.solve(prob, maxiters=5, flag_check_inf_nan=False)
    ref, _ = _oracle_ref(m, orc, prob, ic0, par0, maxiters=5)
    assert np.all(sol.retcode == m.abi.RET_MAXITERS) and np.array_equal(sol.retcode, ref["retcode"])
    assert np.isnan(sol.feat).all()
    # logistic with r > 4 escapes to -inf: never NaN in the state, so retcode stays Success and the
    # measures are Inf/NaN (what check_inf_nan reports, eval_mc_prob.jl:264-279) - same as the oracle
    pars = m.logistic_parameters(4.5)
    dp = m.DiscreteProblem(m.logistic, np.array([0.5]), (0.0, 2000.0), pars)
    prob = m.DEMCBBProblem(dp, lambda: 0.3, 4, pars, ("r", [4.5, 4.6]), m.eval_ode_run, 0.8)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    with pytest.warns(UserWarning, match="Inf or NaN"):
        sol = m.solve(prob)
    ref, _ = _oracle_ref(m, orc, prob, ic0, par0)
    assert np.array_equal(sol.retcode, ref["retcode"])
    assert np.array_equal(np.isnan(sol.feat), np.isnan(ref["feat"]))
    assert np.array_equal(np.isinf(sol.feat), np.isinf(ref["feat"]))
    assert len(m.check_inf_nan(sol)["Inf"]) == 8 and len(m.check_inf_nan(sol)["NaN"]) == 16


def test_error_behaviour(gpu):
    m = gpu
    prob, _ = cases.c1_problem(m, n_ics=4)
    with pytest.raises(m.MCBBError, match="fixed-step"):
        m.solve(prob, m.RK4())
    with pytest.raises(m.MCBBError, match="At least two data points"):
        m.dbscan(np.zeros((1, 1)), 1.0, 2)
    with pytest.raises(m.MCBBError, match="eps must be a positive value"):
        m.dbscan(np.zeros((3, 3)), 0.0, 2)
    with pytest.raises(m.MCBBError, match="minpts must be positive"):
        m.dbscan(np.zeros((3, 3)), 1.0, 0)


def test_distance_matrix_matches_oracle(gpu, orc):
    m = gpu
    rng = np.random.default_rng(0)
    for n, glen, gw in [(257, [6, 6, 6, 1], [1, 0.5, 0.5, 1]), (1000, [10, 10, 1], [1.0, 0.0, 1.0]),
                        (63, [1], [2.0]), (130, [33, 30, 1, 1], [0.25, 1.0, 0.3, 1.0])]:
        X = np.asfortranarray(rng.normal(size=(n, sum(glen))))
        gl, w = np.asarray(glen, dtype=np.int32), np.asarray(gw, dtype=np.float64)
        D = np.zeros((n, n), order="F")
        m.check(m.lib().mcbb_distance_matrix(X.ctypes.data_as(m.abi.c_double_p), n, len(gl),
                                             gl.ctypes.data_as(m.abi.c_int32_p), w.ctypes.data_as(m.abi.c_double_p),
                                             D.ctypes.data_as(m.abi.c_double_p)))
        Dref = orc.distance_matrix(X, gl, w)
        assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
        assert np.allclose(D, Dref, rtol=1e-12, atol=1e-13)


def _dbscan_cases():
    rng = np.random.default_rng(42)
    out = []
    # blobs + noise, generic
    pts = np.concatenate([rng.normal(c, 0.3, size=(60, 2)) for c in [(0, 0), (4, 0), (0, 4)]] +
                         [rng.uniform(-3, 7, size=(40, 2))])
    out.append(("blobs", pts, 0.5, 5))
    # integer lattice: many exact ties d == eps (strict '<' matters) and border points adjacent to two clusters
    g = np.array([(i, j) for i in range(12) for j in range(12)], dtype=float)
    g = g[rng.permutation(len(g))][:110]
    out.append(("lattice_ties", g, 1.0, 3))
    out.append(("lattice_ties2", g, 2.0, 4))
    out.append(("minpts1", pts[:50], 0.4, 1))
    out.append(("all_noise", pts[:80], 1e-6, 3))
    out.append(("single_cluster", pts[:80], 100.0, 3))
    # chain where border points touch two clusters
    chain = np.array([[0, 0], [0.1, 0], [0.2, 0], [1.0, 0], [1.8, 0], [1.9, 0], [2.0, 0]], dtype=float)
    out.append(("two_clusters_shared_border", chain, 0.85, 3))
    return out


@pytest.mark.parametrize("name,pts,eps,minpts", _dbscan_cases(), ids=[c[0] for c in _dbscan_cases()])
def test_dbscan_bit_exact_vs_oracle(gpu, orc, name, pts, eps, minpts):
    """DBSCAN labels from the oracle's distance matrix must be bit-exact (seeds, assignments, counts), for the
    dense path and for the fused feature path."""
    m = gpu
    X = np.asfortranarray(pts)
    gl = np.array([X.shape[1]], dtype=np.int32)
    gw = np.array([1.0])
    Dref = orc.distance_matrix(X, gl, gw)
    ref = orc.dbscan(Dref, eps, minpts)
    res = m.dbscan(Dref, eps, minpts)
    assert np.array_equal(res.assignments, ref["assignments"])
    assert np.array_equal(res.seeds, ref["seeds"]) and np.array_equal(res.counts, ref["counts"])
    fm = m.B200FeatureMatrix(X, gl, gw, [1.0], False)
    res2 = m.dbscan(fm, eps, minpts)
    assert np.array_equal(res2.assignments, ref["assignments"])
    assert np.array_equal(res2.seeds, ref["seeds"]) and np.array_equal(res2.counts, ref["counts"])


def test_dbscan_random_matrices_property(gpu, orc):
    """Parallel restatement == sequential transcription on random symmetric D with quantised (tie-rich) values."""
    m = gpu
    rng = np.random.default_rng(9)
    for trial in range(12):
        n = int(rng.integers(2, 300))
        U = np.round(rng.random((n, n)) * 20) / 20
        D = np.triu(U, 1)
        D = np.asfortranarray(D + D.T)
        eps = float(rng.choice([0.05, 0.1, 0.15, 0.3]))
        minpts = int(rng.integers(1, 8))
        ref = orc.dbscan(D, eps, minpts)
        res = m.dbscan(D, eps, minpts)
        assert np.array_equal(res.assignments, ref["assignments"]), (trial, n, eps, minpts)
        assert np.array_equal(res.seeds, ref["seeds"]) and np.array_equal(res.counts, ref["counts"])


def test_end_to_end_c1_cluster_membership(gpu, orc):
    """solve -> distance_matrix -> dbscan -> cluster_membership against the oracle pipeline (ARI >= 0.99)."""
    from sklearn.metrics import adjusted_rand_score
    m = gpu
    prob, weights = cases.c1_problem(m, n_ics=300, ks=np.arange(0.25, 3.01, 0.25), tspan=(0.0, 200.0),
                                     eval_funcs=(m.mean, m.std))  # KL excluded: its bin-count quirk is 1-ulp sensitive
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    tolkw = dict(reltol=1e-9, abstol=1e-9)
    sol = m.solve(prob, **tolkw)
    D = m.distance_matrix(sol, prob, weights)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **tolkw)
    rsol = m.DEMCBBSol(ref["feat"], None, None, ref["retcode"], ref["nsteps"], 400, None)
    Xr, gl, gw = m.feature_groups(rsol, prob, weights)
    Dref = orc.distance_matrix(Xr, gl, gw)
    eps = float(np.median(np.sort(Dref, axis=0)[15]))
    res = m.dbscan(D, eps, 5)
    rdb = orc.dbscan(Dref, eps, 5)
    assert np.allclose(D.data, Dref, rtol=1e-5, atol=1e-6)
    assert adjusted_rand_score(rdb["assignments"], res.assignments) >= 0.99
    cm = m.cluster_membership(prob, res, 0.5, 0.25)
    mesh, rcm = orc.cluster_membership(prob.par[:, 0], res.assignments, len(res.seeds), 0.5, 0.25)
    assert np.allclose(cm.par, mesh) and np.allclose(cm.data, rcm, equal_nan=True)


def test_fused_dbscan_large_roundtrip_properties(gpu):
    """Full-size style check without an oracle pass: the fused path on 20000 points must equal the dense path on
    the library's own materialised matrix (same distance function), and labels must be permutation-consistent."""
    m = gpu
    rng = np.random.default_rng(4)
    n, F = 20000, 12
    centers = rng.normal(0, 4, size=(6, F))
    X = np.asfortranarray(centers[rng.integers(0, 6, n)] + rng.normal(0, 0.4, size=(n, F)))
    gl = np.array([5, 6, 1], dtype=np.int32)
    gw = np.array([1.0, 0.5, 1.0])
    fm = m.B200FeatureMatrix(X, gl, gw, [1, 0.5, 1], False)
    res = m.dbscan(fm, 2.2, 10)
    D = np.zeros((n, n), order="F")
    m.check(m.lib().mcbb_distance_matrix(X.ctypes.data_as(m.abi.c_double_p), n, 3, gl.ctypes.data_as(m.abi.c_int32_p),
                                         gw.ctypes.data_as(m.abi.c_double_p), D.ctypes.data_as(m.abi.c_double_p)))
    res_d = m.dbscan(D, 2.2, 10)
    assert np.array_equal(res.assignments, res_d.assignments)
    assert np.array_equal(res.seeds, res_d.seeds) and np.array_equal(res.counts, res_d.counts)
    assert res.counts.sum() + m.cluster_n_noise(res) == n
    assert np.all(np.diff(res.seeds) > 0)  # clusters numbered by ascending seed
    assert np.array_equal(res.assignments[res.seeds - 1], np.arange(1, len(res.seeds) + 1))


def test_c5_batched_kuramoto_network_n100(gpu, orc):
    """C5 shape (Kuramoto N=100, Erdos-Renyi p=0.2, K~U(0,8), mean/std with cyclic_setback) through the
    block-cooperative batched kernel, against the oracle's literal N^2-sin right-hand side.

    At the default tolerances (reltol 1e-3) large-K runs are stability-limited: the controller keeps EEst
    near 1, every accept/reject decision is a discontinuity, and ANY two FP64 implementations (different
    rounding of the same formula) drift apart at the 1e-3 level.  Those runs are identified by the oracle's
    own sensitivity probe and excluded; the converged variant below checks every run."""
    m = gpu
    prob, _ = cases.c5_problem(m, 200, tspan=(0.0, 200.0))
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    sol = m.solve(prob)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0)
    assert np.array_equal(sol.retcode, ref["retcode"])
    rel = np.abs(sol.nsteps[:, 0] - ref["nsteps"][:, 0]) / ref["nsteps"][:, 0]
    assert rel.max() < 0.25, rel.max()
    stats = cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.05)
    print("c5 default-tol parity", stats)


def test_c5_batched_converged(gpu, orc):
    """Same problem integrated to convergence (reltol = abstol = 1e-9): every run must match to 1e-6."""
    m = gpu
    prob, _ = cases.c5_problem(m, 64, tspan=(0.0, 100.0))
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-9, abstol=1e-9, maxiters=10**7)
    sol = m.solve(prob, **kw)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **kw)
    assert np.array_equal(sol.retcode, ref["retcode"])
    stats = cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.9)
    print("c5 converged parity", stats)


def test_c5_full_size_properties(gpu):
    """The bench workload at its full single-GPU size (C5: 125 000 runs, Kuramoto N = 100, tspan (0, 1000), default
    tolerances) through the tcgen05 kernel, checked by size-independent properties instead of an oracle pass:
    * uncoupled runs (K = 0) have the closed form u_i(t) = u_i(0) + w_i t -> mean / std incl. the cyclic setback;
    * a run's result does not depend on where it sits in the ensemble or on its batch neighbours: twins placed far
      apart are bitwise equal, and a shuffled sub-ensemble solved alone reproduces its rows bitwise;
    * every run succeeds and reports a plausible step count."""
    m = gpu
    n = 125000
    N = 100
    A = cases.erdos_renyi(N, 0.2, 10)
    w = np.random.default_rng(11).normal(0.5, 0.1, N)
    pars = m.kuramoto_network_parameters(0.5, w, N, A)
    rp = m.ODEProblem(m.kuramoto_network, np.zeros(N), (0.0, 1000.0), pars)
    ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std), cyclic_setback=True)
    rng = np.random.default_rng(12)
    prob = m.DEMCBBProblem(rp, lambda: 0.0, 1, pars, ("K", lambda: 0.0), ev, 0.9)  # shell; arrays set below
    prob.ic = np.asfortranarray(rng.uniform(-np.pi, np.pi, size=(n, N)))
    prob.par = np.asfortranarray(rng.uniform(0.0, 8.0, size=(n, 1)))
    prob.N_mc = n
    free = np.arange(0, n, 1000)
    prob.par[free, 0] = 0.0
    src, dst = np.arange(7, n, 2500), np.arange(7, n, 2500)[::-1] + 1  # 50 twins, far apart
    prob.ic[dst], prob.par[dst] = prob.ic[src], prob.par[src]
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    before = m.kernel_launch_count()
    feat, _, _, ret, ns = m.custom_solve_b200(prob, ts)
    assert m.kernel_launch_count() - before == 1  # one launch of the batched kernel for the whole ensemble
    assert np.all(ret == 0) and np.isfinite(feat).all()
    assert ns[:, 0].min() >= 3 and ns[:, 0].max() < 100000 and np.all(ns[:, 1] <= ns[:, 0])
    # closed form of the uncoupled runs
    t = ts[1:]
    x1 = prob.ic[free] + w[None, :] * t[0]
    v = np.mod(x1, 2 * np.pi)
    v = np.where(v > np.pi, v - 2 * np.pi, v)
    want_mean = v + w[None, :] * (t - t[0]).mean()
    want_std = np.abs(w)[None, :] * t.std(ddof=1)
    near_cut = np.abs(np.abs(v) - np.pi) < 1e-6  # a sample within rounding of the branch cut may wrap either way
    assert np.abs(feat[free, :, 0] - want_mean)[~near_cut].max() < 1e-7
    assert np.abs(feat[free, :, 1] - want_std).max() < 1e-7
    # twins
    assert np.array_equal(feat[src], feat[dst]) and np.array_equal(ns[src], ns[dst])
    # a shuffled sub-ensemble solved alone
    pick = rng.permutation(n)[:6000]
    sub = m.DEMCBBProblem(rp, lambda: 0.0, 1, pars, ("K", lambda: 0.0), ev, 0.9)
    sub.ic, sub.par, sub.N_mc = np.asfortranarray(prob.ic[pick]), np.asfortranarray(prob.par[pick]), len(pick)
    f2, _, _, r2, n2 = m.custom_solve_b200(sub, ts)
    assert np.array_equal(f2, feat[pick]) and np.array_equal(n2, ns[pick]) and np.all(r2 == 0)


@pytest.mark.parametrize("N,filt", [(30, None), (40, None), (64, None), (97, [0, 5, 63, 64, 95, 96]), (128, None)])
def test_tensor_core_kernel_network_sizes(gpu, orc, N, filt):
    """The tcgen05 integrator maps oscillator r to row r of a 128-row MMA tile: sizes that leave whole warps idle
    (N = 40, 64), a size with a one-row last warp plus a state_filter straddling the warp edges (N = 97), and the full
    tile (N = 128) — converged tolerances, every run must match the oracle to 1e-6.  N = 30 is below the tcgen05
    kernel's range and exercises its mma.sync predecessor (csrc/batched_imma.cu)."""
    m = gpu
    prob, _ = cases.c5_problem(m, 48, N=N, seed=20 + N, tspan=(0.0, 60.0))
    if filt is not None:
        prob.eval_ode_func = m.EvalOdeRun(state_filter=[f + 1 for f in filt], eval_funcs=(m.mean, m.std), cyclic_setback=True)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-9, abstol=1e-9, maxiters=10**7)
    sol = m.solve(prob, **kw)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **kw)
    assert np.array_equal(sol.retcode, ref["retcode"])
    assert sol.feat.shape[1] == (N if filt is None else len(filt))
    cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.9)


def test_batched_ring_and_filter(gpu, orc):
    """non_local_kuramoto_ring N=40 (phase_delay swept) and a state_filter subset on the batched kernel."""
    m = gpu
    N = 40
    rng = np.random.default_rng(8)
    pars = m.non_local_kuramoto_ring_parameters(N, 0.2, 0.3, lambda x: np.exp(-2 * abs(x)) / (2 * np.pi))
    rp = m.ODEProblem(m.non_local_kuramoto_ring, np.zeros(N), (0.0, 40.0), pars)
    ev = m.EvalOdeRun(state_filter=list(range(3, 33, 2)), eval_funcs=(m.mean, m.std), cyclic_setback=True)
    prng = np.random.default_rng(9)
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 130, pars,
                           ("phase_delay", lambda: prng.uniform(0.0, 0.6)), ev, 0.8)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-9, abstol=1e-9, maxiters=10**7)
    sol = m.solve(prob, **kw)
    assert sol.feat.shape == (130, 15, 2)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **kw)
    assert np.array_equal(sol.retcode, ref["retcode"])
    cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.3)


def test_solve_is_deterministic(gpu):
    """test/io.jl:53-62 expects a re-solve to be element-wise equal."""
    m = gpu
    for build in (lambda: cases.c1_problem(m, n_ics=64, seed=5)[0], lambda: cases.c5_problem(m, 70, tspan=(0.0, 60.0))[0]):
        a = m.solve(build())
        b = m.solve(build())
        assert np.array_equal(a.feat, b.feat, equal_nan=True) and np.array_equal(a.nsteps, b.nsteps)


def test_all_measures_correlation_ecdf_and_sum(gpu, orc):
    """test/ode_prob_example.jl:49-64 `eval_ode_run_all_test`: [mean, std] + correlation_ecdf + global sum, then the
    distance matrix over per-dim, matrix, global and parameter groups."""
    m = gpu
    prob, _ = cases.c1_problem(m, n_ics=60, seed=13, N=5)
    prob.eval_ode_func = m.EvalOdeRun(eval_funcs=(m.mean, m.std), matrix_eval_funcs=[m.correlation_ecdf],
                                      global_eval_funcs=["sum"])
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    sol = m.solve(prob, reltol=1e-9, abstol=1e-9, flag_check_inf_nan=False)
    assert (sol.N_meas, sol.N_meas_dim, sol.N_meas_matrix, sol.N_meas_global) == (4, 2, 1, 1)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, reltol=1e-9, abstol=1e-9)
    cases.compare_features(sol.feat, ref, orc, *ctx)
    # the ECDF is a step function of the correlations: equal wherever no |cor| sits within rounding of a bin edge
    # (fully synchronised runs have |cor| == 1 everywhere: every entry falls outside [0,1) and the ECDF is 0/0 = NaN,
    # in the reference as well)
    both_nan = np.isnan(sol.matrix) & np.isnan(ref["matrix"])
    with np.errstate(invalid="ignore"):
        same = np.all((np.abs(sol.matrix - ref["matrix"]) < 1e-12) | both_nan, axis=1)
    assert same.mean() >= 0.95, same.mean()
    ok = ~np.isnan(sol.matrix[:, -1])
    assert ok.sum() >= 10 and np.allclose(sol.matrix[ok, -1], 1.0) and np.all(np.diff(sol.matrix[ok], axis=1) >= 0)
    assert np.allclose(sol.glob, ref["glob"], rtol=1e-8)
    weights = [1.0, 0.5, 0.5, 0.25, 1.0]
    D = m.distance_matrix(sol, prob, weights)
    X, gl, gw = m.feature_groups(sol, prob, weights)
    assert list(gl) == [5, 5, 30, 1, 1]
    assert np.allclose(D.data, orc.distance_matrix(X, gl, gw), rtol=1e-12, atol=1e-13, equal_nan=True)


def test_correlation_hist_and_ecdf_on_power_grid_frequencies(gpu, orc):
    """correlation_hist / correlation_ecdf (eval_mc_prob.jl:462-488) where the |cor| values actually spread over the
    bins: the ten frequencies of the second-order Kuramoto grid during their transient (90 off-diagonal pairs)."""
    m = gpu
    out = {}
    for f in (m.correlation_hist, m.correlation_ecdf):
        prob, _ = cases.c3_problem(m, n_mc=60, seed=5, N=10, tspan=(0.0, 60.0), tail=0.5)
        prob.eval_ode_func = m.EvalOdeRun(state_filter=list(range(11, 21)), eval_funcs=(m.mean, m.std),
                                          matrix_eval_funcs=[f])
        ic0, par0 = prob.ic.copy(), prob.par.copy()
        sol = m.solve(prob, reltol=1e-9, abstol=1e-9)
        ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, reltol=1e-9, abstol=1e-9)
        cases.compare_features(sol.feat, ref, orc, *ctx)
        out[f] = (sol.matrix, ref["matrix"])
    w, w_ref = out[m.correlation_hist]
    e, e_ref = out[m.correlation_ecdf]
    assert np.all(w == np.rint(w)) and np.all(w.sum(axis=1) == 90) and np.all(w >= 0)
    assert (w[:, :-1].sum(axis=1) > 0).sum() >= 40  # not just the last bin
    # a |cor| within rounding of a bin edge may land on either side: exact for (nearly) every run
    assert np.all(w == w_ref, axis=1).mean() >= 0.95, np.all(w == w_ref, axis=1).mean()
    assert np.abs(w - w_ref).sum(axis=1).max() <= 4
    assert np.all(np.abs(e - e_ref) < 1e-15, axis=1).mean() >= 0.95
    assert np.array_equal(e, np.cumsum(w, axis=1) / 90.0)  # ecdf_hist of the weights, eval_clustering.jl:759-762


def test_complex_state_correlation_measures(gpu, orc):
    """correlation_hist / correlation_ecdf of COMPLEX states (the Stuart-Landau configuration of the paper notebooks):
    |cor| of the complex series, from the device's online co-moments of the (re, im) channels.  Measured over the
    transient (T = 10, nothing discarded): once the amplitudes synchronise every |cor| sits within rounding of 1, the
    open end of the last bin, and the oracle itself flips bins under 1e-12 perturbations of the initial condition."""
    m = gpu
    Ns = 8
    A1 = np.roll(np.eye(Ns), 1, 0) + np.roll(np.eye(Ns), -1, 0)
    A2 = A1 + np.roll(np.eye(Ns), 2, 0) + np.roll(np.eye(Ns), -2, 0)
    kw = dict(reltol=1e-9, abstol=1e-9)
    out = {}
    for f, filt in ((m.correlation_hist, None), (m.correlation_ecdf, None), (m.correlation_hist, [1, 3, 4, 8])):
        rng, prng = np.random.default_rng(31), np.random.default_rng(32)
        pars = m.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.9, Ns, 1, 2, A1, A2)
        ev = m.EvalOdeRun(state_filter=filt, eval_funcs=("mean_real", "std_real", "mean_imag", "std_imag"),
                          matrix_eval_funcs=[f])
        rp = m.ODEProblem(m.stuart_landau_sathiyadevi, np.zeros(Ns, dtype=complex), (0.0, 10.0), pars)
        prob = m.DEMCBBProblem(rp, lambda: complex(rng.uniform(-1, 1), rng.uniform(-1, 1)), 64, pars,
                               ("eps", lambda: prng.uniform(1.8, 2.1)), ev, 0.0)
        ic0, par0 = prob.ic.copy(), prob.par.copy()
        sol = m.solve(prob, **kw)
        ic0 = np.stack([ic0.real, ic0.imag], axis=2).reshape(ic0.shape[0], -1)
        ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **kw)
        assert np.array_equal(sol.retcode, ref["retcode"])
        cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.8)
        out[(f, filt is None)] = (sol.matrix, ref["matrix"])
    for key, npairs in (((m.correlation_hist, True), 56), ((m.correlation_hist, False), 12)):
        w, w_ref = out[key]
        assert np.all(w == np.rint(w)) and np.all(w >= 0) and np.all(w.sum(axis=1) == npairs)
        same = np.all(w == w_ref, axis=1)
        assert same.mean() >= 0.95, (same.mean(), w[~same][:2], w_ref[~same][:2])
        assert np.abs(w - w_ref).sum(axis=1).max() <= 4
    w = out[(m.correlation_hist, True)][0]
    assert (w[:, :-1].sum(axis=1) > 0).all()  # spread over the bins, not just the last one
    e, e_ref = out[(m.correlation_ecdf, True)]
    assert np.array_equal(e, np.cumsum(w, axis=1) / 56.0)
    assert np.all(np.abs(e - e_ref) < 1e-15, axis=1).mean() >= 0.95


def test_get_trajectory_and_end_state_export(gpu, orc):
    """get_trajectory (eval_clustering.jl:1484-1500): the samples `sol.solve_command(prob_i)` saves at t_save, for a
    given run and for a random member of a cluster; plus the end-state export (sol[end]) of a whole ensemble."""
    m = gpu
    prob, weights = cases.c1_problem(m, n_ics=40, seed=23)
    sol = m.solve(prob)  # sorted: prob.ic / prob.par now follow the sorted order
    psys = prob.pars.pack(prob.par_names)
    o = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    for i_run in (0, 57, prob.N_mc - 1):
        tr = m.get_trajectory(prob, sol, i_run)
        ref, ret, _, nsv = orc.trajectory(psys, o, prob.ic[i_run], prob.par[i_run], ts)
        assert tr.retcode == ret == 0 and nsv == 400 and np.array_equal(tr.t, ts)
        assert np.asarray(tr).shape == (6, 400) and len(tr) == 400
        assert np.allclose(np.asarray(tr), ref, rtol=1e-9, atol=1e-9)
        # the measures of the run are the statistics of exactly these samples (first one dropped)
        assert np.allclose(sol.feat[i_run, :, 0], np.asarray(tr)[:, 1:].mean(axis=1), rtol=1e-12, atol=1e-12)
    D = m.distance_matrix(sol, prob, weights)
    db = m.dbscan(D, 0.5, 4)
    k = 2 if len(db.counts) >= 1 else 1
    tr, sub, i_run = m.get_trajectory(prob, sol, db, k, only_sol=False, rng=np.random.default_rng(0))
    assert db.assignments[i_run] == k - 1 and sub.N_mc == 1 and np.array_equal(sub.ic[0], prob.ic[i_run])
    with pytest.raises(m.MCBBError):
        m.get_trajectory(prob, sol, db, len(db.counts) + 5)
    # whole-ensemble exports through the C-ABI: every sample, and only the last one (the ic layout)
    ts20 = ts[:20]
    *_, ret_a, _, all_s = m.api._solve_b200(prob, ts20, None, m.abi.SAVE_ALL, {})
    *_, ret_l, _, last = m.api._solve_b200(prob, ts20, None, m.abi.SAVE_LAST, {})
    assert all_s.shape == (prob.N_mc, 6, 20) and last.shape == (prob.N_mc, 6)
    assert np.array_equal(all_s[:, :, -1], last) and np.all(ret_a == 0) and np.all(ret_l == 0)
    # a failed run exports NaN for the samples it never reached
    *_, ret_f, _, cut = m.api._solve_b200(prob, ts, None, m.abi.SAVE_ALL, dict(maxiters=8))
    assert np.all(ret_f == m.abi.RET_MAXITERS) and np.isnan(cut[:, :, -1]).all()
    # complex states come back complex
    Ns = 8
    A1 = np.roll(np.eye(Ns), 1, 0) + np.roll(np.eye(Ns), -1, 0)
    pars = m.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.9, Ns, 1, 1, A1, A1)
    rng, prng = np.random.default_rng(3), np.random.default_rng(4)
    rp = m.ODEProblem(m.stuart_landau_sathiyadevi, np.zeros(Ns, dtype=complex), (0.0, 10.0), pars)
    ev = m.EvalOdeRun(eval_funcs=("mean_real", "std_imag"))
    probc = m.DEMCBBProblem(rp, lambda: complex(rng.uniform(-1, 1), rng.uniform(-1, 1)), 8, pars,
                            ("eps", lambda: prng.uniform(1.8, 2.1)), ev, 0.5)
    solc = m.solve(probc)
    trc = m.get_trajectory(probc, solc, 5)
    ic = np.stack([probc.ic[5].real, probc.ic[5].imag], axis=1).reshape(-1)
    oc = m.abi.solver_opts(m.abi.ALG_TSIT5, probc.p.tspan)
    refc, _, _, _ = orc.trajectory(probc.pars.pack(probc.par_names), oc, ic, probc.par[5],
                                   m.tsave_array(probc.p, 400, 0.5))
    assert np.iscomplexobj(trc.u) and trc.u.shape == (Ns, 400)
    assert np.allclose(trc.u, refc[0::2] + 1j * refc[1::2], rtol=1e-9, atol=1e-9)


def test_continuation_response_analysis(gpu, orc):
    """eval_ode_run with the response analysis (eval_mc_prob.jl:202-257): after each run, three short runs restart
    from its end state at p - eps, p, p + eps; D[1,2] and D[2,3] of their measures are appended as scalars.  The
    oracle side is composed from its single-trajectory, featurize and distance functions."""
    m = gpu
    eps, bounds, w_inner = 0.25, (1.0, 3.0), [1.0, 0.5, 1.0]
    for return_pm in (True, False):
        prob, _ = cases.c1_problem(m, n_ics=12, seed=29, eval_funcs=(m.mean, m.std))
        prob.eval_ode_func = m.EvalOdeRun(eval_funcs=(m.mean, m.std),
                                          continuation=m.ContinuationResponse(eps, bounds, w_inner, N_t=200,
                                                                              return_pm=return_pm))
        sol = m.solve(prob)
        n_resp = 2 if return_pm else 1
        assert (sol.N_meas_dim, sol.N_meas_global, sol.N_meas) == (2, n_resp, 2 + n_resp)
        assert sol.glob.shape == (prob.N_mc, n_resp) and len(sol[0]) == 2 + n_resp
        # oracle composition, in the sorted run order solve() leaves behind
        psys, pfeat = prob.pars.pack(prob.par_names), m.EvalOdeRun(eval_funcs=(m.mean, m.std)).pack()
        o = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
        ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
        n = prob.N_mc
        end = np.stack([orc.trajectory(psys, o, prob.ic[i], prob.par[i], ts)[0][:, -1] for i in range(n)])
        p0 = prob.par[:, 0]
        par3 = np.stack([np.maximum(p0 - eps, bounds[0]), p0, np.minimum(p0 + eps, bounds[1])], axis=1)
        assert (par3[:, 0] == p0).any() and (par3[:, 2] == p0).any()  # both bounds clip somewhere
        t0, t1 = prob.p.tspan
        o2 = m.abi.solver_opts(m.abi.ALG_TSIT5, (t0, t0 + 0.15 * (t1 - t0)))
        rp2 = m.ODEProblem(m.kuramoto_network, np.zeros(6), (t0, t0 + 0.15 * (t1 - t0)), prob.pars)
        ts2 = m.tsave_array(rp2, 200, 0.1)
        ref3 = orc.solve_features(psys, o2, pfeat, np.repeat(end, 3, axis=0), par3.reshape(-1, 1), ts2)["feat"]
        X = np.concatenate([ref3[:, :, 0], ref3[:, :, 1], par3.reshape(-1, 1)], axis=1)
        d = np.zeros((n, 2))
        for i in range(n):
            D3 = orc.distance_matrix(X[3 * i:3 * i + 3], [6, 6, 1], w_inner)
            d[i] = D3[0, 1], D3[1, 2]
        want = d if return_pm else d.mean(axis=1, keepdims=True)
        assert np.allclose(sol.glob, want, rtol=1e-6, atol=1e-9), np.abs(sol.glob - want).max()
        # the response vanishes only where the bound clipped the parameter onto p itself ... and not elsewhere
        if return_pm:
            assert np.all(sol.glob[par3[:, 0] != p0, 0] > 0)
    # the responses enter distance_matrix as global measures
    D = m.distance_matrix(sol, prob, [1.0, 0.5, 0.25, 1.0])
    assert D.data.shape == (prob.N_mc, prob.N_mc) and np.isfinite(D.data).all()


def test_solver_parameter_var(gpu, orc):
    """test/solver_parametervar.jl: Kuramoto network N=20, `SolverParameterVar(:reltol, ()->rand(Uniform(1e-6,1e-3)))`,
    random ICs, default eval_ode_run — every run equals the oracle solved alone with that run's reltol.  Also a
    per-run fixed step (RK4 dt) and that the request is one-shot (the next plain solve is unaffected)."""
    m = gpu
    rng, vgen = np.random.default_rng(51), np.random.default_rng(52)
    N = 20
    pars = m.kuramoto_network_parameters(0.5, rng.normal(0.5, 0.005, N), N, np.ones((N, N)))
    rp = m.ODEProblem(m.kuramoto_network, np.zeros(N), (0.0, 40.0), pars)
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 100, pars,
                           m.SolverParameterVar("reltol", lambda: vgen.uniform(1e-6, 1e-3)), m.eval_ode_run, 0.9)
    sol = m.solve(prob, m.Tsit5())
    assert np.all(np.diff(prob.par[:, 0]) >= 0) and np.all(sol.retcode == 0)  # sorted by the solver option
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    tol_ok = 0
    for i in range(0, prob.N_mc, 7):
        oi = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan, reltol=float(prob.par[i, 0]))
        ref = orc.solve_features(psys, oi, pfeat, prob.ic[i:i + 1], prob.par[i:i + 1], ts)
        assert np.array_equal(sol.nsteps[i], ref["nsteps"][0]), i
        tol_ok += np.allclose(sol.feat[i, :, :2], ref["feat"][0, :, :2], rtol=1e-6, atol=1e-7)
    assert tol_ok >= 13  # 15 runs checked; K = 0.5 is incoherent, a run may amplify rounding differences
    assert len(set(sol.nsteps[:, 0].tolist())) > 10  # tighter reltol -> more steps
    assert sol.nsteps[:10, 0].mean() > sol.nsteps[-10:, 0].mean()
    # one-shot: a plain parameter sweep right after is untouched by the previous request
    prob2, _ = cases.c1_problem(m, n_ics=8, seed=3)
    ic0, par0 = prob2.ic.copy(), prob2.par.copy()
    ref2, _ = _oracle_ref(m, orc, prob2, ic0, par0)
    assert np.array_equal(m.solve(prob2).nsteps, ref2["nsteps"])
    # per-run fixed step
    prob3 = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 5, pars,
                            m.SolverParameterVar("dt", [0.01, 0.02, 0.04]), m.EvalOdeRun(eval_funcs=(m.mean, m.std)), 0.9)
    with pytest.raises(m.MCBBError):
        m.solve(prob3, m.RK4(), dt=0.03)  # both varied and given
    sol3 = m.solve(prob3, m.RK4())
    assert np.all(sol3.retcode == 0) and sol3.feat.shape == (15, N, 2)
    want = np.rint(40.0 / prob3.par[:, 0]).astype(np.int64)  # fixed steps of the run's own dt
    assert np.all(np.abs(sol3.nsteps[:, 0] - want) <= 1), (sol3.nsteps[:, 0], want)


def test_two_swept_parameters_end_to_end(gpu, orc):
    """MultiDimParameterVar (setup_mc_prob.jl:176-236): coupling AND damping of the power grid drawn per run; solve,
    distance matrix with one weight per parameter, DBSCAN and the two-parameter cluster_membership."""
    m = gpu
    N = 6
    rng = np.random.default_rng(9)
    E = cases.ring_incidence(N, [(0, 3)])
    pars = m.second_order_kuramoto_parameters(N, 0.1, 1.0, E, np.array([1.0, -1.0] * (N // 2)))
    rp = m.ODEProblem(m.second_order_kuramoto, np.zeros(2 * N), (0.0, 60.0), pars)
    ev = m.EvalOdeRun(state_filter=list(range(N + 1, 2 * N + 1)), eval_funcs=(m.mean, m.std))
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 300, pars,
                           [("coupling", lambda: rng.uniform(0.5, 4.0)), ("damping", lambda: rng.uniform(0.05, 0.5))],
                           ev, 0.5)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-8, abstol=1e-8)
    sol = m.solve(prob, **kw)
    assert np.all(np.diff(prob.par[:, 0]) >= 0)  # sorted by the FIRST parameter, setup_mc_prob.jl:520-536
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **kw)
    assert ctx[0].desc.n_par == 2 and np.array_equal(sol.nsteps, ref["nsteps"])
    cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.5)
    weights = [1.0, 0.5, 1.0, 2.0]
    D = m.distance_matrix(sol, prob, weights)
    X, gl, gw = m.feature_groups(sol, prob, weights)
    assert list(gl) == [N, N, 1, 1] and np.allclose(D.data, orc.distance_matrix(X, gl, gw), rtol=1e-12, atol=1e-13)
    db = m.dbscan(D, float(np.median(m.k_dist(D, 4))), 4)
    cm = m.cluster_membership(prob, db, [1.0, 0.15], [0.5, 0.075])
    assert cm.multidim_flag and cm.data.shape[2] == len(db.seeds) + 1
    filled = ~np.isnan(cm.data[..., 0])
    assert filled.any() and np.allclose(cm.data[filled].sum(axis=-1), 1.0)


def test_k_dist_and_knn_helpers(gpu):
    """k_dist / KNN_dist / KNN_dist_relative (eval_clustering.jl:1504-1553) against numpy on the same matrix."""
    m = gpu
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(700, 3))
    D = np.abs(pts[:, None, :] - pts[None, :, :]).sum(-1)
    Ds = np.sort(D, axis=1)
    assert np.array_equal(m.k_dist(D, 4), np.sort(Ds[:, 3])[::-1])
    assert np.allclose(m.KNN_dist(D, 7)[:, 0], Ds[:, :7].sum(1), rtol=1e-14)
    K = int(round(700 * 0.005))
    assert np.allclose(m.KNN_dist_relative(D)[:, 0], Ds[:, :K].sum(1), rtol=1e-14)
    with pytest.raises(m.MCBBError, match="k out of range"):
        m.k_dist(D, 0)
    # the same from the features, D never materialised: row blocks (ragged last block) sorted one block at a time
    X = np.asfortranarray(np.concatenate([pts, rng.normal(size=(700, 2))], axis=1))
    gl, gw = np.array([3, 1, 1], dtype=np.int32), np.array([1.0, 0.5, 2.0])
    fm = m.B200FeatureMatrix(X, gl, gw, [1.0, 0.5, 2.0], False)
    Dm = m.lib().mcbb_distance_matrix
    Dd = np.zeros((700, 700), order="F")
    m.check(Dm(X.ctypes.data_as(m.abi.c_double_p), 700, 3, gl.ctypes.data_as(m.abi.c_int32_p),
               gw.ctypes.data_as(m.abi.c_double_p), Dd.ctypes.data_as(m.abi.c_double_p)))
    assert np.array_equal(m.k_dist(fm, 4), m.k_dist(Dd, 4))
    assert np.array_equal(m.KNN_dist(fm, 9), m.KNN_dist(Dd, 9))
    assert np.array_equal(m.KNN_dist_relative(fm, 0.01), m.KNN_dist_relative(Dd, 0.01))
    for br in (1, 64, 333, 700, 5000):
        kth, cum = m.api._knn(fm, 5, 11, block_rows=br)
        kd, cd = m.api._knn(Dd, 5, 11)
        assert np.array_equal(kth, kd) and np.array_equal(cum, cd), br


def test_failure_handling_repeat_and_inf(gpu):
    """failure_handling = :Repeat redraws the ICs of failed runs, at most 10 times (eval_mc_prob.jl:120-123,
    setup_mc_prob.jl:790-813); :Inf fills the measures of DtLessThanMin / Unstable runs with Inf (:97-119)."""
    m = gpu
    # one K value: the number of step attempts then depends on the initial condition only, so a redraw can succeed
    mk = lambda: cases.c1_problem(m, n_ics=200, seed=21, ks=(2.0,), eval_funcs=(m.mean, m.std))[0]
    prob = mk()
    base = m.solve(prob, sort_results=False)
    limit = int(np.percentile(base.nsteps[:, 0], 80))  # ~20 % of the runs exceed this many step attempts
    prob = mk()
    plain = m.solve(prob, maxiters=limit, sort_results=False, flag_check_inf_nan=False)
    failed = plain.retcode == m.abi.RET_MAXITERS
    assert 0.02 < failed.mean() < 0.5 and np.isnan(plain.feat[failed]).all()
    prob = mk()
    prob.eval_ode_func = m.EvalOdeRun(eval_funcs=(m.mean, m.std), failure_handling="Repeat")
    ic_before = prob.ic.copy()
    sol = m.solve(prob, maxiters=limit, sort_results=False)
    assert np.all(sol.retcode == m.abi.RET_SUCCESS) and np.all(np.isfinite(sol.feat))
    changed = np.any(prob.ic != ic_before, axis=1)
    assert np.array_equal(changed, failed)  # exactly the failed runs got new initial conditions
    # non-random ICs cannot be repeated
    N = 6
    pars = prob.pars
    rp = m.ODEProblem(m.kuramoto_network, np.zeros(N), (0.0, 40.0), pars)
    ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std), failure_handling="Repeat")
    prob2 = m.DEMCBBProblem(rp, [np.array([0.1, 2.0])] * N, pars, ("K", [1.0, 2.0]), ev, 0.9)
    with pytest.raises(m.MCBBError, match="IC are not random"):
        m.solve(prob2, maxiters=5, flag_check_inf_nan=False)


def test_c4_stuart_landau_n100_global_state_path(gpu, orc):
    """C4's system at its real size with SURVEY §8(d)'s inputs: Stuart-Landau N=100 (200 real state dimensions), A_1 / A_2
    Watts-Strogatz small worlds (degree 2 and 44, p_r = 0.04, seed 7), eps ~ U(1.8, 2.1), z0 ~ U(-1,1) + i U(-1,1),
    measures mean/std/KL of real and imaginary parts + correlation_ecdf.  Short time span at converged tolerances: every
    run against the oracle (the time span of the paper, T = 200, is covered by tests/test_gpu_c4.py)."""
    m = gpu
    prob, _ = cases.c4_problem(m, 48, tspan=(0.0, 20.0))
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-8, abstol=1e-8, maxiters=10 ** 6)
    sol = m.solve(prob, **kw)
    assert sol.feat.shape == (48, 100, 6)
    ic0 = np.stack([ic0.real, ic0.imag], axis=2).reshape(ic0.shape[0], -1)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **kw)
    assert np.array_equal(sol.retcode, ref["retcode"])
    cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.5)
    e, e_ref = sol.matrix, ref["matrix"]  # ECDF of 9900 |cor| values over 30 bins
    assert e.shape == (48, 30) and np.allclose(e[:, -1], 1.0) and np.all(np.diff(e, axis=1) >= 0)
    assert np.abs(e - e_ref).max() <= 4 / 9900 + 1e-15 and (np.abs(e - e_ref).max(axis=1) < 1e-15).mean() >= 0.8


def test_histogram_mode_distance_matrix(gpu, orc):
    """distance_matrix(sol, prob, weights; histograms=true, k_bin) (eval_clustering.jl:187-211, 261-268): Float32
    ECDF rows over common IQR-rule edges, L1 * bin_width.  The reference sums the |ECDF| differences in Float32,
    the device in FP64 -> 1e-6 relative agreement."""
    import warnings
    m = gpu
    prob, _ = cases.c1_problem(m, n_ics=40, seed=17, eval_funcs=(m.mean, m.std))
    sol = m.solve(prob)
    weights = [1.0, 0.5, 1.0]
    for k_bin in (1.0, 0.25):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Dh = m.distance_matrix(sol, prob, weights, histograms=True, k_bin=k_bin)
        assert isinstance(Dh, m.api.DistanceMatrixHist) and len(Dh.hist_edges) == 2
        ref = orc.distance_matrix_hist(sol.feat, prob.par, weights, k_bin=k_bin, nbin_default=50)
        assert np.array_equal(Dh.data, Dh.data.T) and np.all(np.diag(Dh.data) == 0)
        assert np.allclose(Dh.data, ref, rtol=2e-6, atol=1e-9)
        res = m.dbscan(Dh, float(np.median(np.sort(ref, axis=0)[5])), 4)
        rdb = orc.dbscan(Dh.data, float(np.median(np.sort(ref, axis=0)[5])), 4)
        assert np.array_equal(res.assignments, rdb["assignments"])


def test_sparse_and_mmap_distance_matrices(gpu, orc, tmp_path):
    """distance_matrix_sparse -> NonzeroSparseMatrix (entries d < threshold stored as
    d - threshold, compressed columns with ascending rows) and distance_matrix_mmap (2 x Int64 header + rows of
    el_type), both against the oracle's dense matrix; dbscan on the sparse matrix == the sequential algorithm on the
    matrix whose unstored entries read as the threshold (sparse_clustering.jl:126-215), including eps > threshold."""
    m = gpu
    prob, weights = cases.c1_problem(m, n_ics=60, seed=21)
    sol = m.solve(prob)
    X, gl, gw = m.feature_groups(sol, prob, weights)
    Dref = orc.distance_matrix(X, gl, gw)
    n = Dref.shape[0]
    thr = float(np.quantile(Dref, 0.15))
    # ---- sparse, Float64 elements: exact
    S = m.distance_matrix_sparse(sol, prob, weights, sparse_threshold=thr, el_type=np.float64).data
    keep = Dref < thr
    assert np.array_equal(np.diff(S.colptr), keep.sum(axis=0))
    dense = S.toarray()
    assert np.allclose(dense[keep], Dref[keep], rtol=1e-12, atol=1e-12) and np.all(dense[~keep] == thr)
    for c in range(0, n, 97):
        rows = S.rowval[S.colptr[c]:S.colptr[c + 1]]
        assert np.all(np.diff(rows) > 0) and np.array_equal(rows, np.nonzero(keep[:, c])[0])
    assert S[3, 3] == 0.0 and S[0, int(np.argmax(Dref[0]))] == thr
    # ---- DBSCAN on the sparse matrix
    for eps, minpts in ((0.6 * thr, 5), (thr, 8), (1.5 * thr, 3)):
        ref = orc.dbscan(np.asfortranarray(dense), eps, minpts)
        res = m.dbscan(S, eps, minpts)
        assert np.array_equal(res.assignments, ref["assignments"]), (eps, minpts)
        assert np.array_equal(res.seeds, ref["seeds"]) and np.array_equal(res.counts, ref["counts"])
    # ---- sparse, default Float32 elements
    S32 = m.distance_matrix_sparse(sol, prob, weights, sparse_threshold=thr).data
    assert S32.nzval.dtype == np.float32 and np.allclose(S32.toarray()[keep], Dref[keep], rtol=1e-6, atol=1e-6 * thr)
    with pytest.raises(m.MCBBError):
        m.distance_matrix_sparse(sol, prob, weights)  # the reference's default threshold Inf stores d - Inf
    # ---- mmap file
    name = str(tmp_path / "mmap-distance-matrix.bin")
    Dm = m.distance_matrix_mmap(sol, prob, weights, save_name=name, block_rows=77)
    raw = np.fromfile(name, dtype=np.int64, count=2)
    assert list(raw) == [n, n] and os.path.getsize(name) == 16 + 4 * n * n
    assert Dm.data.dtype == np.float32 and np.allclose(np.asarray(Dm.data), Dref, rtol=1e-6, atol=1e-6)
    D64 = m.distance_matrix_mmap(sol, prob, weights, save_name=name, el_type=np.float64)
    assert np.allclose(np.asarray(D64.data), Dref, rtol=1e-12, atol=1e-12)
    again = m.reload_mmap_distance_matrix(D64, name, el_type=np.float64)
    assert np.array_equal(np.asarray(again.data), np.asarray(D64.data)) and again.weights == D64.weights
    res = m.dbscan(np.asarray(D64.data, dtype=np.float64), 0.6 * thr, 5)
    assert np.array_equal(res.assignments, orc.dbscan(Dref, 0.6 * thr, 5)["assignments"])


def test_tensor_core_kernel_default_measures_with_kl(gpu, orc):
    """The reference's default eval_ode_run measures (mean, std, empirical_1D_KL_divergence_hist) on a unit-weight
    Kuramoto network go through the tcgen05 integrator too: the saved samples of the runs in flight are kept in a
    device scratch and binned once mean / std are final.  Converged tolerances; KL compared with its 1-ulp bin-count
    quirk handled by compare_features."""
    m = gpu
    prob, _ = cases.c5_problem(m, 60, N=48, seed=31, tspan=(0.0, 80.0))
    prob.eval_ode_func = m.EvalOdeRun(eval_funcs=(m.mean, m.std, m.empirical_1D_KL_divergence_hist), cyclic_setback=True)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-9, abstol=1e-9, maxiters=10**7)
    before = m.kernel_launch_count()
    sol = m.solve(prob, **kw)
    assert m.kernel_launch_count() - before <= 3  # one integrator launch: no second (replay) pass of the lane kernels
    assert sol.feat.shape == (60, 48, 3)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **kw)
    assert np.array_equal(sol.retcode, ref["retcode"])
    cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.8)
    assert np.isfinite(sol.feat[:, :, 2]).mean() > 0.9


@pytest.mark.parametrize("system", ["ring", "weighted_network"])
def test_fp64_batched_kernel_default_measures_with_kl(gpu, orc, system):
    """The reference's default measures (mean, std, empirical_1D_KL_divergence_hist; eval_mc_prob.jl:191-198) for the
    systems the FP64 block-cooperative kernel integrates — non_local_kuramoto_ring and a WEIGHTED kuramoto_network — run
    in that kernel (one launch), the KL from the saved samples of the runs in flight; against the oracle and against
    the lane kernels (hook)."""
    m = gpu
    rng = np.random.default_rng(41)
    prng = np.random.default_rng(42)
    if system == "ring":
        N = 40
        pars = m.non_local_kuramoto_ring_parameters(N, 0.2, 0.3, lambda x: np.exp(-2 * abs(x)) / (2 * np.pi))
        rp = m.ODEProblem(m.non_local_kuramoto_ring, np.zeros(N), (0.0, 60.0), pars)
        sweep = ("phase_delay", lambda: prng.uniform(0.0, 0.6))
    else:
        N = 36
        A = cases.erdos_renyi(N, 0.3, 43) * np.random.default_rng(44).uniform(0.5, 1.5, (N, N))
        pars = m.kuramoto_network_parameters(0.5, np.random.default_rng(45).normal(0.5, 0.1, N), N, A)
        rp = m.ODEProblem(m.kuramoto_network, np.zeros(N), (0.0, 60.0), pars)
        sweep = ("K", lambda: prng.uniform(0.0, 6.0))
    ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std, m.empirical_1D_KL_divergence_hist), cyclic_setback=True)
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 70, pars, sweep, ev, 0.8)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-9, abstol=1e-9, maxiters=10**7)
    before = m.kernel_launch_count()
    sol = m.solve(prob, **kw)
    assert m.kernel_launch_count() - before == 2, "expected k_solve_batched + k_batched_kl_post"
    assert sol.feat.shape == (70, N, 3)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, **kw)
    assert np.array_equal(sol.retcode, ref["retcode"]) and np.all(sol.retcode == 0)
    cases.compare_features(sol.feat, ref, orc, *ctx, min_checked=0.5)
    assert np.isfinite(sol.feat[:, :, 2]).mean() > 0.9
    # the KL binning runs after the solve, chunk by chunk (k_batched_kl_post): a 1 MB sample scratch (8 runs per chunk)
    # must give bitwise the same measures as one chunk
    m.lib().mcbb_test_hook_set(b"batched_post_mb", 1)
    try:
        prob.ic, prob.par = ic0.copy(), par0.copy()
        before = m.kernel_launch_count()
        sol3 = m.solve(prob, **kw)
        assert m.kernel_launch_count() - before >= 10
    finally:
        m.lib().mcbb_test_hook_set(b"batched_post_mb", -2 ** 63)
    assert np.array_equal(sol3.feat, sol.feat, equal_nan=True) and np.array_equal(sol3.nsteps, sol.nsteps)
    m.lib().mcbb_test_hook_set(b"no_batched_kl", 1)
    try:
        prob.ic, prob.par = ic0.copy(), par0.copy()
        sol2 = m.solve(prob, **kw)
    finally:
        m.lib().mcbb_test_hook_set(b"no_batched_kl", -2 ** 63)
    assert np.nanmax(np.abs(sol2.feat[:, :, :2] - sol.feat[:, :, :2])) < 1e-6


def test_tensor_core_kernel_retcodes_and_fixed_dt(gpu, orc):
    """Failure and option paths of the tcgen05 integrator against the oracle: MaxIters at the exact attempt count
    (a mix of finished and truncated runs), NaN measures for truncated runs, and a user-given initial dt."""
    m = gpu
    prob, _ = cases.c5_problem(m, 40, N=40, seed=41, tspan=(0.0, 30.0))
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    full = m.solve(prob, flag_check_inf_nan=False, sort_results=False)
    limit = int(np.median(full.nsteps[:, 0]))  # about half of the runs need more attempts than this
    sol = m.solve(prob, maxiters=limit, flag_check_inf_nan=False)
    ref, ctx = _oracle_ref(m, orc, prob, ic0, par0, maxiters=limit)
    assert np.array_equal(sol.retcode, ref["retcode"])
    bad = sol.retcode == m.abi.RET_MAXITERS
    assert 0 < bad.sum() < len(bad)
    assert np.isnan(sol.feat[bad]).all() and np.isfinite(sol.feat[~bad]).all()
    assert np.array_equal(sol.nsteps[bad, 0], ref["nsteps"][bad, 0])
    # user-given initial step (no Hairer-Wanner probe): converged run must still match
    prob2, _ = cases.c5_problem(m, 24, N=40, seed=43, tspan=(0.0, 30.0))
    ic0, par0 = prob2.ic.copy(), prob2.par.copy()
    kw = dict(reltol=1e-9, abstol=1e-9, maxiters=10**7, dt=1e-3)
    sol2 = m.solve(prob2, **kw)
    ref2, ctx2 = _oracle_ref(m, orc, prob2, ic0, par0, **kw)
    assert np.array_equal(sol2.retcode, ref2["retcode"])
    cases.compare_features(sol2.feat, ref2, orc, *ctx2, min_checked=0.9)


def test_cuda_path_against_committed_golden_vectors(gpu, orc):
    """The CUDA path through the C-ABI against tests/golden/oracle_golden.json — the committed vectors, not a fresh
    oracle solve (the oracle only tells compare_features which runs are too sensitive to compare): C1- and C2-shape
    measures and step counts, the correlation measures and one saved trajectory of the power grid, and DBSCAN labels /
    seeds on the tie-rich lattice (dense and fused) with the lattice distances bit-exact."""
    import json
    import types
    from tests.golden.make_golden import build_inputs
    m = gpu
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")))
    I = build_inputs(m.abi)

    def tsave(opts, n_t, tail):
        shim = types.SimpleNamespace(discrete=opts.alg == m.abi.ALG_FUNCTIONMAP, tspan=(opts.t0, opts.t1))
        return m.tsave_array(shim, n_t, tail)

    def solve(psys, opts, pfeat, ic, par, ts, save_mode):
        prob = types.SimpleNamespace(
            par_names=[], pars=types.SimpleNamespace(pack=lambda names: psys),
            eval_ode_func=types.SimpleNamespace(pack=lambda: pfeat, failure_handling="None"),
            ic=np.asfortranarray(ic), par=np.asfortranarray(par), solver_par=None,
            p=types.SimpleNamespace(tspan=(opts.t0, opts.t1), discrete=opts.alg == m.abi.ALG_FUNCTIONMAP))
        kw = {k: getattr(opts, k) for k in ("reltol", "abstol") if getattr(opts, k) > 0}
        feat, mat, glob, ret, ns, saved = m.api._solve_b200(prob, ts, None, save_mode, kw)
        return dict(feat=feat, matrix=mat, glob=glob, retcode=ret, nsteps=ns, saved=saved)

    cases.check_solves_against_golden(G, I, orc, solve, tsave)
    X = np.asfortranarray(I["lattice"])
    gl, gw = np.array([2], dtype=np.int32), np.array([1.0])
    D = np.zeros((50, 50), order="F")
    m.check(m.lib().mcbb_distance_matrix(X.ctypes.data_as(m.abi.c_double_p), 50, 1, gl.ctypes.data_as(m.abi.c_int32_p),
                                         gw.ctypes.data_as(m.abi.c_double_p), D.ctypes.data_as(m.abi.c_double_p)))
    assert np.array_equal(D[0], np.array(G["D_lattice_row0"])) and np.array_equal(D, D.T)
    fm = m.B200FeatureMatrix(X, gl, gw, [1.0], False)
    for eps, mp in I["lattice_cases"]:
        for res in (m.dbscan(D, eps, mp), m.dbscan(fm, eps, mp)):
            assert np.array_equal(res.assignments, np.array(G["db_%s_%d" % (eps, mp)]))
            assert np.array_equal(res.seeds, np.array(G["db_seeds_%s_%d" % (eps, mp)]))


@pytest.mark.parametrize("n,gl,gw", [(1000, [6, 6, 6, 1], [1.0, 0.5, 0.5, 1.0]), (2570, [6, 6, 6, 1], [1.0, 0.5, 0.5, 1.0]),
                                     (1536, [1, 1, 1, 1], [1.0, 0.75, 0.5, 1.0]), (1001, [3, 1], [1.0, 2.0]),
                                     (4100, [40, 40, 1], [1.0, 0.3, 1.0])])
def test_tma_distance_kernel_equals_tile_kernel_bitwise(gpu, orc, n, gl, gw):
    """The persistent TMA kernel of the materialising path (csrc/dist_mat.cu) against the general tile kernel the
    on-the-fly DBSCAN uses (pair_tiles.cuh): bit-identical distances (tiles on the diagonal, partial tiles at the edge,
    a partial last feature slab, an odd n that falls back), and both within 1e-12 of the oracle."""
    m = gpu
    rng = np.random.default_rng(n)
    F = int(np.sum(gl))
    X = np.asfortranarray(rng.normal(size=(n, F)) * rng.uniform(0.1, 3.0, F))
    X[rng.integers(0, n, 5), rng.integers(0, F, 5)] = 0.0
    glen, gwt = np.asarray(gl, dtype=np.int32), np.asarray(gw, dtype=np.float64)

    def dist():
        D = np.full((n, n), np.nan, order="F")
        m.check(m.lib().mcbb_distance_matrix(X.ctypes.data_as(m.abi.c_double_p), n, len(gl), glen.ctypes.data_as(m.abi.c_int32_p),
                                             gwt.ctypes.data_as(m.abi.c_double_p), D.ctypes.data_as(m.abi.c_double_p)))
        return D
    before = m.kernel_launch_count()
    D_new = dist()  # default: the TMA kernel where the shape allows it (n even, >= 256): 64-row tiles, two CTAs per SM
    m.lib().mcbb_test_hook_set(b"dist_tma_r", 128)  # its one-CTA-per-SM variant with 128-row tiles (256 threads)
    m.lib().mcbb_test_hook_set(b"dist_tma", 0)      # ... and the general tile kernel
    try:
        D_old = dist()
        m.lib().mcbb_test_hook_set(b"dist_tma", 1)
        D_new8 = dist()
        m.lib().mcbb_test_hook_set(b"dist_tma_rpt", 4)  # 128-row tiles, 512 threads
        assert np.array_equal(dist(), D_old)
    finally:
        m.lib().mcbb_test_hook_set(b"dist_tma", -2 ** 63)
        m.lib().mcbb_test_hook_set(b"dist_tma_r", -2 ** 63)
        m.lib().mcbb_test_hook_set(b"dist_tma_rpt", -2 ** 63)
    assert np.array_equal(D_new8, D_old)
    assert m.kernel_launch_count() - before >= 2
    assert np.array_equal(D_new, D_old), np.abs(D_new - D_old).max()
    assert np.array_equal(D_new, D_new.T) and np.all(np.diag(D_new) == 0)
    sub = rng.choice(n, 300, replace=False)
    Dref = orc.distance_matrix(X[sub], glen, gwt)
    assert np.allclose(D_new[np.ix_(sub, sub)], Dref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("n,gl,gw,k,K", [(3000, [5, 6, 1], [1.0, 0.5, 1.0], 4, 15), (20000, [5, 6, 1], [1.0, 0.5, 1.0], 4, 64),
                                         (1037, [40, 40, 1], [1.0, 0.25, 1.0], 21, 5), (200, [3], [1.0], 1, 1)])
def test_fused_topk_equals_sorted_rows(gpu, orc, n, gl, gw, k, K):
    """k_dist / KNN_dist from the features (src/eval_clustering.jl:1504-1553): the fused per-row K-smallest kernel
    (csrc/topk.cuh: no row of D materialised or sorted) against the row-block + segmented-sort path, bitwise, and both
    against numpy on the oracle's matrix for a sample of rows."""
    m = gpu
    rng = np.random.default_rng(n + k)
    F = int(np.sum(gl))
    centers = rng.normal(0, 3, size=(5, F))
    X = centers[rng.integers(0, 5, n)] + rng.normal(0, 0.5, size=(n, F))
    X = np.asfortranarray(X[np.argsort(X[:, -1], kind="stable")])
    glen, gwt = np.asarray(gl, dtype=np.int32), np.asarray(gw, dtype=np.float64)

    def knn():
        kth, cum = np.full(n, np.nan), np.full(n, np.nan)
        m.check(m.lib().mcbb_knn_dist_features(X.ctypes.data_as(m.abi.c_double_p), n, len(gl), glen.ctypes.data_as(m.abi.c_int32_p),
                                               gwt.ctypes.data_as(m.abi.c_double_p), k, K, 0, kth.ctypes.data_as(m.abi.c_double_p),
                                               cum.ctypes.data_as(m.abi.c_double_p)))
        return kth, cum
    kth_f, cum_f = knn()
    m.lib().mcbb_test_hook_set(b"no_fused_topk", 1)
    try:
        kth_s, cum_s = knn()
    finally:
        m.lib().mcbb_test_hook_set(b"no_fused_topk", -2 ** 63)
    assert np.array_equal(kth_f, kth_s), np.abs(kth_f - kth_s).max()
    assert np.array_equal(cum_f, cum_s), np.abs(cum_f - cum_s).max()
    sub = np.sort(rng.choice(n, min(n, 150), replace=False))
    Dref = orc.distance_matrix(X, glen, gwt)[sub] if n <= 3000 else None
    if Dref is not None:
        srt = np.sort(Dref, axis=1)
        assert np.allclose(kth_f[sub], srt[:, k - 1], rtol=1e-12, atol=1e-12)
        assert np.allclose(cum_f[sub], srt[:, :K].sum(axis=1), rtol=1e-12, atol=1e-12)


def test_failure_handling_inf_uses_the_divergence_test(gpu):
    """failure_handling = :Inf (eval_mc_prob.jl:97-119): a DtLessThanMin / Unstable run becomes Inf only when its solution
    is empty or its last saved state exceeds 1e11; a run that merely stalls keeps its (NaN) measures and a warning is
    raised.  Logistic-free check with the Roessler network: a huge `a` makes the y components blow up (Unstable)."""
    import warnings
    m = gpu
    Nr = 3
    Lp = 2 * np.eye(Nr) - np.roll(np.eye(Nr), 1, 0) - np.roll(np.eye(Nr), -1, 0)
    pars = m.roessler_parameters(40.0 * np.ones(Nr), 0.2 * np.ones(Nr), 5.7 * np.ones(Nr), 0.05, Lp, Nr)
    rng = np.random.default_rng(3)
    rp = m.ODEProblem(m.roessler_network, np.zeros(3 * Nr), (0.0, 60.0), pars)
    ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std), failure_handling="Inf")
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-1.0, 1.0), 24, pars, ("K", lambda: rng.uniform(0.0, 0.1)), ev, 0.5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = m.solve(prob, flag_check_inf_nan=False)
    failed = (sol.retcode == m.abi.RET_UNSTABLE) | (sol.retcode == m.abi.RET_DTLESSTHANMIN)
    assert failed.any(), sol.retcode
    # exponential blow-up at rate 40: by the time the state is non-finite the last saved sample (if any) is far beyond 1e11,
    # or no sample was saved at all -> Inf, as in the reference
    assert np.isinf(sol.feat[failed]).all()
    assert np.isfinite(sol.feat[sol.retcode == 0]).all()  # (a MaxIters run is not touched by :Inf: it keeps its NaN measures)


@pytest.mark.parametrize("k_bin,zero_weight", [(1.0, False), (0.25, True)])
def test_histogram_features_on_the_device_equal_the_host_path(gpu, k_bin, zero_weight):
    """normalize(sol) + sort!(sol, prob) + the histogram-mode feature assembly of distance_matrix (setup_mc_prob.jl:520-536,
    635-660; eval_clustering.jl:187-211, 587-642) done on the device — extrema and quartiles by a device sort, Float32 ECDF
    rows by mcbb_hist_rows_dev — are bit-identical to the host mirror, including a measure whose IQR is zero (the
    nbin_default fallback) and one that is skipped for its zero weight; the host-buffer entry agrees as well."""
    import warnings

    import torch
    m = gpu
    n, n_dim = 3001, 37
    rng = np.random.default_rng(17)
    feat = np.empty((n, n_dim, 4))
    feat[:, :, 0] = rng.normal(0.3, 1.0, (n, n_dim))
    feat[:, :, 1] = np.where(rng.random((n, n_dim)) < 0.4, rng.normal(-2, 0.1, (n, n_dim)), rng.normal(3, 0.5, (n, n_dim)))
    feat[:, :, 2] = np.where(rng.random((n, n_dim)) < 0.1, rng.uniform(0, 5, (n, n_dim)), 1.25)  # IQR = 0
    feat[:, :, 3] = rng.exponential(2.0, (n, n_dim))
    par = rng.uniform(0, 1, (n, 1))
    weights = [1.0, 0.5, 0.25, 0.0 if zero_weight else 0.75, 1.0]
    order = np.argsort(par[:, 0], kind="stable")
    sol = m.DEMCBBSol(np.asfortranarray(feat[order]), None, None, np.zeros(n, dtype=np.int32), None, 100, None)
    prob, _ = cases.c1_problem(m, n_ics=4)
    prob.par = np.asfortranarray(par[order])
    prob.N_mc = n
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fm = m.distance_matrix(m.normalize(sol), prob, weights, histograms=True, k_bin=k_bin, materialize=False)
    dev = torch.device("cuda:0")
    feat_t = torch.from_numpy(np.ascontiguousarray(feat.transpose(2, 1, 0)).reshape(4 * n_dim, n)).to(dev)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        X_t, gl, gw, edges, bws = m.dist.histogram_features_device(feat_t, 4, n_dim, weights, torch.from_numpy(par.T.copy()).to(dev),
                                                                   torch.from_numpy(order).to(dev), k_bin=k_bin)
    assert np.array_equal(gl, fm.group_len) and np.array_equal(gw, fm.group_weight)
    assert (edges[3] is None) == zero_weight
    X = X_t.cpu().numpy().T
    assert X.shape == fm.X.shape
    assert np.array_equal(X, fm.X, equal_nan=True)
    # host-buffer entry on one measure
    e0 = np.ascontiguousarray(edges[0])
    slab = np.ascontiguousarray(feat[:, :, 0].T)
    rows = np.empty((len(e0) - 1, n))
    lo, hi = feat[:, :, 0].min(), feat[:, :, 0].max()
    m.check(m.lib().mcbb_hist_rows(slab.ctypes.data_as(m.abi.c_double_p), n, n_dim, order.ctypes.data_as(m.abi.c_int64_p), 1,
                                   float(lo), float(hi - lo), e0.ctypes.data_as(m.abi.c_double_p), len(e0), 1,
                                   rows.ctypes.data_as(m.abi.c_double_p)))
    assert np.array_equal(rows.T, fm.X[:, :len(e0) - 1])
    st = np.zeros(3)
    nn = np.zeros(1, dtype=np.int64)
    rk = np.array([0, slab.size // 2, slab.size - 1], dtype=np.int64)
    m.check(m.lib().mcbb_order_stats(slab.ctypes.data_as(m.abi.c_double_p), slab.size, 3, rk.ctypes.data_as(m.abi.c_int64_p),
                                     st.ctypes.data_as(m.abi.c_double_p), nn.ctypes.data_as(m.abi.c_int64_p)))
    assert np.array_equal(st, np.sort(slab.ravel())[rk]) and nn[0] == 0


def test_host_mirror_builds_large_histogram_rows_on_the_device(gpu):
    """api._hist_rows sends a measure of >= 2^18 values through mcbb_hist_rows (host buffers) and keeps numpy for toy
    sizes: both give the same Float32 rows, with and without the ECDF."""
    m = gpu
    rng = np.random.default_rng(5)
    meas = np.asfortranarray(rng.normal(0.0, 1.0, (4100, 70)))
    assert meas.size >= m.api._HIST_ROWS_ON_DEVICE
    edges, _ = m.api._compute_hist_edges(meas, 70, 0.5, 50)
    for ecdf in (True, False):
        before = m.kernel_launch_count()
        rows = m.api._hist_rows(meas, edges, ecdf)
        assert m.kernel_launch_count() - before == 1
        ref = m.api._compute_hist_weights(meas, edges, ecdf)
        assert rows.dtype == ref.dtype == np.float32 and np.array_equal(rows, ref, equal_nan=True)


def _solve_both_ways(m, prob, ts, **kw):
    """(group kernel, lane kernel) results of custom_solve_b200 on the same problem."""
    a = m.custom_solve_b200(prob, ts, **kw)
    m.lib().mcbb_test_hook_set(b"no_group", 1)
    try:
        b = m.custom_solve_b200(prob, ts, **kw)
    finally:
        m.lib().mcbb_test_hook_set(b"no_group", -2 ** 63)
    return a, b


@pytest.mark.parametrize("case", ["c1_default", "c1_cyclic_filter", "net12_weighted", "net20_sparse", "c3", "grid7_filter_kl",
                                  "c1_maxiters", "grid30_paper", "meanfield9", "c1_dt_dtmax_tight"])
def test_group_kernel_equals_lane_kernel_bitwise(gpu, orc, case):
    """csrc/group.cu (a group of 8 / 16 / 32 lanes per trajectory, one oscillator per lane) takes the sums over the state
    in the lane kernel's order, so its measures, return codes and step counts are bitwise those of the lane kernel — for
    kuramoto_network (C1 and larger, weighted / sparse couplings) and second_order_kuramoto (C3 and a small grid), with
    KL-hist, cyclic_setback, a state_filter, a MaxIters failure and the paper's 30-node grid (45 edges: two per lane)."""
    m = gpu
    kw = {}
    if case.startswith("c1"):
        prob, _ = cases.c1_problem(m, n_ics=60, seed=3, cyclic=(case == "c1_cyclic_filter"))
        if case == "c1_cyclic_filter":
            prob.eval_ode_func = m.EvalOdeRun(state_filter=[2, 5, 6], eval_funcs=(m.mean, m.std), cyclic_setback=True)
        if case == "c1_maxiters":
            kw = dict(maxiters=5)
        if case == "c1_dt_dtmax_tight":  # user dt (no initdt probe), a dtmax that clips, tight tolerances (rejections)
            kw = dict(dt=0.05, dtmax=0.4, reltol=1e-8, abstol=1e-8)
    elif case.startswith("net"):
        N = 12 if case == "net12_weighted" else 20
        rng = np.random.default_rng(N)
        A = cases.erdos_renyi(N, 0.4, N + 1)
        if case == "net12_weighted":
            A = A * rng.uniform(0.5, 1.5, (N, N))
        pars = m.kuramoto_network_parameters(0.5, rng.normal(0.5, 0.2, N), N, A)
        rp = m.ODEProblem(m.kuramoto_network, np.zeros(N), (0.0, 50.0), pars)
        ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std, m.empirical_1D_KL_divergence_hist), cyclic_setback=True)
        kr = np.random.default_rng(N + 2)
        prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 150, pars, ("K", lambda: kr.uniform(0.0, 6.0)), ev, 0.7)
    elif case == "meanfield9":  # all-to-all kuramoto (systems.jl:85-94)
        N = 9
        rng = np.random.default_rng(90)
        pars = m.kuramoto_parameters(0.5, rng.normal(0.5, 0.3, N), N)
        rp = m.ODEProblem(m.kuramoto, np.zeros(N), (0.0, 60.0), pars)
        ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std, m.empirical_1D_KL_divergence_hist), cyclic_setback=True)
        kr = np.random.default_rng(91)
        prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 140, pars, ("K", lambda: kr.uniform(0.0, 4.0)), ev, 0.7)
    elif case == "c3":
        prob, _ = cases.c3_problem(m, n_mc=300, tspan=(0.0, 120.0))
    elif case == "grid30_paper":  # the paper's size: N = 30, 3-regular, 45 edges -> two edges per lane of a 32-lane group
        prob, _ = cases.c3_problem(m, n_mc=160, N=30, seed=9, tspan=(0.0, 60.0))
        assert prob.pars.pack(prob.par_names).desc.n_edges == 45
    else:
        N = 7
        rng = np.random.default_rng(70)
        E = cases.ring_incidence(N, [(0, 3), (2, 5)])
        pars = m.second_order_kuramoto_parameters(N, 0.2, 1.0, E, rng.normal(0.0, 0.5, N))
        rp = m.ODEProblem(m.second_order_kuramoto, np.zeros(2 * N), (0.0, 80.0), pars)
        ev = m.EvalOdeRun(state_filter=[1, 3, 8, 9, 14], eval_funcs=(m.mean, m.std, m.empirical_1D_KL_divergence_hist))
        lam = np.random.default_rng(71)
        prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 200, pars, ("coupling", lambda: lam.uniform(0.05, 3.0)),
                               ev, 0.6)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    (f0, _, _, r0, n0), (f1, _, _, r1, n1) = _solve_both_ways(m, prob, ts, **kw)
    assert np.array_equal(r0, r1) and np.array_equal(n0, n1)
    assert np.array_equal(f0, f1, equal_nan=True), [float(np.nanmax(np.abs(f0[..., q] - f1[..., q]))) for q in range(f0.shape[2])]
    # KL measures are binned after the solve, chunk by chunk (k_group_kl_post): a 1 MB sample scratch = many chunks
    m.lib().mcbb_test_hook_set(b"group_post_mb", 1)
    try:
        f2, _, _, r2, n2 = m.custom_solve_b200(prob, ts, **kw)
    finally:
        m.lib().mcbb_test_hook_set(b"group_post_mb", -2 ** 63)
    assert np.array_equal(f0, f2, equal_nan=True) and np.array_equal(r0, r2) and np.array_equal(n0, n2)
    if case == "c1_maxiters":
        assert np.all(r0 == m.abi.RET_MAXITERS) and np.isnan(f0).all()
    else:
        assert np.all(r0 == 0) and np.isfinite(f0[..., :2]).all()


def test_order_stats_counts_nans_and_the_device_assembly_rejects_them(gpu):
    """mcbb_order_stats ranks the finite values in ascending order (negative zero, infinities and subnormals included)
    and reports how many NaNs the measure holds; histogram_features_device refuses a measure with NaNs (the reference's
    quantile would be NaN and its edge range meaningless) with a MCBBError, like every other argument error."""
    import torch
    m = gpu
    rng = np.random.default_rng(12)
    v = rng.normal(size=5000)
    v[:7] = [np.inf, -np.inf, 0.0, -0.0, 5e-324, -5e-324, 1e308]
    v[100:103] = np.nan
    rk = np.array([0, 1, 2500, 4995, 4996], dtype=np.int64)
    out = np.zeros(len(rk))
    nn = np.zeros(1, dtype=np.int64)
    m.check(m.lib().mcbb_order_stats(v.ctypes.data_as(m.abi.c_double_p), v.size, len(rk), rk.ctypes.data_as(m.abi.c_int64_p),
                                     out.ctypes.data_as(m.abi.c_double_p), nn.ctypes.data_as(m.abi.c_int64_p)))
    ref = np.sort(v[~np.isnan(v)])
    assert nn[0] == 3 and np.array_equal(out, ref[rk])
    dev = torch.device("cuda:0")
    feat = torch.from_numpy(v[:4000].reshape(8, 500).copy()).to(dev)
    par = torch.from_numpy(rng.uniform(size=(1, 500))).to(dev)
    with pytest.raises(m.MCBBError, match="NaN"):
        m.dist.histogram_features_device(feat, 1, 8, [1.0, 1.0], par)
    with pytest.raises(m.MCBBError, match="rank out of range"):
        bad = np.array([5000], dtype=np.int64)
        m.check(m.lib().mcbb_order_stats(v.ctypes.data_as(m.abi.c_double_p), v.size, 1, bad.ctypes.data_as(m.abi.c_int64_p),
                                         out.ctypes.data_as(m.abi.c_double_p), nn.ctypes.data_as(m.abi.c_int64_p)))
