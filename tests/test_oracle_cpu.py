"""CPU tests (-m "not gpu") of the ORACLE: what pins the restatement, given that the reference's own tests hold no
golden vectors and Julia cannot run here.  Analytic known answers, scipy cross-checks, order
conditions of the Tsit5 tableau, sklearn DBSCAN on tie-free inputs, and the committed golden vectors."""
import json
import os

import numpy as np
import pytest

from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))


def _kur(abi, N=6, seed=1, K=2.0):
    rng = np.random.default_rng(seed)
    A = np.ones((N, N))
    w = rng.normal(0.5, 0.01, N)
    return abi.PackedSystem(abi.SYS_KURAMOTO_NETWORK, N, N, [K], [0], mat_a=A, vec_a=w), A, w


def test_tsit5_order_conditions():
    """Row sums = c, b (7th row) satisfies orders 1-5, btilde sums to 0, interpolant reproduces b at theta = 1."""
    a = np.zeros((7, 7))
    vals = [0.161, -0.008480655492356989, 0.335480655492357, 2.8971530571054935, -6.359448489975075,
            4.3622954328695815, 5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525,
            5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383,
            0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]
    k = 0
    for i in range(1, 7):
        for j in range(i):
            a[i, j] = vals[k]
            k += 1
    c = np.array([0, 0.161, 0.327, 0.9, 0.9800255409045097, 1, 1])
    assert np.allclose(a.sum(1), c, atol=1e-14)
    b = a[6]
    assert abs(b.sum() - 1) < 1e-14
    for p in range(1, 5):
        assert abs(b @ c ** p - 1 / (p + 1)) < 1e-13
    assert abs(b @ (a @ c) - 1 / 6) < 1e-13 and abs(b @ (a @ c ** 2) - 1 / 12) < 1e-13
    bt = np.array([-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                   0.5823571654525552, -0.45808210592918697, 0.015151515151515152])
    assert abs(bt.sum()) < 1e-15


def test_tsit5_converges_to_scipy(orc):
    abi = orc.abi
    ps, A, w = _kur(abi)
    from scipy.integrate import solve_ivp
    o = abi.solver_opts(abi.ALG_TSIT5, (0.0, 40.0), abstol=1e-12, reltol=1e-12, maxiters=10 ** 7)
    ts = orc.tsave_array(o, 400, 0.9)
    assert len(ts) == 400 and ts[0] == 36.0 and abs(ts[-1] - 39.99) < 1e-12
    u0 = np.random.default_rng(2).uniform(-np.pi, np.pi, 6)
    saved, ret, ns, nsv = orc.trajectory(ps, o, u0, [2.0], ts)
    s = solve_ivp(lambda t, u: 2.0 / 6 * (np.sin(u[None, :] - u[:, None]) * A).sum(1) + w, (0, 40), u0, method="DOP853",
                  rtol=1e-13, atol=1e-13, t_eval=ts)
    assert ret == 0 and nsv == 400
    assert np.abs(saved - s.y).max() < 1e-9
    # RK4 fixed step with Hermite saveat
    o4 = abi.solver_opts(abi.ALG_RK4, (0.0, 40.0), dt=0.01)
    saved4, ret4, ns4, _ = orc.trajectory(ps, o4, u0, [2.0], ts)
    assert ret4 == 0 and ns4[0] == 4000 and np.abs(saved4 - s.y).max() < 1e-8


def test_interpolant_is_fourth_order_accurate(orc):
    """Default tolerances: saved samples come from the Tsit5 interpolant, error stays at the integration tolerance."""
    abi = orc.abi
    ps, A, w = _kur(abi)
    u0 = np.random.default_rng(3).uniform(-np.pi, np.pi, 6)
    ts = orc.tsave_array(abi.solver_opts(abi.ALG_TSIT5, (0.0, 40.0)), 400, 0.9)
    coarse, _, ns, _ = orc.trajectory(ps, abi.solver_opts(abi.ALG_TSIT5, (0.0, 40.0)), u0, [2.0], ts)
    fine, _, _, _ = orc.trajectory(ps, abi.solver_opts(abi.ALG_TSIT5, (0.0, 40.0), abstol=1e-12, reltol=1e-12,
                                                         maxiters=10 ** 7), u0, [2.0], ts)
    assert ns[0] < 60 and np.abs(coarse - fine).max() < 5e-2


def test_uncoupled_kuramoto_closed_form(orc):
    """K = 0: u_i(t) = u_i0 + w_i t, so mean / std over the tail grid have closed forms (incl. cyclic setback)."""
    abi = orc.abi
    N = 4
    w = np.array([0.3, -0.2, 0.5, 0.1])
    ps = abi.PackedSystem(abi.SYS_KURAMOTO_NETWORK, N, N, [0.0], [0], mat_a=np.ones((N, N)), vec_a=w)
    o = abi.solver_opts(abi.ALG_TSIT5, (0.0, 100.0))
    ts = orc.tsave_array(o, 400, 0.9)
    u0 = np.array([[0.1, -2.0, 3.0, 1.0]])
    for cyc in (False, True):
        pf = abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD], cyclic_setback=cyc)
        r = orc.solve_features(ps, o, pf, u0, np.array([[0.0]]), ts)
        x = u0[0][:, None] + w[:, None] * ts[None, 1:]
        if cyc:
            first = np.mod(x[:, 0], 2 * np.pi)
            first = np.where(first > np.pi, first - 2 * np.pi, first)
            x = x - (x[:, :1] - first[:, None])
        assert np.allclose(r["feat"][0, :, 0], x.mean(1), rtol=1e-9, atol=1e-9)
        assert np.allclose(r["feat"][0, :, 1], x.std(1, ddof=1), rtol=1e-9, atol=1e-9)


def test_logistic_known_answers(orc):
    abi = orc.abi
    pl = abi.PackedSystem(abi.SYS_LOGISTIC, 1, 1, [2.8], [0])
    ol = abi.solver_opts(abi.ALG_FUNCTIONMAP, (0.0, 1000.0))
    tl = orc.tsave_array(ol, 400, 0.8)
    assert len(tl) == 200 and tl[0] == 800 and tl[-1] == 999  # N_t ignored for DiscreteProblem
    pf = abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD, abi.MEAS_KL_HIST])
    r = orc.solve_features(pl, ol, pf, np.array([[0.3], [0.5]]), np.array([[2.8], [3.2]]), tl)
    # r = 2.8: fixed point 1 - 1/r; std below sig_tol => KL exactly 0 (eval_mc_prob.jl:314-316)
    assert abs(r["feat"][0, 0, 0] - (1 - 1 / 2.8)) < 1e-12 and r["feat"][0, 0, 1] < 1e-10 and r["feat"][0, 0, 2] == 0.0
    # r = 3.2: period-2 orbit; 199 samples alternate between the two closed-form branches
    xs = ((3.2 + 1) + np.array([1, -1]) * np.sqrt((3.2 - 3) * (3.2 + 1))) / (2 * 3.2)
    orbit = np.array([xs[i % 2] for i in range(199)])
    m_a, m_b = orbit.mean(), np.array([xs[(i + 1) % 2] for i in range(199)]).mean()
    assert min(abs(r["feat"][1, 0, 0] - m_a), abs(r["feat"][1, 0, 0] - m_b)) < 1e-9
    assert abs(r["feat"][1, 0, 1] - orbit.std(ddof=1)) < 1e-9
    assert np.all(r["nsteps"][:, 0] == 1000)


def test_kl_hist_against_numpy_restatement(orc):
    """empirical_1D_KL_divergence_hist against an independent numpy restatement, including the missing-mu last edge."""
    rng = np.random.default_rng(0)
    for mu, sig in [(0.0, 1.0), (5.0, 0.5), (-3.0, 2.0), (100.0, 0.01)]:
        u = rng.normal(mu, sig, 399)
        m, s = u.mean(), u.std(ddof=1)
        for n_c in (31, 30):
            k = 15.0
            bw = 3 * s / k
            centers = (m - k * bw) + bw * np.arange(n_c)
            edges = np.concatenate([centers - bw / 2, [k * bw + bw / 2]])
            idx = np.array([_ssl(edges, x) for x in u])
            keep = (idx >= 1) & (idx <= n_c)
            h = np.bincount(idx[keep] - 1, minlength=n_c).astype(float)
            p = h / h.sum()
            q = np.exp(-((centers - m) / s) ** 2 / 2)
            q /= q.sum()
            ref = np.sum(np.where(p > 0, p * np.log(np.where(p > 0, p, 1) / q), 0.0))
            got = orc.kl_hist(u, m, s, force_len=n_c)
            assert abs(got - ref) < 1e-10 * max(1, abs(ref)), (mu, sig, n_c)
    assert orc.kl_hist(np.ones(10), 1.0, 1e-5) == 0.0


def test_correlation_hist_and_ecdf_against_numpy(orc):
    """correlation_hist / correlation_ecdf (eval_mc_prob.jl:462-488) against numpy corrcoef + histogram: the weights
    of |cor| over 0:1/30:1 (left-closed bins, so the unit diagonal falls outside), and ecdf_hist of them.  The first
    saved sample is dropped before measuring (eval_mc_prob.jl:95)."""
    abi = orc.abi
    rng = np.random.default_rng(3)
    N, n_t, n_mc = 7, 120, 12
    psys, _, _ = _kur(abi, N=N)
    mix = rng.normal(size=(n_mc, N, N))
    saved = np.einsum("rij,rjt->rit", mix, rng.normal(size=(n_mc, N, n_t)))
    out = {}
    for kind in (abi.MAT_CORR_HIST, abi.MAT_CORR_ECDF):
        pfeat = abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD], matrix_meas=kind, corr_nbins=30)
        out[kind] = orc.featurize_saved(psys, pfeat, saved)["matrix"]
    for r in range(n_mc):
        c = np.abs(np.corrcoef(saved[r][:, 1:]))
        # numpy's last bin is closed on the right, Julia's is not: the unit diagonal is dropped by hand
        w = np.histogram(c[~np.eye(N, dtype=bool)], bins=np.arange(31) / 30.0)[0]
        assert w.sum() == N * (N - 1)
        assert np.array_equal(out[abi.MAT_CORR_HIST][r], w.astype(float))
        assert np.allclose(out[abi.MAT_CORR_ECDF][r], np.cumsum(w) / w.sum(), rtol=0, atol=1e-15)


def test_complex_state_correlation_against_numpy(mcbb, orc):
    """|cor| of complex series (Stuart-Landau states, interleaved re/im): Statistics.cor conjugates the second factor,
    clampcor leaves complex values alone, abs gives the magnitude — the same as numpy's complex corrcoef."""
    abi = orc.abi
    rng = np.random.default_rng(4)
    N, n_t, n_mc = 6, 90, 10
    ring = np.roll(np.eye(N), 1, 0) + np.roll(np.eye(N), -1, 0)
    psys = mcbb.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.9, N, 1, 1, ring, ring).pack(["eps"])
    mix = rng.normal(size=(n_mc, N, N)) + 1j * rng.normal(size=(n_mc, N, N))
    z = np.einsum("rij,rjt->rit", mix, rng.normal(size=(n_mc, N, n_t)) + 1j * rng.normal(size=(n_mc, N, n_t)))
    saved = np.stack([z.real, z.imag], axis=2).reshape(n_mc, 2 * N, n_t)  # dimension index: re_0, im_0, re_1, ...
    filt = [2, 3, 5, 6]
    for state_filter in (None, filt):
        sel = np.arange(N) if state_filter is None else np.array(filt) - 1
        out = {}
        for kind in (abi.MAT_CORR_HIST, abi.MAT_CORR_ECDF):
            pfeat = abi.PackedFeat([abi.MEAS_MEAN, abi.MEAS_STD], meas_comp=[0, 1], state_filter=state_filter,
                                   matrix_meas=kind, corr_nbins=30)
            out[kind] = orc.featurize_saved(psys, pfeat, saved)["matrix"]
        for r in range(n_mc):
            c = np.abs(np.corrcoef(z[r][sel][:, 1:]))
            w = np.histogram(c[~np.eye(len(sel), dtype=bool)], bins=np.arange(31) / 30.0)[0]
            assert w.sum() == len(sel) * (len(sel) - 1)
            assert np.array_equal(out[abi.MAT_CORR_HIST][r], w.astype(float))
            assert np.allclose(out[abi.MAT_CORR_ECDF][r], np.cumsum(w) / w.sum(), rtol=0, atol=1e-15)


def _ssl(v, x):  # Base.searchsortedlast on a Vector, binary search as in sort.jl
    lo, hi = 0, len(v) + 1
    while lo < hi - 1:
        m = (lo + hi) >> 1
        if x < v[m - 1]:
            hi = m
        else:
            lo = m
    return lo


def test_complex_state_error_norm_is_rotation_invariant(orc):
    """DiffEqBase's default norm works per ELEMENT of the state array: for a Complex{Float64} state the error scale is
    abstol + |z_i| reltol with the complex modulus and the RMS runs over N complex elements.  Uncoupled Stuart-Landau
    oscillators (eps = 0) are invariant under z -> z e^{i phi}; with the element-wise complex norm so is every
    accept / reject decision and every step size — a per-real-component norm would break that."""
    abi = orc.abi
    N = 5
    ring = np.roll(np.eye(N), 1, 0) + np.roll(np.eye(N), -1, 0)
    psys = abi.PackedSystem(abi.SYS_STUART_LANDAU, N, 2 * N, [0.0, 1.0, 2.0, 1.0, 1.0], [], mat_a=ring, mat_b=ring)
    opts = abi.solver_opts(abi.ALG_TSIT5, (0.0, 30.0))
    rng = np.random.default_rng(3)
    z = rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)
    ts = orc.tsave_array(opts, 50, 0.5)
    res = []
    for phi in (0.0, 0.7, 2.1):
        w = z * np.exp(1j * phi)
        u0 = np.stack([w.real, w.imag], axis=1).ravel()
        saved, ret, ns, nsv = orc.trajectory(psys, opts, u0, np.zeros(0), ts)
        res.append((saved[0::2] + 1j * saved[1::2], ns.copy()))
        assert ret == 0 and nsv == len(ts)
    for (zz, ns), phi in zip(res[1:], (0.7, 2.1)):
        assert np.array_equal(ns, res[0][1]), "step sequence depends on the global phase: per-component norm?"
        assert np.allclose(zz, res[0][0] * np.exp(1j * phi), rtol=0, atol=1e-12)


def test_distance_matrix_and_dbscan_vs_sklearn(orc):
    from sklearn.cluster import DBSCAN
    rng = np.random.default_rng(5)
    X = np.concatenate([rng.normal(c, 0.25, size=(50, 3)) for c in [(0, 0, 0), (3, 0, 0), (0, 3, 1)]] +
                       [rng.uniform(-2, 5, size=(30, 3))])
    gl, gw = np.array([2, 1], dtype=np.int32), np.array([1.0, 0.5])
    D = orc.distance_matrix(X, gl, gw)
    ref = np.abs(X[:, None, :2] - X[None, :, :2]).sum(-1) + 0.5 * np.abs(X[:, None, 2] - X[None, :, 2])
    assert np.allclose(D, ref, rtol=1e-13, atol=1e-14) and np.array_equal(D, D.T)
    for eps, mp in [(0.45, 5), (0.6, 3)]:
        assert not np.any(np.isclose(D, eps, rtol=0, atol=1e-12))  # tie-free: strict '<' == sklearn's '<='
        res = orc.dbscan(D, eps, mp)
        sk = DBSCAN(eps=eps, min_samples=mp, metric="precomputed").fit(D)
        core = np.zeros(len(X), bool)
        core[sk.core_sample_indices_] = True
        a = res["assignments"]
        # same core partition (border points may legitimately differ between visit orders; cores may not)
        lab_map = {}
        for i in np.nonzero(core)[0]:
            lab_map.setdefault(sk.labels_[i], a[i])
            assert lab_map[sk.labels_[i]] == a[i] and a[i] > 0
        assert len(set(lab_map.values())) == len(lab_map) == len(res["seeds"])
        assert np.array_equal(a == 0, sk.labels_ == -1)
        assert np.array_equal(res["counts"], np.bincount(a, minlength=len(res["seeds"]) + 1)[1:])


def test_dbscan_parallel_restatement_property(orc):
    """The order-independent restatement the GPU implements (components of the core graph numbered by smallest
    core index; border -> smallest adjacent cluster) equals the sequential transcription, on tie-rich random D."""
    rng = np.random.default_rng(11)
    for _ in range(60):
        n = int(rng.integers(2, 80))
        U = np.round(rng.random((n, n)) * 10) / 10
        D = np.triu(U, 1)
        D = D + D.T
        eps = float(rng.choice([0.1, 0.2, 0.3, 0.5]))
        mp = int(rng.integers(1, 6))
        res = orc.dbscan(D, eps, mp)
        nb = D < eps
        core = nb.sum(0) >= mp
        parent = np.arange(n)

        def find(x):
            while parent[x] != x:
                x = parent[x]
            return x
        for i in range(n):
            for j in range(i + 1, n):
                if core[i] and core[j] and nb[i, j]:
                    a, b = find(i), find(j)
                    if a != b:
                        parent[max(a, b)] = min(a, b)
        roots = sorted({find(i) for i in range(n) if core[i]})
        rank = {r: k + 1 for k, r in enumerate(roots)}
        lab = np.zeros(n, dtype=np.int64)
        for i in range(n):
            if core[i]:
                lab[i] = rank[find(i)]
            else:
                adj = [rank[find(j)] for j in range(n) if core[j] and nb[i, j]]
                lab[i] = min(adj) if adj else 0
        assert np.array_equal(lab, res["assignments"])
        assert np.array_equal(np.array(roots) + 1, res["seeds"])


def test_cluster_membership_windows(orc):
    par = np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0])
    a = np.array([1, 1, 1, 0, 2, 2, 2, 2, 0, 1, 1])
    mesh, out = orc.cluster_membership(par, a, 2, 0.3, 0.1, normalize=False)
    # Julia: 0.0:0.1:0.7 has 8 elements (rational lift of the float range), windows_mins[end] == 0.7
    assert len(mesh) == 8 and np.allclose(mesh, np.arange(8) / 10.0, atol=1e-15)
    # window (0, 0.3): members 0.1, 0.2 only (both bounds strict, eval_clustering.jl:1222)
    assert np.array_equal(out[0], [0, 2, 0])
    assert np.array_equal(out[3], [0, 0, 2])  # (0.3, 0.6): 0.4, 0.5
    assert np.array_equal(out[7], [1, 1, 0])  # (0.7, 1.0): 0.8, 0.9


def test_golden_vectors(orc):
    """Committed golden vectors (tests/golden/, generated by tests/golden/make_golden.py from this oracle):
    guards the oracle itself against silent drift."""
    path = os.path.join(HERE, "golden", "oracle_golden.json")
    G = json.load(open(path))
    from tests.golden.make_golden import compute_all
    now = compute_all(orc)
    for k, v in G.items():
        assert np.allclose(np.array(now[k], dtype=float), np.array(v, dtype=float), rtol=1e-12, atol=1e-13,
                           equal_nan=True), k
