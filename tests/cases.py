"""Seeded problem builders (shapes of the reference's test scripts / BASELINE configs) and the parity checker
used by tests/, __graft_entry__.smoke() and bench.py.  Inputs come from numpy PCG64 streams so the oracle and the
GPU path consume identical bytes."""
import numpy as np

RTOL = 1e-6          # north_star: features within 1e-6 relative (FP64)
ATOL_SCALE = 1e-10   # absolute floor = ATOL_SCALE * max(1, max |mean| of the run)  (SURVEY §7 hard part 3)


def c1_problem(m, n_ics=500, seed=1, N=6, tspan=(0.0, 40.0), ks=(1.0, 1.5, 2.0, 2.5, 3.0), tail=0.9,
               eval_funcs=None, cyclic=False):
    """C1: Kuramoto network N=6, A=ones, w~N(0.5,0.01), random ICs x K range (test/ode_prob_example.jl:36-41)."""
    rng = np.random.default_rng(seed)
    w = rng.normal(0.5, 0.01, N)
    A = np.ones((N, N))
    pars = m.kuramoto_network_parameters(0.5, w, N, A)
    rp = m.ODEProblem(m.kuramoto_network, np.zeros(N), tspan, pars)
    ev = m.EvalOdeRun(eval_funcs=eval_funcs or (m.mean, m.std, m.empirical_1D_KL_divergence_hist),
                      cyclic_setback=cyclic)
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), n_ics, pars, ("K", list(ks)), ev, tail)
    weights = [1.0] + [0.5] * (len(ev.eval_funcs) - 1) + [1.0]
    return prob, weights


def c2_problem(m, n_mc=2000, seed=3, tspan=(0.0, 1000.0), tail=0.8):
    """C2: logistic map, r~U(2.5,4), x0~U(0.1,0.99) (test/discrete_prob_example.jl, docs/src/basic_usage.md:26-78)."""
    rng = np.random.default_rng(seed)
    pars = m.logistic_parameters(2.5)
    dp = m.DiscreteProblem(m.logistic, np.array([0.5]), tspan, pars)
    r_gen = np.random.default_rng(seed + 1)
    prob = m.DEMCBBProblem(dp, lambda: rng.uniform(0.1, 0.99), n_mc, pars, ("r", lambda: r_gen.uniform(2.5, 4.0)),
                           m.eval_ode_run, tail)
    return prob, [1.0, 0.75, 0.5, 1.0]


def random_regular_graph(N, d, seed):
    """Simple d-regular graph on N nodes from the pairing model (stubs shuffled with numpy PCG64(seed); a draw with a
    self-loop or a double edge is rejected and redrawn).  Returns the edge list (a < b), sorted."""
    rng = np.random.default_rng(seed)
    assert (N * d) % 2 == 0
    while True:
        stubs = rng.permutation(np.repeat(np.arange(N), d))
        pairs = stubs.reshape(-1, 2)
        a, b = pairs.min(axis=1), pairs.max(axis=1)
        if np.any(a == b):
            continue
        edges = sorted(set(zip(a.tolist(), b.tolist())))
        if len(edges) == N * d // 2:
            return edges


def incidence(N, edges):
    """Oriented incidence matrix (N x n_edges): -1 at the tail, +1 at the head (LightGraphs incidence_matrix(g,
    oriented=true), paper/kuramoto/1-simulate.jl:25)."""
    E = np.zeros((N, len(edges)))
    for e, (a, b) in enumerate(edges):
        E[a, e] = -1.0
        E[b, e] = 1.0
    return E


def watts_strogatz(N, k, beta, seed):
    """Watts-Strogatz small world as LightGraphs.watts_strogatz(n, k, beta) builds it: ring lattice with k/2
    neighbours on each side; every lattice edge (s, s+j) is rewired with probability beta to (s, random target)
    avoiding self-loops and existing edges.  Symmetric 0/1 adjacency (p_r = 0.04, seed 7)."""
    rng = np.random.default_rng(seed)
    A = np.zeros((N, N), dtype=bool)
    for j in range(1, k // 2 + 1):
        for s in range(N):
            t = (s + j) % N
            if rng.random() < beta:
                while True:
                    t = int(rng.integers(0, N))
                    if t != s and not A[s, t]:
                        break
            A[s, t] = A[t, s] = True
    return A.astype(np.float64)


def ring_incidence(N, chords):
    """Ring plus chords: the small hand-made grids of the unit tests (the C3 case itself uses random_regular_graph)."""
    return incidence(N, [(i, (i + 1) % N) for i in range(N)] + list(chords))


def c3_problem(m, n_mc=400, seed=5, N=10, tspan=(0.0, 200.0), tail=0.9, lam_max=2.0):
    """C3: second-order Kuramoto power grid on a random 3-regular graph (seed 5), drive +-1 alternating,
    damping 0.1, coupling swept (paper/kuramoto/1-simulate.jl:22-26, 61), features on the frequencies."""
    rng = np.random.default_rng(seed + 1)
    E = incidence(N, random_regular_graph(N, 3, seed))
    drive = np.array([1.0, -1.0] * (N // 2))
    pars = m.second_order_kuramoto_parameters(N, 0.1, 1.0, E, drive)
    rp = m.ODEProblem(m.second_order_kuramoto, np.zeros(2 * N), tspan, pars)
    ev = m.EvalOdeRun(state_filter=list(range(N + 1, 2 * N + 1)), eval_funcs=(m.mean, m.std))
    lam = np.random.default_rng(seed + 2)
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), n_mc, pars,
                           ("coupling", lambda: lam.uniform(0.05, lam_max)), ev, tail)
    return prob, [1.0, 0.0, 1.0]


def c4_problem(m, n_mc, N=100, tspan=(0.0, 200.0), tail=0.7, seed=7, corr=True, p_r=0.04, P1=None):
    """C4 (paper/stuartlandau/stuart_landau_sathiyadevi-standard-config.ipynb cell 1): Stuart-Landau ring,
    lambda=1, omega=2, r1=0.01 (P1=1), r2=0.22 (P2=22), A1/A2 Watts-Strogatz p_r=0.04 (seed 7 / 7+1000), eps~U(1.8,2.1)
    (seed 8), z0~U(-1,1)+iU(-1,1) (seed 9), measures mean/std/KL of re and im + correlation_ecdf."""
    P1, P2 = P1 or int(round(0.01 * N)) or 1, int(round(0.22 * N))
    A1 = watts_strogatz(N, 2 * P1, p_r, seed)
    A2 = watts_strogatz(N, 2 * P2, p_r, seed + 1000)
    pars = m.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.925, N, P1, P2, A1, A2)
    ev = m.EvalOdeRun(eval_funcs=("mean_real", "std_real", "kl_real", "mean_imag", "std_imag", "kl_imag"),
                      matrix_eval_funcs=[m.correlation_ecdf] if corr else [])
    rp = m.ODEProblem(m.stuart_landau_sathiyadevi, np.zeros(N, dtype=complex), tspan, pars)
    erng = np.random.default_rng(seed + 1)
    zrng = np.random.default_rng(seed + 2)
    prob = m.DEMCBBProblem(rp, lambda: complex(zrng.uniform(-1, 1), zrng.uniform(-1, 1)), n_mc, pars,
                           ("eps", lambda: erng.uniform(1.8, 2.1)), ev, tail)
    weights = [0.25, 1.0, 0.0, 0.25, 1.0, 0.0] + ([0.0] if corr else []) + [1.0]
    return prob, weights


def erdos_renyi(N, p, seed):
    rng = np.random.default_rng(seed)
    U = np.triu(rng.random((N, N)) < p, 1)
    return (U | U.T).astype(np.float64)


def c5_problem(m, n_mc, N=100, seed=10, tspan=(0.0, 1000.0), tail=0.9, kmax=8.0):
    """C5: Kuramoto network N=100, Erdos-Renyi p=0.2, w~N(0.5,0.1), K~U(0,8) (docs/src/hpc.md:44-59 shape),
    measures [mean, std] with cyclic_setback."""
    A = erdos_renyi(N, 0.2, seed)
    w = np.random.default_rng(seed + 1).normal(0.5, 0.1, N)
    pars = m.kuramoto_network_parameters(0.5, w, N, A)
    rp = m.ODEProblem(m.kuramoto_network, np.zeros(N), tspan, pars)
    ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std), cyclic_setback=True)
    rng = np.random.default_rng(seed + 2)
    kr = np.random.default_rng(seed + 3)
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi, N), n_mc, pars, ("K", lambda: kr.uniform(0, kmax)),
                           ev, tail)
    return prob, [1.0, 0.5, 1.0]


def feature_tolerance(ref_feat):
    """|a-b| <= RTOL*|b| + ATOL_SCALE*scale, scale = max(1, max |value| of the run's first measure)."""
    scale = np.maximum(1.0, np.nanmax(np.abs(ref_feat[:, :, 0]), axis=1))
    return RTOL * np.abs(ref_feat) + ATOL_SCALE * scale[:, None, None]


def compare_features(feat, ref, orc, psys, opts, pfeat, ic, par, t_save, min_checked=0.6, allow_bad=0):
    """Parity of GPU features against the oracle on identical inputs.

    * mean / std (and any non-KL measure): within feature_tolerance on every NON-CHAOTIC run, where a run counts as
      chaotic / ill-conditioned when the oracle itself moves by more than half the tolerance under 1e-12 relative
      perturbations of the initial condition.
    * KL-hist: Julia's `mu-k*bw:bw:mu+k*bw` has 31 or 30 elements depending on the last bits of (mu, sig)
      ("KL range-length quirk"); the GPU value must match the oracle under one of the two lengths.
    Returns a dict of counts; raises AssertionError on mismatch."""
    abi = orc.abi
    ref_feat = ref["feat"]
    tol = feature_tolerance(ref_feat)
    is_kl = np.array([pfeat.opts.meas[m] == abi.MEAS_KL_HIST for m in range(pfeat.n_meas)])
    nk = ~is_kl
    # sensitivity probe: three seeded random relative perturbations of size 1e-12 (about 100x the rounding
    # differences between two correct FP64 implementations); a run is "non-chaotic" when the oracle stays
    # within half the tolerance under all of them
    prng = np.random.default_rng(12345)
    stable = np.ones(ref_feat.shape[0], dtype=bool)
    ic_a = np.asarray(ic, dtype=np.float64)
    for _ in range(3):
        pert_ic = ic_a * (1.0 + 1e-12 * prng.standard_normal(ic_a.shape))
        pert = orc.solve_features(psys, opts, pfeat, pert_ic, par, t_save)["feat"]
        with np.errstate(invalid="ignore"):
            stable &= (np.abs(pert[:, :, nk] - ref_feat[:, :, nk]) <= 0.5 * tol[:, :, nk]).all(axis=(1, 2))
    both_nan = np.isnan(feat) & np.isnan(ref_feat)
    with np.errstate(invalid="ignore"):
        ok = (np.abs(feat - ref_feat) <= tol) | both_nan
    bad_runs = ~ok[:, :, nk].all(axis=(1, 2)) & stable
    stats = dict(n=len(stable), checked=int(stable.sum()), bad=int(bad_runs.sum()))
    assert stats["checked"] >= min_checked * stats["n"], "too few non-chaotic runs to check: %s" % stats
    assert stats["bad"] <= allow_bad, "features differ from the oracle on non-chaotic runs: %s, worst %s" % (
        stats, np.nanmax(np.abs(feat - ref_feat)[stable][:, :, nk] / tol[stable][:, :, nk]))
    if is_kl.any():
        nb = pfeat.opts.kl_bins if pfeat.opts.kl_bins > 0 else 31
        nb += (nb % 2 == 0)
        v_a = orc.solve_features(psys, opts, pfeat, ic, par, t_save, kl_force_len=nb)["feat"]
        v_b = orc.solve_features(psys, opts, pfeat, ic, par, t_save, kl_force_len=nb - 1)["feat"]
        kl_tol = lambda v: RTOL * np.abs(v) + 1e-9
        with np.errstate(invalid="ignore"):
            okk = (np.abs(feat - v_a) <= kl_tol(v_a)) | (np.abs(feat - v_b) <= kl_tol(v_b)) | both_nan
        # the sig < 1e-4 => 0 switch (eval_mc_prob.jl:314-316) is a discontinuity: skip dims sitting on it
        std_idx = [m for m in range(pfeat.n_meas) if pfeat.opts.meas[m] == abi.MEAS_STD]
        on_switch = np.zeros(feat.shape[:2], dtype=bool)
        for m in std_idx:
            on_switch |= np.abs(ref_feat[:, :, m] - 1e-4) < 1e-9
        bad_kl = (~okk[:, :, is_kl]).any(axis=2) & ~on_switch & stable[:, None]
        stats["kl_bad"] = int(bad_kl.sum())
        assert stats["kl_bad"] <= allow_bad, "KL-hist differs from both oracle variants: %s" % stats
    return stats


def check_solves_against_golden(G, I, orc, solve, tsave):
    """The solve-related golden vectors (tests/golden/oracle_golden.json) against `solve(psys, opts, pfeat, ic, par,
    t_save, save_mode) -> dict(feat, matrix, nsteps, retcode, saved)` — the CUDA path through the C-ABI, or the host
    build of the lane source.  `tsave(opts, N_t, tail)` is the product's save grid.  The oracle is only used by
    compare_features to tell which runs are too sensitive to compare (and for the KL range-length variants)."""
    abi = orc.abi
    # C1 shape: measures within the feature tolerance, identical step sequences
    ts = tsave(I["c1_opts"], 400, 0.9)
    ctx = (I["c1_sys"], I["c1_opts"], I["feat_default"], I["c1_ic"], I["c1_par"], ts)
    got = solve(*ctx, 0)
    assert np.all(got["retcode"] == 0) and np.array_equal(got["nsteps"], np.array(G["c1_nsteps"]))
    stats = compare_features(got["feat"], {"feat": np.array(G["c1_feat"])}, orc, *ctx)
    assert stats["checked"] >= 12
    # C2 shape: the map is evaluated with the reference's literal expression
    tl = tsave(I["c2_opts"], 400, 0.8)
    ctx = (I["c2_sys"], I["c2_opts"], I["feat_default"], I["c2_ic"], I["c2_par"], tl)
    got = solve(*ctx, 0)
    compare_features(got["feat"], {"feat": np.array(G["c2_feat"])}, orc, *ctx, min_checked=0.3)
    # correlation_hist / correlation_ecdf of the grid frequencies: a |cor| within rounding of a bin edge may move
    tg = tsave(I["c3_opts"], 400, 0.5)
    for name in ("hist", "ecdf"):
        got = solve(I["c3_sys"], I["c3_opts"], I["c3_feat"][name], I["c3_ic"], I["c3_par"], tg, 0)
        want = np.array(G["c3_corr_" + name])
        same = np.all(np.abs(got["matrix"] - want) < 1e-12, axis=1)
        assert same.sum() >= 7, (name, same)
        assert np.abs(got["matrix"] - want).max() <= (4 if name == "hist" else 4 / 30 + 1e-12)
    # the saved samples of run 0 (get_trajectory)
    got = solve(I["c3_sys"], I["c3_opts"], I["c3_feat"]["hist"], I["c3_ic"][:1], I["c3_par"][:1], tg, abi.SAVE_ALL)
    want = np.array(G["c3_trajectory_run0"])
    assert np.abs(got["saved"][0][:, ::40] - want).max() <= 1e-7 * max(1.0, np.abs(want).max())
