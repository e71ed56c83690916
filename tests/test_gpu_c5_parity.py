"""GPU parity of the BENCHMARKED path (-m gpu): C5 (Kuramoto network N = 100, Erdos-Renyi p = 0.2, K ~ U(0, 8), tspan
(0, 1000), tail 0.9, N_t = 400, Tsit5 at the DiffEq default tolerances abstol 1e-6 / reltol 1e-3, measures [mean, std]
with cyclic_setback) through the tcgen05 kernel `k_solve_batched_umma` — the configuration, tolerances and time span
bench.py times — against the CPU oracle on identical seeded inputs.

At reltol 1e-3 most large-K runs are stability-limited: the step-size controller keeps EEst near 1 and every
accept / reject decision is a discontinuity, so ANY two correct FP64 implementations drift apart by far more than
1e-6 on them (the oracle itself does, under a 1e-12 relative perturbation of the initial condition).  Instead of
excluding those runs, EVERY run is bounded by the oracle's own perturbation spread:

    |gpu - oracle|  <=  C * max_k |oracle(ic (1 + 1e-12 xi_k)) - oracle(ic)|  +  1e-6 * max(1, |oracle|)

and the end-to-end acceptance of north_star is asserted on the same ensemble: solve -> distance_matrix -> DBSCAN ->
cluster_membership on both sides, adjusted Rand index >= 0.99 and membership curves equal within a stated tolerance
(src/eval_clustering.jl:1206-1265)."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

N_RUNS = 1536      # the oracle needs ~12 s per solve of this ensemble on the GPU box's host cores
N_PERT = 4         # perturbed oracle solves that measure every run's own sensitivity
C_SPREAD = 10.0    # allowed multiple of that spread


@pytest.fixture(scope="module")
def c5_bench_ensemble(gpu, orc):
    m = gpu
    prob, weights = cases.c5_problem(m, N_RUNS)  # tspan (0, 1000), tail 0.9: the bench workload
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    before = m.kernel_launch_count()
    sol = m.solve(prob)  # default tolerances, sorted by parameter like MCBB.solve
    assert m.kernel_launch_count() - before == 1, "expected one launch of the batched tensor-core integrator"
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    o = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    idx = np.argsort(par0[:, 0], kind="stable")
    ic, par = ic0[idx], par0[idx]
    ref = orc.solve_features(psys, o, pfeat, ic, par, ts)
    prng = np.random.default_rng(2025)
    perts = [orc.solve_features(psys, o, pfeat, ic * (1.0 + 1e-12 * prng.standard_normal(ic.shape)), par, ts)["feat"]
             for _ in range(N_PERT)]
    return dict(m=m, prob=prob, weights=weights, sol=sol, ref=ref, perts=perts)


def test_c5_bench_tolerances_every_run_bounded_by_oracle_spread(c5_bench_ensemble):
    e = c5_bench_ensemble
    sol, ref, perts = e["sol"], e["ref"], e["perts"]
    assert np.array_equal(sol.retcode, ref["retcode"]) and np.all(sol.retcode == 0)
    rf = ref["feat"]
    err = np.abs(sol.feat - rf).max(axis=(1, 2))                                  # per run
    spread = np.max([np.abs(p - rf).max(axis=(1, 2)) for p in perts], axis=0)      # per run
    scale = np.maximum(1.0, np.abs(rf).max(axis=(1, 2)))
    bound = C_SPREAD * spread + 1e-6 * scale
    ok = err <= bound
    tight = spread <= 1e-6 * scale  # runs the oracle calls insensitive: these must match to 1e-6 outright
    stats = dict(n=len(err), within_bound=int(ok.sum()), frac_within_bound=float(ok.mean()),
                 insensitive=int(tight.sum()), insensitive_ok=int((ok & tight).sum()),
                 median_err=float(np.median(err)), median_spread=float(np.median(spread)),
                 worst_ratio=float((err / bound).max()))
    print("c5 bench-tolerance parity (T=1000, default tolerances):", stats)
    assert (ok | ~tight).all(), "an insensitive run differs from the oracle by more than 1e-6: %s" % stats
    # a run can exceed C x its measured spread only if none of the N_PERT probes happened to flip the same
    # accept / reject decision the GPU's rounding flipped; with 4 probes that has to stay rare
    assert stats["frac_within_bound"] >= 0.97, stats
    # the error distribution of the GPU must look like the oracle's own perturbation distribution
    assert np.median(err) <= 3.0 * np.median(spread) + 1e-6, stats
    assert np.quantile(err, 0.9) <= 3.0 * np.quantile(spread, 0.9) + 1e-6, stats


def test_c5_bench_tolerances_step_counts(c5_bench_ensemble):
    """roofline.achieved in bench.py is built on the kernel's own step-attempt count: it has to be the oracle's."""
    e = c5_bench_ensemble
    gs, os_ = e["sol"].nsteps[:, 0].astype(float), e["ref"]["nsteps"][:, 0].astype(float)
    ratio = gs.sum() / os_.sum()
    rel = np.abs(gs - os_) / os_
    print("c5 step attempts: gpu mean %.2f, oracle mean %.2f, ratio %.5f, per-run |rel| median %.4f max %.4f, "
          "identical on %.1f %% of the runs" % (gs.mean(), os_.mean(), ratio, np.median(rel), rel.max(),
                                                100.0 * (gs == os_).mean()))
    assert abs(ratio - 1.0) < 0.02
    assert np.median(rel) < 0.02 and rel.max() < 0.25
    rj_g, rj_o = e["sol"].nsteps[:, 1].sum(), e["ref"]["nsteps"][:, 1].sum()
    assert abs(rj_g / max(rj_o, 1) - 1.0) < 0.05, (rj_g, rj_o)


def test_c5_bench_tolerances_cluster_membership_ari(c5_bench_ensemble, orc):
    """north_star acceptance on the benchmarked arithmetic: DBSCAN labels from the GPU's features against DBSCAN
    labels from the oracle's features (same eps from the reference's KNN heuristic on the oracle's matrix), adjusted
    Rand index >= 0.99, and cluster_membership curves (src/eval_clustering.jl:1206-1265) within 0.01."""
    from sklearn.metrics import adjusted_rand_score
    e = c5_bench_ensemble
    m, prob, weights, sol, ref = e["m"], e["prob"], e["weights"], e["sol"], e["ref"]
    D = m.distance_matrix(sol, prob, weights)
    rsol = m.DEMCBBSol(ref["feat"], None, None, ref["retcode"], ref["nsteps"], 400, None)
    Xr, gl, gw = m.feature_groups(rsol, prob, weights)
    Dref = orc.distance_matrix(Xr, gl, gw)
    n = Dref.shape[0]
    K = int(round(0.005 * n))
    eps = float(np.median(np.sort(Dref, axis=0)[:K].sum(axis=0)))  # median(KNN_dist_relative(D)), :1526-1553
    minpts = 20
    res = m.dbscan(D, eps, minpts)
    rdb = orc.dbscan(Dref, eps, minpts)
    ari = adjusted_rand_score(rdb["assignments"], res.assignments)
    # the library's DBSCAN on the ORACLE's matrix must be bit-exact (integer work)
    res_on_ref = m.dbscan(np.asfortranarray(Dref), eps, minpts)
    assert np.array_equal(res_on_ref.assignments, rdb["assignments"])
    assert np.array_equal(res_on_ref.seeds, rdb["seeds"]) and np.array_equal(res_on_ref.counts, rdb["counts"])
    cm = m.cluster_membership(prob, res, 0.8, 0.2)
    mesh, rcm = orc.cluster_membership(prob.par[:, 0], rdb["assignments"], len(rdb["seeds"]), 0.8, 0.2)
    k = min(cm.data.shape[1], rcm.shape[1])
    diff = float(np.nanmax(np.abs(cm.data[:, :k] - rcm[:, :k]))) if cm.data.shape == rcm.shape else float("inf")
    print("c5 membership: eps %.4f, clusters gpu %d / oracle %d, noise gpu %d / oracle %d, ARI %.5f, "
          "membership max |diff| %.4f, max |D - Dref| %.3e" % (
              eps, len(res.seeds), len(rdb["seeds"]), int((res.assignments == 0).sum()),
              int((rdb["assignments"] == 0).sum()), ari, diff, float(np.abs(D.data - Dref).max())))
    assert ari >= 0.99
    assert np.allclose(cm.par, mesh) and cm.data.shape == rcm.shape and diff <= 0.01


def test_c5_default_tolerances_structured_ensemble_ari(gpu, orc):
    """The bench ensemble above collapses to one cluster under the reference's eps heuristic (the mean of a phase that
    has drifted to ~500 rad carries +-30 rad of integration noise at reltol 1e-3, in the oracle itself), so its ARI of 1
    says little.  A shorter horizon (tspan (0, 100), tail 0.5) of the SAME network at the SAME default tolerances keeps
    cluster structure (a synchronised cluster, a small second one, ~25 % noise).  Even there the reference path is not
    reproducible to ARI 0.99 against ITSELF at every eps (oracle vs oracle with a 1e-12 perturbed IC: 0.93 at 1x eps),
    so the bar is north_star's 0.99 wherever the oracle reproduces itself that well, and the oracle's own
    reproducibility (minus 0.03) otherwise — both printed."""
    from sklearn.metrics import adjusted_rand_score
    m = gpu
    n = 640
    prob, weights = cases.c5_problem(m, n, tspan=(0.0, 100.0), tail=0.5, seed=30)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    sol = m.solve(prob)
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    o = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    idx = np.argsort(par0[:, 0], kind="stable")
    ic, par = ic0[idx], par0[idx]
    ref = orc.solve_features(psys, o, pfeat, ic, par, ts)
    prng = np.random.default_rng(77)
    perts = [orc.solve_features(psys, o, pfeat, ic * (1.0 + 1e-12 * prng.standard_normal(ic.shape)), par, ts)
             for _ in range(2)]

    def dist_of(feat):
        rsol = m.DEMCBBSol(feat, None, None, ref["retcode"], ref["nsteps"], 400, None)
        X, gl, gw = m.feature_groups(rsol, prob, weights)
        return orc.distance_matrix(X, gl, gw)
    Dref = dist_of(ref["feat"])
    Dp = [dist_of(p["feat"]) for p in perts]
    Dg = m.distance_matrix(sol, prob, weights).data
    K = int(round(0.005 * n))
    eps0 = float(np.median(np.sort(Dref, axis=0)[:K].sum(axis=0)))  # median(KNN_dist_relative(D))
    for fac in (1.0, 2.0, 4.0):
        eps = fac * eps0
        lab_ref = orc.dbscan(Dref, eps, 5)["assignments"]
        lab_gpu = m.dbscan(np.asfortranarray(Dg), eps, 5).assignments
        self_ari = min(adjusted_rand_score(lab_ref, orc.dbscan(D, eps, 5)["assignments"]) for D in Dp)
        ari = adjusted_rand_score(lab_ref, lab_gpu)
        print("structured C5 ensemble: eps %.2f (%.0fx heuristic): %d clusters, %d noise; ARI gpu-vs-oracle %.4f, "
              "oracle-vs-perturbed-oracle %.4f" % (eps, fac, lab_ref.max(), int((lab_ref == 0).sum()), ari, self_ari))
        assert ari >= min(0.99, self_ari - 0.03), (fac, ari, self_ari)


@pytest.mark.parametrize("N", [28, 40, 100])
def test_batched_sizes_with_a_matrix_measure_fall_through(gpu, orc, N):
    """kuramoto_network sizes that the batched kernels take for (mean, std) must NOT take a problem that also asks for a
    matrix measure and no global one (no batched kernel computes it): correlation_ecdf has to come back from the lane
    kernels, equal to the oracle's, never as unwritten scratch."""
    m = gpu
    prob, _ = cases.c5_problem(m, 24, N=N, seed=40 + N, tspan=(0.0, 30.0))
    prob.eval_ode_func = m.EvalOdeRun(eval_funcs=(m.mean, m.std), matrix_eval_funcs=[m.correlation_ecdf],
                                      cyclic_setback=True)
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    kw = dict(reltol=1e-9, abstol=1e-9, maxiters=10 ** 7)
    sol = m.solve(prob, **kw)
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    o = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan, **kw)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    idx = np.argsort(par0[:, 0], kind="stable")
    ref = orc.solve_features(psys, o, pfeat, ic0[idx], par0[idx], ts)
    assert sol.matrix is not None and sol.matrix.shape == (24, 30)
    assert np.isfinite(sol.matrix).all() and np.allclose(sol.matrix[:, -1], 1.0)
    npairs = N * (N - 1)
    d = np.abs(sol.matrix - ref["matrix"])
    # a |cor| within rounding of a bin edge may move one pair (two ordered pairs) to the neighbouring bin
    assert d.max() <= 4.0 / npairs + 1e-15 and (d.max(axis=1) < 1e-15).mean() >= 0.7, d.max()
    cases.compare_features(sol.feat, ref, orc, psys, o, pfeat, ic0[idx], par0[idx], ts, min_checked=0.8)


@pytest.mark.parametrize("N", [100, 97])
def test_light_warp_variant_is_bitwise_equal(gpu, N):
    """csrc/batched_umma_lw.cuh (experiment switch `umma_lw`; measured slower): the warp of the short row block
    carries one element per lane and the row blocks rotate over the tensor-memory quadrants.  Every reduction adds in the
    order of the plain mapping, so measures and step counts must be bitwise equal to the default kernel."""
    m = gpu
    prob, _ = cases.c5_problem(m, 700, N=N, seed=70 + N, tspan=(0.0, 120.0))
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    f0, _, _, r0, n0 = m.custom_solve_b200(prob, ts)
    out = []
    for mode in (1, 2):
        m.lib().mcbb_test_hook_set(b"umma_lw", mode)
        try:
            out.append(m.custom_solve_b200(prob, ts))
        finally:
            m.lib().mcbb_test_hook_set(b"umma_lw", -2 ** 63)
    for f1, _, _, r1, n1 in out:
        assert np.array_equal(f0, f1, equal_nan=True) and np.array_equal(r0, r1) and np.array_equal(n0, n1)
    assert np.all(r0 == 0) and np.isfinite(f0).all()


def test_five_ctas_of_four_trajectories_is_bitwise_equal(gpu):
    """Experiment switch umma_jb = 4, umma_per_sm = 5 (two tensor-memory allocations per CTA, 64 + 32 columns, so that five
    CTAs fit an SM; measured slower than six trajectories x four CTAs): results do not depend on how the
    trajectories are batched, so measures and step counts are bitwise equal to the default configuration."""
    m = gpu
    prob, _ = cases.c5_problem(m, 900, N=100, seed=77, tspan=(0.0, 120.0))
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    f0, _, _, r0, n0 = m.custom_solve_b200(prob, ts)
    m.lib().mcbb_test_hook_set(b"umma_jb", 4)
    m.lib().mcbb_test_hook_set(b"umma_per_sm", 5)
    try:
        f1, _, _, r1, n1 = m.custom_solve_b200(prob, ts)
    finally:
        m.lib().mcbb_test_hook_set(b"umma_jb", -2 ** 63)
        m.lib().mcbb_test_hook_set(b"umma_per_sm", -2 ** 63)
    assert np.array_equal(f0, f1, equal_nan=True) and np.array_equal(r0, r1) and np.array_equal(n0, n1)
    assert np.all(r0 == 0)


def test_kl_measures_are_binned_chunk_by_chunk(gpu):
    """The KL-hist measure of the tensor-core integrator is evaluated after the solve by k_umma_kl_post from the saved
    samples of a chunk of runs.  A 20 MB scratch (65 runs per chunk at N = 100, 400 save times) and a chunk size that does
    not divide the ensemble give bitwise the results of one chunk."""
    m = gpu
    prob, _ = cases.c5_problem(m, 300, N=100, seed=81, tspan=(0.0, 100.0))
    prob.eval_ode_func = m.EvalOdeRun(eval_funcs=(m.mean, m.std, m.empirical_1D_KL_divergence_hist), cyclic_setback=True)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    before = m.kernel_launch_count()
    f0, _, _, r0, n0 = m.custom_solve_b200(prob, ts)
    assert m.kernel_launch_count() - before == 2  # k_solve_batched_umma + k_umma_kl_post
    for mb, launches in ((20, 10), (61, 4)):  # 65 runs per chunk -> 5 chunks; 199 per chunk -> 2
        m.lib().mcbb_test_hook_set(b"umma_post_mb", mb)
        try:
            before = m.kernel_launch_count()
            f1, _, _, r1, n1 = m.custom_solve_b200(prob, ts)
            assert m.kernel_launch_count() - before == launches
        finally:
            m.lib().mcbb_test_hook_set(b"umma_post_mb", -2 ** 63)
        assert np.array_equal(f0, f1, equal_nan=True) and np.array_equal(r0, r1) and np.array_equal(n0, n1)
    assert np.all(r0 == 0) and np.isfinite(f0[..., 2]).mean() > 0.9
