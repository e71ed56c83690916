"""CPU test (-m "not gpu") of the PRODUCT's per-lane integrator source: tests/host_emul/lane_emul.cu compiles
csrc/lane_core.cuh (the code every `k_solve_lane<SYS>` instantiates) for the host and runs it one lane at a time, so
the state machine, the fused statistics, the correlation measures and the sample export are compared with the oracle
in the CPU-only container as well.  The harness is test infrastructure: the product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_emul", "lane_emul.cu")
OUT = os.path.join(HERE, "host_emul", "_build", "liblane_emul.so")


@pytest.fixture(scope="module")
def emul(mcbb):
    deps = [SRC, mcbb.LIB_PATH] + [os.path.join(ROOT, "mcbb.jl_b200", "csrc", f)
                                   for f in ("lane_core.cuh", "systems.cuh", "common.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        libdir = os.path.dirname(mcbb.LIB_PATH)
        cmd = ["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler",
               "-fPIC", "-I" + os.path.join(ROOT, "include"), "-o", OUT, SRC, "-L" + libdir,
               "-l:" + os.path.basename(mcbb.LIB_PATH), "-Xlinker", "-rpath", "-Xlinker", libdir]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    L = C.CDLL(OUT)
    abi = mcbb.abi
    dp = abi.c_double_p
    L.emul_solve_features.argtypes = [C.POINTER(abi.SystemDesc), C.POINTER(abi.SolverOpts), C.POINTER(abi.FeatOpts),
                                      C.c_int64, dp, dp, dp, C.c_int32, dp, dp, abi.c_int32_p, abi.c_int64_p, dp,
                                      C.c_int32, dp, C.c_int32, dp]

    def solve(psys, opts, pfeat, ic, par, t_save, save_mode=0, solver_var=None):
        ic = np.asfortranarray(ic, dtype=np.float64)
        par = np.asfortranarray(np.asarray(par, dtype=np.float64).reshape(ic.shape[0], -1))
        t_save = np.ascontiguousarray(t_save, dtype=np.float64)
        n_mc, nf = ic.shape[0], pfeat.n_filter(psys)
        feat = np.full((n_mc, nf, pfeat.n_meas), np.nan, order="F")
        mat = np.full((n_mc, pfeat.corr_nbins), np.nan, order="F")
        glob = np.full((n_mc, max(1, pfeat.n_global)), np.nan, order="F")
        ret = np.zeros(n_mc, dtype=np.int32)
        ns = np.zeros((n_mc, 2), dtype=np.int64, order="F")
        saved = None
        if save_mode == abi.SAVE_ALL:
            saved = np.full((n_mc, len(t_save), psys.n_dim), np.nan)
        elif save_mode == abi.SAVE_LAST:
            saved = np.full((n_mc, psys.n_dim), np.nan, order="F")
        field, vals = 0, None
        if solver_var is not None:
            field, vals = abi.SOLVER_FIELDS[solver_var[0]], np.ascontiguousarray(solver_var[1], dtype=np.float64)
        rc = L.emul_solve_features(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n_mc, ic.ctypes.data_as(dp),
                                   par.ctypes.data_as(dp), t_save.ctypes.data_as(dp), len(t_save),
                                   feat.ctypes.data_as(dp), glob.ctypes.data_as(dp), ret.ctypes.data_as(abi.c_int32_p),
                                   ns.ctypes.data_as(abi.c_int64_p), mat.ctypes.data_as(dp), int(save_mode),
                                   saved.ctypes.data_as(dp) if saved is not None else dp(), field,
                                   vals.ctypes.data_as(dp) if vals is not None else dp())
        assert rc == 0
        if save_mode == abi.SAVE_ALL:
            saved = np.transpose(saved, (0, 2, 1))
        return dict(feat=feat, matrix=mat, glob=glob, retcode=ret, nsteps=ns, saved=saved)
    return solve


def _setup(m, prob, **kw):
    psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
    a = m.abi.ALG_FUNCTIONMAP if prob.p.discrete else m.abi.ALG_TSIT5
    o = m.abi.solver_opts(a, prob.p.tspan, **kw)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    ic = prob.ic
    if np.iscomplexobj(ic):
        ic = np.stack([ic.real, ic.imag], axis=2).reshape(ic.shape[0], -1)
    return psys, o, pfeat, ic, prob.par, ts


def test_lane_source_default_measures_match_oracle(mcbb, orc, emul):
    prob, _ = cases.c1_problem(mcbb, n_ics=12, seed=41)
    ctx = _setup(mcbb, prob)
    got, ref = emul(*ctx), orc.solve_features(*ctx)
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["nsteps"], ref["nsteps"])
    cases.compare_features(got["feat"], ref, orc, *ctx)
    prob, _ = cases.c2_problem(mcbb, n_mc=64)
    ctx = _setup(mcbb, prob)
    got, ref = emul(*ctx), orc.solve_features(*ctx)
    cases.compare_features(got["feat"], ref, orc, *ctx)


def test_lane_source_sample_export_matches_oracle_trajectories(mcbb, orc, emul):
    abi = mcbb.abi
    prob, _ = cases.c1_problem(mcbb, n_ics=6, seed=42, eval_funcs=(mcbb.mean, mcbb.std))
    ctx = _setup(mcbb, prob)
    psys, o, pfeat, ic, par, ts = ctx
    plain = emul(*ctx)
    all_s = emul(*ctx, save_mode=abi.SAVE_ALL)
    last = emul(*ctx, save_mode=abi.SAVE_LAST)
    assert np.array_equal(all_s["feat"], plain["feat"]) and np.array_equal(last["feat"], plain["feat"])
    assert np.array_equal(all_s["saved"][:, :, -1], last["saved"])
    for i in range(prob.N_mc):
        ref, ret, _, nsv = orc.trajectory(psys, o, ic[i], par[i], ts)
        assert ret == 0 and nsv == len(ts)
        assert np.allclose(all_s["saved"][i], ref, rtol=1e-10, atol=1e-10)
    # the measures are the statistics of these samples, first one dropped
    assert np.allclose(plain["feat"][:, :, 0], all_s["saved"][:, :, 1:].mean(axis=2), rtol=1e-12, atol=1e-12)
    assert np.allclose(plain["feat"][:, :, 1], all_s["saved"][:, :, 1:].std(axis=2, ddof=1), rtol=1e-9, atol=1e-12)
    # a run that hits maxiters keeps NaN for the samples it never reached
    ctx_f = _setup(mcbb, prob, maxiters=8)
    cut = emul(*ctx_f, save_mode=abi.SAVE_ALL)
    assert np.all(cut["retcode"] == abi.RET_MAXITERS) and np.isnan(cut["saved"][:, :, -1]).all()


def test_lane_source_correlation_measures_real_and_complex(mcbb, orc, emul):
    m = mcbb
    for f in (m.correlation_hist, m.correlation_ecdf):
        prob, _ = cases.c3_problem(m, n_mc=10, seed=5, N=10, tspan=(0.0, 60.0), tail=0.5)
        prob.eval_ode_func = m.EvalOdeRun(state_filter=list(range(11, 21)), eval_funcs=(m.mean, m.std),
                                          matrix_eval_funcs=[f])
        ctx = _setup(m, prob, reltol=1e-9, abstol=1e-9)
        got, ref = emul(*ctx), orc.solve_features(*ctx)
        assert np.allclose(got["matrix"], ref["matrix"], rtol=0, atol=1e-15)
        if f == m.correlation_hist:
            assert np.all(got["matrix"].sum(axis=1) == 90) and (got["matrix"][:, :-1].sum(axis=1) > 0).all()
    Ns = 8
    A1 = np.roll(np.eye(Ns), 1, 0) + np.roll(np.eye(Ns), -1, 0)
    A2 = A1 + np.roll(np.eye(Ns), 2, 0) + np.roll(np.eye(Ns), -2, 0)
    for f, filt in ((m.correlation_hist, None), (m.correlation_ecdf, [1, 3, 4, 8])):
        rng, prng = np.random.default_rng(31), np.random.default_rng(32)
        pars = m.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.9, Ns, 1, 2, A1, A2)
        ev = m.EvalOdeRun(state_filter=filt, eval_funcs=("mean_real", "std_real", "mean_imag", "std_imag"),
                          matrix_eval_funcs=[f])
        rp = m.ODEProblem(m.stuart_landau_sathiyadevi, np.zeros(Ns, dtype=complex), (0.0, 10.0), pars)
        prob = m.DEMCBBProblem(rp, lambda: complex(rng.uniform(-1, 1), rng.uniform(-1, 1)), 12, pars,
                               ("eps", lambda: prng.uniform(1.8, 2.1)), ev, 0.0)
        ctx = _setup(m, prob, reltol=1e-9, abstol=1e-9)
        got, ref = emul(*ctx), orc.solve_features(*ctx)
        cases.compare_features(got["feat"], ref, orc, *ctx, min_checked=0.8)
        assert np.allclose(got["matrix"], ref["matrix"], rtol=0, atol=1e-15)
        if filt is None:
            assert np.all(got["matrix"].sum(axis=1) == 56) and (got["matrix"][:, :-1].sum(axis=1) > 0).all()


def test_lane_source_per_run_solver_options(mcbb, orc, emul):
    """SolverParameterVar (src/solver_parametervar.jl; test/solver_parametervar.jl varies :reltol per run): every run
    equals the oracle solved alone with that run's option."""
    m = mcbb
    rng = np.random.default_rng(7)
    w = rng.normal(0.5, 0.005, 6)
    pars = m.kuramoto_network_parameters(0.5, w, 6, np.ones((6, 6)))
    rp = m.ODEProblem(m.kuramoto_network, np.zeros(6), (0.0, 40.0), pars)
    vgen = np.random.default_rng(8)
    for name, gen, kw in (("reltol", lambda: vgen.uniform(1e-6, 1e-3), {}),
                          ("abstol", lambda: 10.0 ** vgen.uniform(-9, -4), {}),
                          ("dtmax", lambda: vgen.uniform(0.05, 0.5), {}),
                          ("dt", [0.01, 0.02, 0.05], dict(alg=m.RK4()))):
        prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 4, pars, m.SolverParameterVar(name, gen),
                               m.EvalOdeRun(eval_funcs=(m.mean, m.std)), 0.9)
        assert prob.par_names == [] and prob.solver_par == name and prob.par.shape == (prob.N_mc, 1)
        psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
        assert psys.desc.n_par == 0
        alg = m.abi.ALG_RK4 if "alg" in kw else m.abi.ALG_TSIT5
        base = dict(dt=0.03) if name == "dt" else {}  # the launch-wide dt is overridden per run
        o = m.abi.solver_opts(alg, prob.p.tspan, **base)
        ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
        got = emul(psys, o, pfeat, prob.ic, prob.par, ts, solver_var=(name, prob.par[:, 0]))
        steps = set()
        for i in range(prob.N_mc):
            oi = m.abi.solver_opts(alg, prob.p.tspan, **{name: float(prob.par[i, 0])})
            ref = orc.solve_features(psys, oi, pfeat, prob.ic[i:i + 1], prob.par[i:i + 1], ts)
            assert got["retcode"][i] == ref["retcode"][0] == 0
            assert np.array_equal(got["nsteps"][i], ref["nsteps"][0]), (name, i)
            assert np.allclose(got["feat"][i], ref["feat"][0], rtol=1e-6, atol=1e-7), (name, i)
            steps.add(int(ref["nsteps"][0, 0]))
        assert len(steps) > 1  # the option really changed the integration
    with pytest.raises(m.MCBBError):
        m.SolverParameterVar("maxiters", [1, 2])


def test_lane_source_two_swept_parameters(mcbb, orc, emul):
    """MultiDimParameterVar: two system parameters (coupling, damping of the power grid) overridden per run."""
    m = mcbb
    N = 6
    rng = np.random.default_rng(9)
    E = cases.ring_incidence(N, [(0, 3)])
    pars = m.second_order_kuramoto_parameters(N, 0.1, 1.0, E, np.array([1.0, -1.0] * (N // 2)))
    rp = m.ODEProblem(m.second_order_kuramoto, np.zeros(2 * N), (0.0, 60.0), pars)
    ev = m.EvalOdeRun(state_filter=list(range(N + 1, 2 * N + 1)), eval_funcs=(m.mean, m.std))
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-np.pi, np.pi), 16, pars,
                           [("coupling", lambda: rng.uniform(0.5, 4.0)), ("damping", lambda: rng.uniform(0.05, 0.5))],
                           ev, 0.5)
    ctx = _setup(m, prob, reltol=1e-8, abstol=1e-8)
    assert ctx[0].desc.n_par == 2
    got, ref = emul(*ctx), orc.solve_features(*ctx)
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["nsteps"], ref["nsteps"])
    cases.compare_features(got["feat"], ref, orc, *ctx, min_checked=0.5)
    # the second parameter matters: the same runs with the damping column held fixed differ
    prob.par[:, 1] = 0.1
    ctx2 = _setup(m, prob, reltol=1e-8, abstol=1e-8)
    assert not np.allclose(emul(*ctx2)["feat"], got["feat"])


def test_lane_source_every_system_features_and_samples(mcbb, orc, emul):
    """All seven systems of src/systems.jl (+ the Stuart-Landau notebook system) through the lane source on the host:
    retcodes, step counts, measures and the exported samples against the oracle."""
    m = mcbb
    rng = np.random.default_rng(21)
    N = 6
    checks = []
    pars = m.kuramoto_parameters(1.0, rng.normal(0.5, 0.05, N), N)
    checks.append((pars, m.kuramoto, N, "K", (0.5, 3.0), (0.0, 30.0), m.EvalOdeRun(eval_funcs=(m.mean, m.std)), False))
    pars = m.non_local_kuramoto_ring_parameters(N, 0.3, 0.2, lambda x: np.exp(-abs(x)) / (2 * np.pi))
    checks.append((pars, m.non_local_kuramoto_ring, N, "phase_delay", (0.0, 1.4), (0.0, 30.0),
                   m.EvalOdeRun(eval_funcs=(m.mean, m.std), cyclic_setback=True), False))
    Nr = 3
    Lp = 2 * np.eye(Nr) - np.roll(np.eye(Nr), 1, 0) - np.roll(np.eye(Nr), -1, 0)
    pars = m.roessler_parameters(0.2 * np.ones(Nr), 0.2 * np.ones(Nr), 5.7 * np.ones(Nr), 0.05, Lp, Nr)
    checks.append((pars, m.roessler_network, 3 * Nr, "K", (0.0, 0.1), (0.0, 15.0),
                   m.EvalOdeRun(eval_funcs=(m.mean, m.std)), False))
    A1 = np.roll(np.eye(N), 1, 0) + np.roll(np.eye(N), -1, 0)
    A2 = A1 + np.roll(np.eye(N), 2, 0) + np.roll(np.eye(N), -2, 0)
    pars = m.stuart_landau_sathiyadevi_pars(1.0, 2.0, 1.9, N, 1, 2, A1, A2)
    ev = m.EvalOdeRun(eval_funcs=("mean_real", "std_real", "kl_real", "mean_imag", "std_imag", "kl_imag"))
    checks.append((pars, m.stuart_landau_sathiyadevi, N, "eps", (1.8, 2.1), (0.0, 20.0), ev, True))
    E = cases.ring_incidence(N, [(0, 3)])
    pars = m.second_order_kuramoto_parameters(N, 0.1, 1.0, E, np.array([1.0, -1.0] * (N // 2)))
    checks.append((pars, m.second_order_kuramoto, 2 * N, "coupling", (0.5, 4.0), (0.0, 30.0),
                   m.EvalOdeRun(state_filter=list(range(N + 1, 2 * N + 1)), eval_funcs=(m.mean, m.std)), False))
    w = rng.normal(0.5, 0.01, N)
    pars = m.kuramoto_network_parameters(0.5, w, N, (rng.random((N, N)) < 0.6).astype(float))
    checks.append((pars, m.kuramoto_network, N, "K", (0.5, 3.0), (0.0, 30.0), m.eval_ode_run, False))
    for pars, f, ndim, pname, prange, tspan, ev, cplx in checks:
        prng = np.random.default_rng(5)
        if cplx:
            gen, u0 = (lambda: complex(rng.uniform(-1, 1), rng.uniform(-1, 1))), np.zeros(ndim, dtype=complex)
        else:
            gen, u0 = (lambda: rng.uniform(-1.0, 1.0)), np.zeros(ndim)
        rp = m.ODEProblem(f, u0, tspan, pars)
        prob = m.DEMCBBProblem(rp, gen, 10, pars, (pname, lambda: prng.uniform(*prange)), ev, 0.7)
        ctx = _setup(m, prob)
        psys, o, pfeat, ic, par, ts = ctx
        got, ref = emul(*ctx, save_mode=m.abi.SAVE_ALL), orc.solve_features(*ctx)
        assert np.array_equal(got["retcode"], ref["retcode"]), f.__name__
        assert np.array_equal(got["nsteps"], ref["nsteps"]), f.__name__
        cases.compare_features(got["feat"], ref, orc, *ctx, min_checked=0.3)
        for i in (0, 9):
            want = orc.trajectory(psys, o, ic[i], par[i], ts)[0]
            scale = max(1.0, np.abs(want).max())
            assert np.abs(got["saved"][i] - want).max() <= 1e-7 * scale, (f.__name__, i)
    # the map: FunctionMap stepping, every iterate after the transient is a sample
    prob, _ = cases.c2_problem(m, n_mc=16, tspan=(0.0, 300.0))
    ctx = _setup(m, prob)
    psys, o, pfeat, ic, par, ts = ctx
    got = emul(*ctx, save_mode=m.abi.SAVE_ALL)
    for i in range(16):
        x, r = ic[i, 0], par[i, 0]
        orbit = []
        for t in range(300):
            if t >= int(ts[0]):
                orbit.append(x)
            x = r * x * (1 - x)
        assert np.array_equal(got["saved"][i, 0], np.array(orbit[:len(ts)]))


def test_lane_source_against_committed_golden_vectors(mcbb, orc, emul):
    """The host build of the lane source against tests/golden/oracle_golden.json (no fresh oracle solve)."""
    import json
    from tests.golden.make_golden import build_inputs
    G = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))
    I = build_inputs(mcbb.abi)

    def tsave(opts, n_t, tail):
        class P:
            discrete = opts.alg == mcbb.abi.ALG_FUNCTIONMAP
            tspan = (opts.t0, opts.t1)
        ts = mcbb.tsave_array(P, n_t, tail)
        assert np.array_equal(ts, orc.tsave_array(opts, n_t, tail))
        return ts
    cases.check_solves_against_golden(G, I, orc, lambda *a: emul(*a[:6], save_mode=a[6]), tsave)


def test_lane_source_option_combinations(mcbb, orc, emul):
    """Seeded random combinations of what eval_ode_run / solve can be asked for — measures (with and without KL),
    state_filter, cyclic_setback, global sum, a matrix measure, Tsit5 at several tolerances or fixed-step RK4,
    maxiters failures — on three systems: retcodes and step counts identical to the oracle, measures within the feature
    tolerance, failed runs NaN."""
    m, abi = mcbb, mcbb.abi
    rng = np.random.default_rng(77)
    N = 5
    E = cases.ring_incidence(N, [(0, 2)])
    systems = [
        lambda: (m.kuramoto_network_parameters(0.5, rng.normal(0.5, 0.02, N), N, (rng.random((N, N)) < 0.7) * 1.0),
                 m.kuramoto_network, N, "K", (0.5, 3.0)),
        lambda: (m.second_order_kuramoto_parameters(N, 0.1, 1.0, E, rng.normal(0, 1, N)), m.second_order_kuramoto,
                 2 * N, "coupling", (0.5, 3.0)),
        lambda: (m.kuramoto_parameters(1.0, rng.normal(0.5, 0.05, N), N), m.kuramoto, N, "K", (0.2, 2.0)),
    ]
    n_fail_cases = 0
    for trial in range(30):
        pars, f, ndim, pname, prange = systems[trial % 3]()
        funcs = [(m.mean, m.std), (m.mean, m.std, m.empirical_1D_KL_divergence_hist), (m.mean,)][rng.integers(3)]
        filt = None if rng.random() < 0.5 else sorted(rng.choice(np.arange(1, ndim + 1), size=3, replace=False).tolist())
        mat = [None, [m.correlation_ecdf], [m.correlation_hist]][rng.integers(3)]
        ev = m.EvalOdeRun(state_filter=filt, eval_funcs=funcs, matrix_eval_funcs=mat,
                          global_eval_funcs=["sum"] if rng.random() < 0.5 else None,
                          cyclic_setback=bool(rng.random() < 0.5))
        rp = m.ODEProblem(f, np.zeros(ndim), (0.0, float(rng.choice([20.0, 40.0]))), pars)
        prng = np.random.default_rng(int(rng.integers(1 << 30)))
        prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-2.0, 2.0), 6, pars, (pname, lambda: prng.uniform(*prange)), ev,
                               float(rng.choice([0.5, 0.8, 0.9])))
        mode = rng.integers(4)
        if mode == 0:
            kw, alg = {}, None
        elif mode == 1:
            kw, alg = dict(reltol=1e-7, abstol=1e-8), None
        elif mode == 2:
            kw, alg = dict(dt=0.05), m.abi.ALG_RK4
        else:
            kw, alg = dict(maxiters=int(rng.integers(3, 40))), None
        psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
        o = abi.solver_opts(alg if alg is not None else abi.ALG_TSIT5, prob.p.tspan, **kw)
        ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
        ctx = (psys, o, pfeat, prob.ic, prob.par, ts)
        got, ref = emul(*ctx), orc.solve_features(*ctx)
        tag = (trial, f.__name__, mode)
        assert np.array_equal(got["retcode"], ref["retcode"]), tag
        assert np.array_equal(got["nsteps"], ref["nsteps"]), tag
        failed = ref["retcode"] != 0
        n_fail_cases += int(failed.any())
        assert np.isnan(got["feat"][failed]).all(), tag
        if (~failed).sum() >= 2:
            ok = ~failed
            sub_ctx = (psys, o, pfeat, prob.ic[ok], prob.par[ok], ts)
            cases.compare_features(got["feat"][ok], {"feat": ref["feat"][ok]}, orc, *sub_ctx, min_checked=0.3)
            if ev.global_eval_funcs:
                assert np.allclose(got["glob"][ok, :1], ref["glob"][ok], rtol=1e-6, atol=1e-8), tag
            if mat is not None:
                # locked oscillators have |cor| within rounding of 1, the open end of the last bin: whether a pair
                # is counted there or dropped hangs on the last bit, so the last bin (and rows whose whole mass sits
                # in it, whose ECDF is 0/0 or 0,...,0,1) are left out of the comparison
                g, r = got["matrix"][ok], ref["matrix"][ok]
                if mat == [m.correlation_hist]:
                    assert np.array_equal(g[:, :-1], r[:, :-1]), tag
                    assert np.all(g[:, -1] <= ndim * (ndim - 1)) and np.all(g >= 0), tag
                else:
                    spread = np.nan_to_num(r[:, -2], nan=0.0) > 0
                    unaffected = spread & (np.abs(g[:, -2] - r[:, -2]) < 1e-12)
                    assert np.all(np.abs(g[unaffected] - r[unaffected]) < 1e-12), tag
                    assert unaffected.sum() >= 0.5 * spread.sum(), (tag, spread, unaffected)
    assert n_fail_cases >= 1  # the maxiters mode really produced failed runs somewhere


def test_lane_source_failure_retcodes(mcbb, orc, emul):
    """Integrator failures are data (retcode per run): MaxIters, DtLessThanMin (a large dtmin against a tight tolerance)
    and Unstable (a diverging Roessler network) — same codes, same step counts and NaN measures as the oracle; the
    escaping logistic orbit (r > 4) stays `Success` with non-finite measures."""
    m, abi = mcbb, mcbb.abi
    prob, _ = cases.c1_problem(m, n_ics=4, seed=3, eval_funcs=(m.mean, m.std))
    seen = set()
    for kw in (dict(maxiters=7), dict(dtmin=0.5, reltol=1e-10, abstol=1e-10), dict(dtmax=0.01, maxiters=500)):
        ctx = _setup(m, prob, **kw)
        got, ref = emul(*ctx), orc.solve_features(*ctx)
        assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["nsteps"], ref["nsteps"]), kw
        assert np.all(ref["retcode"] != abi.RET_SUCCESS) and np.isnan(got["feat"]).all(), kw
        seen |= set(ref["retcode"].tolist())
    assert {abi.RET_MAXITERS, abi.RET_DTLESSTHANMIN} <= seen
    Nr = 3
    Lp = 2 * np.eye(Nr) - np.roll(np.eye(Nr), 1, 0) - np.roll(np.eye(Nr), -1, 0)
    pars = m.roessler_parameters(0.2 * np.ones(Nr), 0.2 * np.ones(Nr), 5.7 * np.ones(Nr), 0.05, Lp, Nr)
    rp = m.ODEProblem(m.roessler_network, np.zeros(3 * Nr), (0.0, 60.0), pars)
    rng = np.random.default_rng(1)
    probr = m.DEMCBBProblem(rp, lambda: rng.uniform(-1, 1), 2, pars, ("K", [0.05, 60.0, 200.0]),
                            m.EvalOdeRun(eval_funcs=(m.mean, m.std)), 0.5)
    ctx = _setup(m, probr)
    got, ref = emul(*ctx), orc.solve_features(*ctx)
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["nsteps"], ref["nsteps"])
    assert (ref["retcode"] == abi.RET_SUCCESS).any() and (ref["retcode"] != abi.RET_SUCCESS).any(), ref["retcode"]
    bad = ref["retcode"] != abi.RET_SUCCESS
    assert np.isnan(got["feat"][bad]).all() and np.isfinite(got["feat"][~bad]).all()
    parsl = m.logistic_parameters(4.5)
    dp = m.DiscreteProblem(m.logistic, np.array([0.5]), (0.0, 500.0), parsl)
    probl = m.DEMCBBProblem(dp, lambda: 0.3, 2, parsl, ("r", [3.9, 4.5]), m.eval_ode_run, 0.8)
    ctx = _setup(m, probl)
    got, ref = emul(*ctx), orc.solve_features(*ctx)
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.all(ref["retcode"] == abi.RET_SUCCESS)
    assert np.array_equal(np.isfinite(got["feat"]), np.isfinite(ref["feat"]))
    assert np.isfinite(got["feat"][probl.par[:, 0] < 4]).all() and not np.isfinite(got["feat"][probl.par[:, 0] > 4]).all()


@pytest.fixture
def cpu_engine(monkeypatch, mcbb, emul):
    """Replaces the two functions of the host mirror that call the CUDA library by the host build of the lane source,
    so that the mirror's own logic (solve: sorting, :Repeat, the response analysis; get_trajectory) runs in the CPU
    suite.  Test-only: the product has no such path."""
    api, abi = mcbb.api, mcbb.abi

    def fake_solve(prob, t_save, alg, save_mode, kwargs):
        psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
        o = api._solver_opts(prob, alg, kwargs)
        ic = prob.ic
        cplx = np.iscomplexobj(ic)
        if cplx:
            ic = np.stack([ic.real, ic.imag], axis=2).reshape(ic.shape[0], -1)
        sv = None
        if getattr(prob, "solver_par", None) is not None:
            sv = (prob.solver_par, prob.par[:, 0])
        r = emul(psys, o, pfeat, ic, prob.par, t_save, save_mode=save_mode, solver_var=sv)
        saved = r["saved"]
        if saved is not None and cplx:
            saved = saved[:, 0::2] + 1j * saved[:, 1::2]
        mat = r["matrix"] if pfeat.matrix_meas else None
        glob = r["glob"] if pfeat.n_global else None
        return r["feat"], mat, glob, r["retcode"], r["nsteps"], saved

    monkeypatch.setattr(api, "_solve_b200", fake_solve)
    monkeypatch.setattr(api, "custom_solve_b200", lambda prob, ts, alg=None, **kw: fake_solve(prob, ts, alg, 0, kw)[:5])
    return mcbb


def test_host_mirror_response_analysis_and_get_trajectory_on_cpu(cpu_engine, orc):
    """solve() with the response analysis (eval_mc_prob.jl:202-257), get_trajectory and failure_handling = :Repeat —
    the host mirror's logic end to end, the lane source standing in for the GPU — against the oracle composed from its
    single-trajectory, featurize and distance functions."""
    m = cpu_engine
    eps, bounds, w_inner = 0.25, (1.0, 3.0), [1.0, 0.5, 1.0]
    prob, _ = cases.c1_problem(m, n_ics=5, seed=29, eval_funcs=(m.mean, m.std))
    prob.eval_ode_func = m.EvalOdeRun(eval_funcs=(m.mean, m.std),
                                      continuation=m.ContinuationResponse(eps, bounds, w_inner, N_t=200))
    sol = m.solve(prob)
    assert (sol.N_meas_dim, sol.N_meas_global) == (2, 2) and np.all(np.diff(prob.par[:, 0]) >= 0)
    psys, pfeat = prob.pars.pack(prob.par_names), m.EvalOdeRun(eval_funcs=(m.mean, m.std)).pack()
    o = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan)
    ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
    n = prob.N_mc
    end = np.stack([orc.trajectory(psys, o, prob.ic[i], prob.par[i], ts)[0][:, -1] for i in range(n)])
    p0 = prob.par[:, 0]
    par3 = np.stack([np.maximum(p0 - eps, bounds[0]), p0, np.minimum(p0 + eps, bounds[1])], axis=1)
    t0, t1 = prob.p.tspan
    span2 = (t0, t0 + 0.15 * (t1 - t0))
    o2 = m.abi.solver_opts(m.abi.ALG_TSIT5, span2)
    ts2 = m.tsave_array(m.ODEProblem(m.kuramoto_network, np.zeros(6), span2, prob.pars), 200, 0.1)
    ref3 = orc.solve_features(psys, o2, pfeat, np.repeat(end, 3, axis=0), par3.reshape(-1, 1), ts2)["feat"]
    X = np.concatenate([ref3[:, :, 0], ref3[:, :, 1], par3.reshape(-1, 1)], axis=1)
    want = np.array([[orc.distance_matrix(X[3 * i:3 * i + 3], [6, 6, 1], w_inner)[a, a + 1] for a in (0, 1)]
                     for i in range(n)])
    assert np.allclose(sol.glob, want, rtol=1e-6, atol=1e-9), np.abs(sol.glob - want).max()
    assert np.all(sol.glob[par3[:, 0] == p0, 0] == 0.0)  # clipped onto p itself: identical twin runs
    # get_trajectory: a given run, and a random member of a "cluster"
    tr = m.get_trajectory(prob, sol, 3)
    ref = orc.trajectory(psys, o, prob.ic[3], prob.par[3], ts)[0]
    assert np.allclose(np.asarray(tr), ref, rtol=1e-9, atol=1e-9) and tr.i_run == 3 and len(tr) == 400
    ca = np.arange(n) % 2
    db = m.DbscanResult(np.array([1]), ca, np.array([int((ca == 1).sum())]))
    tr2, sub, i_run = m.get_trajectory(prob, sol, db, 2, only_sol=False, rng=np.random.default_rng(1))
    assert ca[i_run] == 1 and sub.N_mc == 1
    # failure_handling = :Repeat: runs that hit maxiters get fresh initial conditions until they pass, at most 10 times
    draws = {"n": 0}
    rng = np.random.default_rng(4)

    def ic_gen():
        draws["n"] += 1
        return rng.uniform(-np.pi, np.pi)
    pars = m.kuramoto_network_parameters(0.5, np.full(6, 0.5), 6, np.ones((6, 6)))
    rp = m.ODEProblem(m.kuramoto_network, np.zeros(6), (0.0, 40.0), pars)
    ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std), failure_handling="Repeat")
    probr = m.DEMCBBProblem(rp, ic_gen, 6, pars, ("K", lambda: 2.0), ev, 0.9)
    first = draws["n"]
    with pytest.raises(m.MCBBError, match="More than 10 Repeats"):
        m.solve(probr, maxiters=5)  # nothing can succeed in five steps
    assert draws["n"] == first + 10 * 6 * 6  # ten redraws of all six runs, six dimensions each
    solr = m.solve(probr, maxiters=100000)
    assert np.all(solr.retcode == 0)


def test_host_mirror_parameter_variations_on_cpu(cpu_engine, orc):
    """solve() for the three kinds of parameter variation — one system parameter, several (MultiDimParameterVar) and a
    solver option (SolverParameterVar) — through the host mirror with the lane source as the engine: results come back
    sorted by the first parameter column and equal the oracle run by run."""
    m = cpu_engine
    rng = np.random.default_rng(6)
    N = 5
    E = cases.ring_incidence(N, [(0, 2)])
    pars = m.second_order_kuramoto_parameters(N, 0.1, 1.0, E, np.array([1.0, -1.0, 1.0, -1.0, 0.0]))
    rp = m.ODEProblem(m.second_order_kuramoto, np.zeros(2 * N), (0.0, 30.0), pars)
    ev = m.EvalOdeRun(state_filter=list(range(N + 1, 2 * N + 1)), eval_funcs=(m.mean, m.std))
    variations = [("coupling", lambda: rng.uniform(0.5, 3.0)),
                  [("coupling", lambda: rng.uniform(0.5, 3.0)), ("damping", lambda: rng.uniform(0.05, 0.4))],
                  m.SolverParameterVar("reltol", lambda: 10.0 ** rng.uniform(-7, -3))]
    for pv in variations:
        prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-1.0, 1.0), 8, pars, pv, ev, 0.6)
        ic0, par0 = prob.ic.copy(), prob.par.copy()
        sol = m.solve(prob)
        idx = np.argsort(par0[:, 0], kind="stable")
        assert np.array_equal(prob.par, par0[idx]) and np.array_equal(prob.ic, ic0[idx])
        psys, pfeat = prob.pars.pack(prob.par_names), prob.eval_ode_func.pack()
        ts = m.tsave_array(prob.p, 400, prob.rel_transient_time)
        same_steps = 0
        for i in range(prob.N_mc):
            kw = {"reltol": float(prob.par[i, 0])} if prob.solver_par else {}
            o = m.abi.solver_opts(m.abi.ALG_TSIT5, prob.p.tspan, **kw)
            ref = orc.solve_features(psys, o, pfeat, prob.ic[i:i + 1], prob.par[i:i + 1], ts)
            # a step whose error estimate sits at the accept threshold may flip between two correct FP64
            # implementations (DESIGN.md, step-size-controller sensitivity): such a run is compared loosely
            if np.array_equal(sol.nsteps[i], ref["nsteps"][0]):
                same_steps += 1
                assert np.allclose(sol.feat[i], ref["feat"][0], rtol=1e-6, atol=1e-7), i
            else:
                assert abs(int(sol.nsteps[i, 0]) - int(ref["nsteps"][0, 0])) <= 0.1 * ref["nsteps"][0, 0]
                assert np.allclose(sol.feat[i], ref["feat"][0], rtol=2e-2, atol=2e-2), i
        assert same_steps >= 6, same_steps
