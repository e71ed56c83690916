"""CPU tests (-m "not gpu") of the host logic: the C-ABI library loads and exports every symbol the header
declares, host-side helpers agree with the oracle, the Python mirror of the reference interface builds the same
problem shapes, and the multi-rank (world_size 2, gloo) DBSCAN bookkeeping reproduces the single-rank labels.
No compute kernel is launched here."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(mcbb):
    hdr = open(os.path.join(ROOT, "include", "mcbb_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mcbb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = mcbb.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(mcbb.EXPORTS) == declared
    assert b"sm_100a" in L.mcbb_version()


def test_library_is_cuda_sm100a_only():
    """The shipped library carries sm_100a SASS and no CPU implementation of the hot path."""
    import subprocess
    so = os.path.join(ROOT, "mcbb.jl_b200", "libmcbb_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert "orc_" not in syms  # the oracle is not linked into the product


def test_tsave_array_matches_oracle_and_julia_ranges(mcbb, orc):
    abi = mcbb.abi
    for alg, tspan, nt, tail in [(abi.ALG_TSIT5, (0.0, 40.0), 400, 0.9), (abi.ALG_TSIT5, (0.0, 1000.0), 400, 0.9),
                                 (abi.ALG_TSIT5, (0.0, 200.0), 400, 0.7), (abi.ALG_TSIT5, (0.0, 5000.0), 200, 0.8),
                                 (abi.ALG_FUNCTIONMAP, (0.0, 1000.0), 400, 0.8),
                                 (abi.ALG_FUNCTIONMAP, (0.0, 5000.0), 400, 0.8)]:
        o = abi.solver_opts(alg, tspan)
        n = C.c_int32(0)
        mcbb.check(mcbb.lib().mcbb_tsave_array(C.byref(o), nt, tail, abi.c_double_p(), C.byref(n)))
        ts = np.zeros(n.value)
        mcbb.check(mcbb.lib().mcbb_tsave_array(C.byref(o), nt, tail, ts.ctypes.data_as(abi.c_double_p), C.byref(n)))
        assert np.array_equal(ts, orc.tsave_array(o, nt, tail))
        if alg == abi.ALG_TSIT5:
            # Julia's float range can come out one short (e.g. tspan (0,200), tail 0.7: 399 points) - same quirk here
            assert abs(len(ts) - nt) <= 1 and ts[-1] < tspan[1]
        else:
            assert len(ts) == round((tspan[1] - tspan[0]) * (1 - tail)) and ts[-1] == tspan[1] - 1
    # Julia range semantics (twiceprecision.jl): rational lift makes 0.0:0.1:0.7 eight elements long
    assert len(mcbb.api._julia_range(0.0, 0.1, 0.7)) == 8
    assert len(mcbb.api._julia_range(0.25, 0.25, 2.5)) == 10
    assert len(mcbb.api._julia_range(1.0, 0.5, 0.0)) == 0


def test_error_reporting_without_gpu(mcbb):
    L = mcbb.lib()
    n = C.c_int32(0)
    assert L.mcbb_tsave_array(None, 400, 0.9, mcbb.abi.c_double_p(), C.byref(n)) != 0
    assert "NULL" in mcbb.last_error()
    with pytest.raises(mcbb.MCBBError, match="No per dimension measures"):
        mcbb.EvalOdeRun(eval_funcs=())
    with pytest.raises(mcbb.MCBBError, match="failure_handling symbol not known"):
        mcbb.EvalOdeRun(failure_handling="Bogus")


def test_no_cpu_fallback(mcbb):
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tests import cases
    prob, w = cases.c1_problem(mcbb, n_ics=4)
    with pytest.raises(mcbb.MCBBError):
        mcbb.solve(prob)
    with pytest.raises(mcbb.MCBBError):
        mcbb.dbscan(np.zeros((4, 4)), 1.0, 2)
    # the product never imports, links or executes the checker
    for f in ("api.py", "dist.py", "_lib.py", "__init__.py", "_abi.py"):
        src = open(os.path.join(ROOT, "mcbb.jl_b200", f)).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
        assert "libmcbb_oracle" not in src and "orc_" not in src, f


def test_problem_shapes_follow_the_reference(mcbb):
    m = mcbb
    N = 3
    pars = m.kuramoto_network_parameters(0.5, np.full(N, 0.5), N, np.ones((N, N)))
    rp = m.ODEProblem(m.kuramoto_network, np.zeros(N), (0.0, 10.0), pars)
    # ranges x range: CartesianIndices order, first IC dimension fastest, parameter slowest (setup_mc_prob.jl:872-880)
    prob = m.DEMCBBProblem(rp, [np.arange(0.0, 1.0, 0.5)] * N, pars, ("K", np.arange(1.0, 2.01, 0.5)), m.eval_ode_run, 0.9)
    assert prob.N_mc == 2 ** 3 * 3 and prob.ic.shape == (24, 3) and prob.par.shape == (24, 1)
    assert np.array_equal(prob.ic[:4, 0], [0, 0.5, 0, 0.5]) and np.array_equal(prob.ic[:4, 1], [0, 0, 0.5, 0.5])
    assert np.array_equal(prob.par[:9, 0], [1.0] * 8 + [1.5])
    # random ICs x range: identical ICs for every parameter value, IC index fastest (:929-969)
    rng = np.random.default_rng(0)
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-1, 1), 5, pars, ("K", [1.0, 2.0, 3.0]), m.eval_ode_run, 0.9)
    assert prob.N_mc == 15 and np.array_equal(prob.ic[:5], prob.ic[5:10]) and np.array_equal(prob.ic[:5], prob.ic[10:])
    assert np.array_equal(prob.par[:, 0], np.repeat([1.0, 2.0, 3.0], 5))
    # random x random (:888-923)
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-1, 1), 7, pars, ("K", lambda: rng.uniform(0, 5)), m.eval_ode_run, 0.9)
    assert prob.N_mc == 7 and len(np.unique(prob.par)) == 7
    assert m.parameter(prob).shape == (7,)
    with pytest.raises(m.MCBBError, match="does not belong"):
        m.ODEProblem(m.logistic, np.zeros(1), (0.0, 1.0), pars)


def test_sort_is_stable_and_in_place(mcbb):
    m = mcbb
    feat = np.asfortranarray(np.arange(6 * 2 * 2, dtype=float).reshape(6, 2, 2))
    sol = m.DEMCBBSol(feat.copy(order="F"), None, None, np.arange(6, dtype=np.int32), np.zeros((6, 2), dtype=np.int64), 400, None)

    class P:
        pass
    prob = P()
    prob.par = np.asfortranarray(np.array([[2.0], [1.0], [2.0], [1.0], [3.0], [1.0]]))
    prob.ic = np.asfortranarray(np.arange(12, dtype=float).reshape(6, 2))
    idx = m.sort_(sol, prob)
    assert list(idx) == [1, 3, 5, 0, 2, 4]  # sortperm is stable (setup_mc_prob.jl:520-536)
    assert np.array_equal(prob.par[:, 0], [1, 1, 1, 2, 2, 3]) and np.array_equal(sol.retcode, idx)
    assert np.array_equal(sol.feat, feat[idx]) and np.array_equal(prob.ic[:, 0], np.arange(12).reshape(6, 2)[idx, 0])
    assert len(sol[0]) == 2 and np.array_equal(sol[0][1], feat[1, :, 1])
    assert np.array_equal(m.get_measure(sol, 2), sol.feat[:, :, 1])


def test_feature_groups_and_cluster_membership_vs_oracle(mcbb, orc):
    m = mcbb
    rng = np.random.default_rng(3)
    n, nd = 40, 4
    sol = m.DEMCBBSol(np.asfortranarray(rng.normal(size=(n, nd, 2))), None, np.asfortranarray(rng.normal(size=(n, 1))),
                      np.zeros(n, np.int32), np.zeros((n, 2), np.int64), 400, None)

    class P:
        pass
    prob = P()
    prob.par = np.asfortranarray(np.sort(rng.uniform(0.5, 3.0, size=(n, 1)), axis=0))
    X, gl, gw = m.feature_groups(sol, prob, [1.0, 0.5, 0.25, 1.0])
    assert X.shape == (n, 2 * nd + 2) and list(gl) == [nd, nd, 1, 1] and list(gw) == [1.0, 0.5, 0.25, 1.0]
    assert np.array_equal(X[:, :nd], sol.feat[:, :, 0]) and np.array_equal(X[:, -1], prob.par[:, 0])
    with pytest.raises(m.MCBBError, match="Amount of weights"):
        m.feature_groups(sol, prob, [1.0, 1.0])
    Xr, _, _ = m.feature_groups(sol, prob, [1.0, 0.5, 0.25, 1.0], relative_parameter=True)
    p = prob.par[:, 0]
    assert np.allclose(Xr[:, -1], (p - p.min()) / p.max())  # sic: divides by max (eval_clustering.jl:576)
    a = rng.integers(0, 4, n)
    res = m.DbscanResult(np.array([1, 2, 3]), a, np.bincount(a, minlength=4)[1:])
    cm = m.cluster_membership(prob, res, 0.5, 0.1)
    mesh, ref = orc.cluster_membership(p, a, 3, 0.5, 0.1)
    assert np.allclose(cm.par, mesh) and np.allclose(cm.data, ref, equal_nan=True)
    assert m.cluster_n_noise(res) == int((a == 0).sum())


def test_shard_ranges(mcbb):
    d = mcbb.dist
    for n, w in [(10, 3), (7, 8), (1000000, 8), (5, 1)]:
        rs = [d.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
        assert d.shard_sizes(n, w) == [b - a for a, b in rs]


class _CheckerOps:
    """CPU stand-in for the stage kernels (numpy, same contracts as the C-ABI stage entry points), so the
    collectives and index bookkeeping of dist.dbscan_features_sharded can be exercised under gloo."""

    def __init__(self, lists=True):
        import torch
        self.torch = torch
        if lists:  # without it dist.dbscan_features_sharded falls back to the pair-space border stage
            self.rows_count_lists = self._rows_count_lists

    def _d(self, Xa, Xb, gl, gw):
        out = np.zeros((Xa.shape[1], Xb.shape[1]))
        f = 0
        for g, w in zip(gl, gw):
            out += w * np.abs(Xa[f:f + g, :, None] - Xb[f:f + g, None, :]).sum(0)
            f += g
        return out

    def rows_count(self, X, gl, gw, eps, minpts, r0, r1):
        Xn = X.numpy()
        return self.torch.from_numpy((self._d(Xn[:, r0:r1], Xn, gl, gw) < eps).sum(1).astype(np.int32))

    def _rows_count_lists(self, X, gl, gw, eps, minpts, r0, r1):
        Xn = X.numpy()
        near = self._d(Xn[:, r0:r1], Xn, gl, gw) < eps
        nbr = np.zeros((r1 - r0, minpts), dtype=np.int32)
        for i in range(r1 - r0):
            idx = np.nonzero(near[i])[0][::-1][:minpts]  # any order, only the first minpts
            nbr[i, :len(idx)] = idx
        cnt = near.sum(1).astype(np.int32)
        return self.torch.from_numpy(cnt), self.torch.from_numpy(nbr), self.torch.from_numpy(cnt.copy())

    def gather(self, X, idx):
        return X[:, idx.long()].contiguous()

    def rows_union(self, Xc, gl, gw, eps, c0, c1):
        Xn = Xc.numpy()
        n = Xn.shape[1]
        parent = np.arange(n, dtype=np.int32)
        D = self._d(Xn[:, c0:c1], Xn, gl, gw) < eps

        def find(x):
            while parent[x] != x:
                x = parent[x]
            return x
        for i in range(c0, c1):
            for j in np.nonzero(D[i - c0])[0]:
                if j > i:
                    a, b = find(i), find(j)
                    if a != b:
                        parent[max(a, b)] = min(a, b)
        return self.torch.from_numpy(parent)

    def rows_union_more(self, Xc, gl, gw, eps, c0, c1, parent):
        extra = self.rows_union(Xc, gl, gw, eps, c0, c1)
        self.merge_parents(parent, extra)

    def merge_parents(self, parent, other):
        p, o = parent.numpy(), other.numpy()

        def find(x):
            while p[x] != x:
                x = p[x]
            return x
        for i in range(len(p)):
            a, b = find(i), find(int(o[i]))
            if a != b:
                p[max(a, b)] = min(a, b)

    def rows_border(self, Xn, b0, b1, Xc, core_seed, gl, gw, eps):
        D = self._d(Xn.numpy()[:, b0:b1], Xc.numpy(), gl, gw) < eps
        seeds = core_seed.numpy()
        best = np.full(b1 - b0, np.iinfo(np.int32).max, dtype=np.int32)
        for i in range(b1 - b0):
            if D[i].any():
                best[i] = seeds[D[i]].min()
        return self.torch.from_numpy(best)

    def labels(self, root_label):
        rl = root_label.numpy()
        seeds = np.array(sorted({int(r) for r in rl if r >= 0 and rl[r] == r}), dtype=np.int64)
        rank = {s: k + 1 for k, s in enumerate(seeds)}
        a = np.array([rank[int(r)] if r >= 0 else 0 for r in rl], dtype=np.int64)
        counts = np.bincount(a, minlength=len(seeds) + 1)[1:]
        return self.torch.from_numpy(a), self.torch.from_numpy(seeds + 1), self.torch.from_numpy(counts)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, X, gl, gw, eps, minpts, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    m = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, s, c = m.dist.dbscan_features_sharded(torch.from_numpy(X), gl, gw, eps, minpts, rank, world, dist,
                                             backend_ops=_CheckerOps())
    q.put((rank, a.numpy().copy(), s.numpy().copy(), c.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_dbscan_world2_gloo(mcbb, orc):
    """world_size 2 over gloo: all-gather of counts, union-find merge of the per-rank parent arrays, all-gather of
    border labels -> labels identical to the sequential oracle on every rank."""
    import torch.multiprocessing as mp
    rng = np.random.default_rng(17)
    pts = np.concatenate([rng.normal(c, 0.3, size=(40, 2)) for c in [(0, 0), (3, 0), (0, 3)]] + [rng.uniform(-2, 5, size=(25, 2))])
    pts = pts[rng.permutation(len(pts))]
    X = np.ascontiguousarray(pts.T)  # [F, n]
    gl, gw = np.array([1, 1], dtype=np.int32), np.array([1.0, 0.5])
    D = orc.distance_matrix(pts, gl, gw)
    eps, minpts = 0.45, 4
    ref = orc.dbscan(D, eps, minpts)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, X, gl, gw, eps, minpts, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, s, c in got:
        assert np.array_equal(a, ref["assignments"]), rank
        assert np.array_equal(s, ref["seeds"]) and np.array_equal(c, ref["counts"])
    # single-rank path through the same code
    import torch
    a1, s1, c1 = mcbb.dist.dbscan_features_sharded(torch.from_numpy(X), gl, gw, eps, minpts, 0, 1, None,
                                                   backend_ops=_CheckerOps())
    assert np.array_equal(a1.numpy(), ref["assignments"])
    a2, s2, c2 = mcbb.dist.dbscan_features_sharded(torch.from_numpy(X), gl, gw, eps, minpts, 0, 1, None,
                                                   backend_ops=_CheckerOps(lists=False))
    assert np.array_equal(a2.numpy(), ref["assignments"]) and np.array_equal(s2.numpy(), ref["seeds"])


def test_hist_edges_and_ecdf_rows_match_oracle(mcbb, orc):
    """Histogram-mode preprocessing (host side, like the reference): common edges (IQR rule + both fallbacks) and
    the per-run Float32 ECDF rows — the product's vectorised numpy code against the oracle's scalar C loops."""
    import ctypes as C
    import warnings
    rng = np.random.default_rng(0)
    n, nd = 90, 12
    cases_ = [rng.normal(0, 1, (n, nd)) * rng.uniform(0.2, 2, (n, 1)), np.abs(rng.normal(0, 0.3, (n, nd))),
              np.full((n, nd), 0.25) + 1e-3 * (rng.random((n, nd)) < 0.05)]  # last: IQR == 0 -> fallback edges
    L = orc.lib()
    L.orc_hist_ecdf.argtypes = [orc.abi.c_double_p, C.c_int32, orc.abi.c_double_p, C.c_int32, C.c_int32,
                                C.POINTER(C.c_float)]
    for data in cases_:
        for k_bin in (1.0, 0.25):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                e1, b1 = mcbb.api._compute_hist_edges(data, nd, k_bin, 50)
            e2, b2 = orc.hist_edges(np.asfortranarray(data), nd, k_bin, 50)
            assert np.array_equal(e1, e2) and b1 == b2
            H = mcbb.api._compute_hist_weights(data, e1, True)
            for i in (0, 7, n - 1):
                row = np.ascontiguousarray(data[i])
                out = np.zeros(len(e1) - 1, dtype=np.float32)
                L.orc_hist_ecdf(row.ctypes.data_as(orc.abi.c_double_p), nd, e2.ctypes.data_as(orc.abi.c_double_p),
                                len(e2), 1, out.ctypes.data_as(C.POINTER(C.c_float)))
                assert np.array_equal(H[i], out, equal_nan=True)


def test_sharding_helpers_partition_exactly(mcbb):
    """dist.cyclic_order is a permutation that groups the 64-point tiles by owner (tile t -> rank t % world) in
    ascending order inside each group; dist.triangle_range blocks are tile aligned, contiguous, cover [0, m) and hold
    roughly equal shares of the pair triangle {col > row}."""
    import torch
    d = mcbb.dist
    for m_, world in ((1, 1), (63, 2), (64, 2), (1000, 3), (12345, 8)):
        perm = d.cyclic_order(m_, world)
        assert torch.equal(torch.sort(perm).values, torch.arange(m_))
        owner = (perm // 64) % world
        assert torch.all(owner[1:] >= owner[:-1])  # grouped by owner, rank 0 first
        for r in range(world):
            mine = perm[owner == r]
            assert torch.all(mine[1:] > mine[:-1])  # original order kept inside a group
        # contiguous shards of the permuted order are what the ranks own
        sizes = d.shard_sizes(m_, world)
        assert sum(sizes) == m_ and max(sizes) - min(sizes) <= 1
    for m_, parts in ((100000, 8), (100000, 32), (5000, 3), (70, 4)):
        edges = [d.triangle_range(m_, k, parts) for k in range(parts)]
        assert edges[0][0] == 0 and edges[-1][1] == m_
        for (a0, a1), (b0, b1) in zip(edges[:-1], edges[1:]):
            assert a1 == b0 and a0 % 64 == 0 and a0 <= a1
        if m_ >= 100000:
            area = [sum(m_ - r for r in (lo, hi - 1)) * (hi - lo) / 2 for lo, hi in edges if hi > lo]
            assert max(area) / min(area) < 1.25


def test_nonzero_sparse_matrix_semantics(mcbb):
    """NonzeroSparseMatrix (src/sparse_clustering.jl:37-110): stored entries are value-shifted, everything else reads
    as `value`; host-side container only (the device fills it)."""
    rng = np.random.default_rng(4)
    n, thr = 37, 0.35
    pts = rng.random((n, 3))
    D = np.abs(pts[:, None, :] - pts[None, :, :]).sum(-1)
    keep = D < thr
    colptr = np.concatenate([[0], np.cumsum(keep.sum(0))]).astype(np.int64)
    rowval = np.concatenate([np.nonzero(keep[:, c])[0] for c in range(n)]).astype(np.int32)
    nzval = np.concatenate([D[keep[:, c], c] - thr for c in range(n)])
    S = mcbb.NonzeroSparseMatrix(colptr, rowval, nzval, thr, n)
    assert S.shape == (n, n)
    dense = S.toarray()
    assert np.allclose(dense[keep], D[keep]) and np.all(dense[~keep] == thr)
    for i, j in ((0, 0), (3, 17), (n - 1, 2), (5, 5)):
        assert np.isclose(S[i, j], D[i, j] if keep[i, j] else thr)
    with pytest.raises(mcbb.MCBBError, match="sparse_threshold"):
        mcbb.distance_matrix_sparse(None, None, [1.0])  # the default threshold (Inf) is rejected before any work


def test_fixed_point_coupling_sum_identities():
    """The arithmetic claims behind the tensor-core coupling sums (csrc/fixedpoint.cuh, batched_umma.cu), restated in
    numpy / Python integers: the magic-number split gives u = round(x 2^50) + 2^51 and its radix-256 digits; digit sums
    over up to 128 unit couplings recombine EXACTLY with the 31-bit packing used on the device (three sums per 32-bit
    word, one 64-bit multiply-add, degree correction in the high word), and the single conversion to double is the
    correctly rounded sum — i.e. more accurate than any FP64 summation order."""
    from fractions import Fraction
    rng = np.random.default_rng(123)
    magic = 6755399441055744.0  # 1.5 * 2^52
    for deg in (1, 7, 20, 100, 128):
        x = rng.uniform(-1.0, 1.0, size=deg)
        x[0] = 1.0 if deg > 1 else x[0]   # extremes of the range
        x[-1] = -1.0
        t = x * 1125899906842624.0 + magic                       # one FMA on the device; exact here as well
        bits = t.view(np.uint64)
        lo = (bits & 0xFFFFFFFF).astype(np.uint64)
        hi = ((bits >> 32) & 0xFFFFF).astype(np.uint64)
        u = (hi << 32) | lo
        expect_u = np.array([int(np.rint(v * 2.0 ** 50)) + 2 ** 51 for v in x], dtype=object)
        assert all(int(a) == int(b) for a, b in zip(u, expect_u))
        digits = np.array([[(int(ui) >> (8 * k)) & 0xFF for k in range(8)] for ui in u])
        assert np.all(digits[:, 7] == 0) and np.all(digits[:, 6] <= 15)
        v = digits.sum(axis=0)                                   # what the int8 GEMM accumulates per digit column
        assert v.max() <= 128 * 255 < 2 ** 15
        # device recombination (32-bit wrap-around arithmetic made explicit)
        m32 = 0xFFFFFFFF
        lo3 = (int(v[0]) + (int(v[1]) << 8) + (int(v[2]) << 16)) & m32
        mid3 = (int(v[3]) + (int(v[4]) << 8) + (int(v[5]) << 16)) & m32
        assert lo3 == int(v[0]) + (int(v[1]) << 8) + (int(v[2]) << 16) < 2 ** 31   # no wrap: the packing is lossless
        w = mid3 * 16777216 + lo3
        degh = deg << 19
        hi_word = ((w >> 32) + ((int(v[6]) << 16) - degh))
        val = (hi_word << 32) | (w & m32)
        exact = sum(int(ui) for ui in u) - deg * 2 ** 51
        assert val == exact
        # one rounding: double(val) * 2^-50 against the exact rational sum of the rounded terms
        got = float(val) * 2.0 ** -50
        want = Fraction(exact, 2 ** 50)
        assert abs(Fraction(got) - want) <= Fraction(abs(exact), 2 ** 50) * Fraction(1, 2 ** 53) + 0
        # and the fixed-point terms are within 2^-51 of the true values
        assert np.max(np.abs(np.array([float(Fraction(int(ui) - 2 ** 51, 2 ** 50)) for ui in u]) - x)) <= 2.0 ** -51


def test_multidim_parameter_var_problem_and_cluster_membership(mcbb):
    """MultiDimParameterVar (setup_mc_prob.jl:176-236, 850-969) and the multi-parameter cluster_membership
    (eval_clustering.jl:1206-1265) against literal loops."""
    m = mcbb
    N = 4
    E = np.zeros((N, N))
    for e in range(N):
        E[e, e], E[(e + 1) % N, e] = -1.0, 1.0
    pars = m.second_order_kuramoto_parameters(N, 0.1, 1.0, E, np.array([1.0, -1.0, 1.0, -1.0]))
    rp = m.ODEProblem(m.second_order_kuramoto, np.zeros(2 * N), (0.0, 10.0), pars)
    ev = m.EvalOdeRun(eval_funcs=(m.mean, m.std))
    # ranges x ranges: cartesian product, IC dimensions first, first index fastest
    ic_ranges = [[0.0, 1.0]] + [[0.5]] * (2 * N - 1)
    prob = m.DEMCBBProblem(rp, ic_ranges, pars, [("coupling", [1.0, 2.0, 3.0]), ("damping", [0.1, 0.2])], ev, 0.5)
    assert prob.N_mc == 12 and prob.par.shape == (12, 2) and prob.par_names == ["coupling", "damping"]
    assert prob.ic[:, 0].tolist() == [0.0, 1.0] * 6
    assert prob.par[:, 0].tolist() == [1.0, 1.0, 2.0, 2.0, 3.0, 3.0] * 2
    assert prob.par[:, 1].tolist() == [0.1] * 6 + [0.2] * 6
    psys = prob.pars.pack(prob.par_names)
    assert psys.desc.n_par == 2 and list(psys.desc.par_slot)[:2] == [1, 0]
    # generators x generators
    rng = np.random.default_rng(0)
    mv = m.MultiDimParameterVar([("coupling", lambda: rng.uniform(0, 4)), ("damping", lambda: rng.uniform(0, 1))])
    prob = m.DEMCBBProblem(rp, lambda: rng.uniform(-1, 1), 500, pars, mv, ev, 0.5)
    assert prob.par.shape == (500, 2) and len(mv) == 2 and mv[1][0] == "damping"
    assert prob.par[:, 0].max() > 3 and prob.par[:, 1].max() <= 1
    with pytest.raises(m.MCBBError, match="only supported for N_par=1"):
        m.DEMCBBProblem(rp, lambda: 0.0, 5, pars, [("coupling", [1.0, 2.0]), ("damping", [0.1])], ev, 0.5)
    with pytest.raises(m.MCBBError):
        m.MultiDimParameterVar([("coupling", [1.0]), ("damping", lambda: 0.1)])
    # sliding windows in two parameters
    ca = rng.integers(0, 4, size=500)
    counts = np.bincount(ca, minlength=4)[1:]
    db = m.DbscanResult(np.arange(3), ca, counts)
    ws, wo = [1.0, 0.25], [0.5, 0.125]
    cm = m.cluster_membership(prob, db, ws, wo)
    assert cm.multidim_flag and cm.data.ndim == 3 and cm.data.shape[2] == 4 and cm.par.shape == cm.data.shape[:2] + (2,)
    p1, p2 = prob.par[:, 0], prob.par[:, 1]
    n1 = len(np.arange(p1.min(), p1.max() - ws[0] + 1e-12, wo[0]))
    assert abs(cm.data.shape[0] - n1) <= 1
    for i in range(cm.data.shape[0]):
        for j in range(cm.data.shape[1]):
            lo1, lo2 = cm.par[i, j]
            ind = (p1 > lo1) & (p1 < lo1 + ws[0]) & (p2 > lo2) & (p2 < lo2 + ws[1])
            want = np.bincount(ca[ind], minlength=4) / max(1, ind.sum())
            if ind.sum() > 0:
                assert np.allclose(cm.data[i, j], want)
    raw = m.cluster_membership(prob, db, ws, wo, normalize=False, min_members=int(counts.min()))
    assert raw.data.shape[2] == 1 + (counts > counts.min()).sum() - (0 if (ca == 0).sum() > counts.min() else 1)
    with pytest.raises(m.MCBBError, match="as many elements"):
        m.cluster_membership(prob, db, 1.0, 0.5)


def test_normalize_sort_show_results_get_measure(mcbb):
    """normalize / sort / show_results / get_measure(state_filter) (setup_mc_prob.jl:505-660) on a synthetic
    solution: literal per-run loops as in the reference."""
    m = mcbb
    rng = np.random.default_rng(5)
    n, nd = 40, 3
    feat = np.asfortranarray(rng.normal(size=(n, nd, 2)) * [3.0, 0.2] + [1.0, 5.0])
    mat = np.asfortranarray(rng.random((n, 30)))
    glob = np.asfortranarray(rng.normal(size=(n, 1)))
    sol = m.DEMCBBSol(feat.copy(order="F"), mat.copy(order="F"), glob.copy(order="F"), np.zeros(n, np.int32),
                      np.zeros((n, 2), np.int64), 400, None)
    assert (sol.N_meas_dim, sol.N_meas_matrix, sol.N_meas_global, sol.N_meas) == (2, 1, 1, 4)
    with pytest.warns(UserWarning, match="Global measures"):
        nsol = m.normalize(sol)
    for q in (1, 2):
        a = m.get_measure(nsol, q)
        want = (feat[:, :, q - 1] - feat[:, :, q - 1].min()) / (feat[:, :, q - 1].max() - feat[:, :, q - 1].min())
        assert np.array_equal(a, want) and a.min() == 0.0 and a.max() == 1.0
    assert np.array_equal(m.get_measure(nsol, 3), (mat - mat.min()) / (mat.max() - mat.min()))
    assert np.array_equal(m.get_measure(nsol, 4), glob[:, 0]) and np.array_equal(sol.feat, feat)  # input untouched
    only2 = m.normalize(sol, [2])
    assert np.array_equal(only2.feat[:, :, 0], feat[:, :, 0]) and only2.feat[:, :, 1].max() == 1.0
    assert np.array_equal(m.get_measure(sol, 1, state_filter=[3, 1]), feat[:, [2, 0], 0])

    class P:  # the fields sort / parameter read
        pass
    prob = P()
    prob.ic = np.asfortranarray(rng.normal(size=(n, nd)))
    prob.par = np.asfortranarray(rng.permutation(np.repeat(np.arange(10.0), 4))[:, None])
    prob.N_mc = n
    ic0, par0 = prob.ic.copy(), prob.par.copy()
    ssol, sprob = m.sort(sol, prob)
    idx = np.argsort(par0[:, 0], kind="stable")
    assert np.array_equal(sprob.par, par0[idx]) and np.array_equal(sprob.ic, ic0[idx])
    assert np.array_equal(ssol.feat, feat[idx]) and np.array_equal(ssol.matrix, mat[idx])
    assert np.array_equal(prob.par, par0) and np.array_equal(sol.feat, feat)  # copies: inputs untouched
    got = m.show_results(sol, prob, 2.0, 5.0)
    rows = [r for r in idx if 2.0 < par0[r, 0] < 5.0]
    assert len(got) == len(rows) == 8
    for g, r in zip(got, rows):
        assert np.array_equal(g[0], feat[r, :, 0]) and np.array_equal(g[2], mat[r]) and g[3] == glob[r, 0]
    m.sort_prob_(prob)
    assert np.array_equal(prob.par, par0[idx])


def test_hist_edges_nbin_and_explicit_bin_edges(mcbb):
    """nbin / bin_edges of distance_matrix(...; histograms=true) (eval_clustering.jl:607-642): (min-bw):bw:(max+bw)
    with bw = (max-min)/(nbin-1), or the edges as given with bin_width = edges[2]-edges[1]; both at once is an error."""
    api = mcbb.api
    rng = np.random.default_rng(2)
    data = rng.normal(size=(50, 7))
    e, bw = api._compute_hist_edges(data, 7, 1.0, 50, nbin=12)
    assert bw == (data.max() - data.min()) / 11 and e[0] == data.min() - bw and abs(e[1] - e[0] - bw) < 1e-15
    assert 13 <= len(e) <= 14 and e[-1] >= data.max()
    e2, bw2 = api._compute_hist_edges(data, 7, 1.0, 50, bin_edges=np.arange(-5.0, 5.01, 0.5))
    assert bw2 == 0.5 and len(e2) == 21 and e2[0] == -5.0
    with pytest.raises(mcbb.MCBBError, match="choose only one"):
        api._compute_hist_edges(data, 7, 1.0, 50, nbin=5, bin_edges=[0.0, 1.0])
    with pytest.raises(mcbb.MCBBError):
        api._compute_hist_edges(data, 7, 1.0, 50, nbin=1)
    # through the feature assembly: one ECDF group per measure with weight w * bin_width
    feat = np.asfortranarray(rng.normal(size=(50, 7, 2)))
    sol = mcbb.DEMCBBSol(feat, None, None, np.zeros(50, np.int32), np.zeros((50, 2), np.int64), 400, None)

    class P:
        pass
    prob = P()
    prob.par = np.asfortranarray(rng.random((50, 1)))
    edges = [np.arange(-5.0, 5.01, 0.5), None]
    X, gl, gw, he, bws = api._assemble_features(sol, prob, [1.0, 2.0, 1.0], False, True, True, 1.0, 50, None, edges)
    assert list(gl)[:2] == [20, len(he[1]) - 1] and gw[0] == 0.5 and gw[1] == 2.0 * bws[1] and bws[0] == 0.5
    with pytest.raises(mcbb.MCBBError, match="one entry"):
        api._assemble_features(sol, prob, [1.0, 2.0, 1.0], False, True, True, 1.0, 50, None, edges[:1])


def test_cluster_membership_result_operations(mcbb):
    """Indexing, sort! and sum of ClusterMembershipResult (eval_clustering.jl:1290-1338)."""
    rng = np.random.default_rng(1)
    data = rng.random((9, 4)) * [1.0, 5.0, 0.1, 2.0]
    cm = mcbb.ClusterMembershipResult(np.arange(9.0), data.copy(), False)
    assert cm.shape == (9, 4) and np.array_equal(cm[2].data, data[:, 1:2]) and np.array_equal(cm[[1, 4]].data, data[:, [0, 3]])
    s = cm.sum([2, 3])
    assert s.data.shape == (9, 1) and np.allclose(s.data[:, 0], data[:, 1] + data[:, 2])
    order = cm.sort_()
    assert list(order) == [2, 0, 3, 1] and np.array_equal(cm.data, data[:, [2, 0, 3, 1]])
    cm2 = mcbb.ClusterMembershipResult(np.arange(9.0), data.copy(), False)
    assert list(cm2.sort_(ignore_first=True)) == [0, 2, 3, 1]
    multi = mcbb.ClusterMembershipResult(np.zeros((3, 3, 2)), rng.random((3, 3, 4)), True)
    assert multi[3].data.shape == (3, 3, 1) and multi.sum([1, 2]).data.shape == (9, 1)


def test_measure_on_parameter_sliding_window(mcbb):
    """measure_on_parameter_sliding_window (eval_clustering.jl:1369-1430) against the reference's literal loops."""
    m = mcbb
    rng = np.random.default_rng(8)
    n, nd = 120, 3
    feat = np.asfortranarray(rng.normal(size=(n, nd, 2)))
    glob = np.asfortranarray(rng.normal(size=(n, 1)))
    sol = m.DEMCBBSol(feat, None, glob, np.zeros(n, np.int32), np.zeros((n, 2), np.int64), 400, None)

    class P:
        pass
    prob = P()
    prob.N_mc = n
    prob.par = np.asfortranarray(np.sort(rng.uniform(0, 4, n))[:, None])
    ca = rng.integers(0, 3, n)
    db = m.DbscanResult(np.arange(2), ca, np.bincount(ca, minlength=3)[1:])
    ws, wo = 1.0, 0.5
    mesh, meas = m.measure_on_parameter_sliding_window(prob, sol, 2, db, ws, wo)
    assert mesh.ndim == 1 and meas.shape == (3, nd, len(mesh))
    p = prob.par[:, 0]
    for k, lo in enumerate(mesh):
        ind = (p > lo) & (p < lo + ws)
        for c in range(3):
            sel = ind & (ca == c)
            want = feat[sel, :, 1].mean(axis=0) if sel.any() else np.full(nd, np.nan)
            assert np.allclose(meas[c, :, k], want, equal_nan=True)
    mesh2, meas2 = m.measure_on_parameter_sliding_window(prob, sol, 3, ws, wo)  # the global measure, all runs
    assert meas2.shape == (1, 1, len(mesh)) and np.array_equal(mesh2, mesh)
    ind0 = (p > mesh[0]) & (p < mesh[0] + ws)
    assert np.isclose(meas2[0, 0, 0], glob[ind0, 0].mean())
    # a cluster absent from a window reads NaN
    ca2 = np.where(p < 2.0, 1, 2)
    db2 = m.DbscanResult(np.arange(2), ca2, np.bincount(ca2, minlength=3)[1:])
    _, meas3 = m.measure_on_parameter_sliding_window(prob, sol, 1, db2, ws, wo)
    assert np.isnan(meas3[0]).all() and np.isnan(meas3[2, :, 0]).all() and np.isfinite(meas3[1, :, 0]).all()


def test_cluster_means_measures_and_per_cluster_getters(mcbb):
    """cluster_means / cluster_measure_mean / cluster_measure_std / get_measure by cluster / cluster_measures
    (eval_clustering.jl:770-906) against literal loops."""
    m = mcbb
    rng = np.random.default_rng(13)
    n, nd = 90, 4
    feat = np.asfortranarray(rng.normal(size=(n, nd, 2)))
    glob = np.asfortranarray(rng.normal(size=(n, 1)))
    sol = m.DEMCBBSol(feat, None, glob, np.zeros(n, np.int32), np.zeros((n, 2), np.int64), 400, None)
    ca = rng.integers(0, 3, n)
    db = m.DbscanResult(np.arange(2), ca, np.bincount(ca, minlength=3)[1:])
    cmn = m.cluster_means(sol, db)
    assert cmn.shape == (3, 2, nd)
    for c in range(3):
        assert np.allclose(cmn[c], feat[ca == c].mean(axis=0).T)
    assert np.allclose(m.cluster_measure_mean(sol, db, 2), [feat[ca == c, :, 1].mean() for c in range(3)])
    assert np.allclose(m.cluster_measure_std(sol, db, 1), [feat[ca == c, :, 0].std(ddof=1) for c in range(3)])
    assert np.allclose(m.cluster_measure_mean(sol, db, 3), [glob[ca == c, 0].mean() for c in range(3)])

    class P:
        pass
    prob = P()
    prob.N_mc = n
    prob.par = np.asfortranarray(rng.uniform(0, 3, n)[:, None])
    got = m.get_cluster_measure(sol, 1, db, 2, prob, 1.0, 2.0)
    sel = (ca == 1) & (prob.par[:, 0] > 1.0) & (prob.par[:, 0] < 2.0)
    assert np.array_equal(got, feat[sel, :, 0]) and np.array_equal(m.get_cluster_measure(sol, 3, db, 1), glob[ca == 0, 0])
    res = m.cluster_measures(prob, sol, db, 1.0, 0.5)
    nw = len(res.par)
    assert res.cluster_measures.shape == (3, 2, nd, nw) and res.cluster_measures_global.shape == (3, 1, nw)
    assert not res.multidim_flag
    lo = res.par[1]
    ind = (prob.par[:, 0] > lo) & (prob.par[:, 0] < lo + 1.0)
    assert np.allclose(res.cluster_measures[2, 1, :, 1], feat[ind & (ca == 2), :, 1].mean(axis=0))
    assert np.isclose(res.cluster_measures_global[0, 0, 1], glob[ind & (ca == 0), 0].mean())
    with pytest.warns(UserWarning, match="No clusters"):
        assert m.cluster_means(sol, m.DbscanResult(np.zeros(0, int), np.zeros(n, int), np.zeros(0, int))) is False


def test_quantile_neighbours_and_lerp_reproduce_numpy_percentile(mcbb):
    """The device path of the histogram edges (dist.histogram_features_device) gets the two order statistics around a
    type-7 quantile from mcbb_order_stats and interpolates on the host: api._quantile_neighbours + api._lerp must be
    numpy's percentile (the host path's StatsBase.iqr stand-in) bit for bit, and _hist_edges_from_stats the rule of
    _compute_hist_edges."""
    api = mcbb.api
    rng = np.random.default_rng(4)
    for n in (2, 5, 17, 1000, 20001):
        x = np.sort(rng.normal(size=n) * rng.uniform(0.1, 50))
        for q in (0.25, 0.75, 0.5, 0.0, 1.0):
            lo, hi, g = api._quantile_neighbours(n, q)
            assert api._lerp(x[lo], x[hi], g) == np.percentile(x, 100 * q), (n, q)
    data = rng.normal(size=(400, 9))
    srt = np.sort(data.ravel())
    l25, h25, g25 = api._quantile_neighbours(srt.size, 0.25)
    l75, h75, g75 = api._quantile_neighbours(srt.size, 0.75)
    e_ref, bw_ref = api._compute_hist_edges(data, 9, 0.5, 50)
    e, bw = api._hist_edges_from_stats(srt[0], srt[-1], api._lerp(srt[l25], srt[h25], g25), api._lerp(srt[l75], srt[h75], g75),
                                       9, 0.5, 50)
    assert bw == bw_ref and np.array_equal(e, e_ref)
