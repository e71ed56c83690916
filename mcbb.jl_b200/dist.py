"""One-process-per-GPU sharding of the hot path (torch.distributed is plumbing only: device memory + NCCL/gloo).

* integrate + featurize: runs are independent -> contiguous blocks of run indices per rank, no collective while
  computing; the small per-run feature vectors are all-gathered afterwards.
* DBSCAN: the pair space shards by point rows; the stage entry points of the C-ABI (mcbb_dbscan_rows_*_dev) do the
  work on the owned rows, this module does the collectives between them:
      counts  -> all_gather -> core flags / index lists
      parents -> all_gather -> mcbb_dbscan_merge_parents_dev (union-find merge of boundary labels)
      border  -> all_gather -> root labels -> mcbb_dbscan_labels_dev
"""
import ctypes as C

import numpy as np

from . import _abi as abi
from ._lib import check, lib


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items owned by `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def _all_gather_rows(t, sizes, dist_mod, group=None):
    """all_gather of row blocks with per-rank sizes along dim 0 (pads to the largest block)."""
    import torch
    world = len(sizes)
    if world == 1:
        return t
    mx = max(sizes)
    pad = t.new_zeros((mx,) + tuple(t.shape[1:]))
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist_mod.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def solve_features_device(psys, opts, pfeat, ic_t, par_t, tsave_t):
    """mcbb_solve_features_dev on torch CUDA tensors (run index contiguous: ic_t is [n_dim, n_mc], par_t
    [n_par, n_mc]).  Returns (feat [n_meas*n_filter, n_mc], glob, retcode, nsteps [2, n_mc])."""
    import torch
    n_mc = ic_t.shape[1]
    nf = pfeat.n_filter(psys)
    dev = ic_t.device
    feat = torch.empty((pfeat.n_meas * nf, n_mc), dtype=torch.float64, device=dev)
    glob = torch.empty((max(pfeat.n_global, 1), n_mc), dtype=torch.float64, device=dev)
    ret = torch.empty((n_mc,), dtype=torch.int32, device=dev)
    nsteps = torch.empty((2, n_mc), dtype=torch.int64, device=dev)
    check(lib().mcbb_solve_features_dev(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n_mc, _ptr(ic_t),
                                        _ptr(par_t), _ptr(tsave_t), tsave_t.numel(), _ptr(feat), C.c_void_p(0),
                                        _ptr(glob) if pfeat.n_global else C.c_void_p(0), _ptr(ret), _ptr(nsteps),
                                        _stream()))
    return feat, (glob if pfeat.n_global else None), ret, nsteps


def dbscan_features_device(X_t, group_len, group_weight, eps, minpts):
    """Single-GPU fused DBSCAN on a torch CUDA feature matrix X_t [F, n].  Returns (assignments, seeds, counts)."""
    import torch
    n = X_t.shape[1]
    gl = np.ascontiguousarray(group_len, dtype=np.int32)
    gw = np.ascontiguousarray(group_weight, dtype=np.float64)
    a = torch.empty((n,), dtype=torch.int64, device=X_t.device)
    s = torch.empty((n,), dtype=torch.int64, device=X_t.device)
    c = torch.empty((n,), dtype=torch.int64, device=X_t.device)
    k = np.zeros(1, dtype=np.int64)
    check(lib().mcbb_dbscan_features_dev(_ptr(X_t), n, len(gl), gl.ctypes.data_as(abi.c_int32_p),
                                         gw.ctypes.data_as(abi.c_double_p), float(eps), int(minpts), _ptr(a), _ptr(s),
                                         _ptr(c), k.ctypes.data_as(abi.c_int64_p), _stream()))
    return a, s[: k[0]], c[: k[0]]


def order_stats_device(v_t, ranks):
    """Values at the given 0-based ranks of the ascending order of a torch CUDA tensor, and its NaN count
    (mcbb_order_stats_dev: one radix sort on the device, a few numbers back)."""
    rk = np.ascontiguousarray(ranks, dtype=np.int64)
    out = np.zeros(len(rk), dtype=np.float64)
    nn = np.zeros(1, dtype=np.int64)
    check(lib().mcbb_order_stats_dev(_ptr(v_t), v_t.numel(), len(rk), rk.ctypes.data_as(abi.c_int64_p),
                                     out.ctypes.data_as(abi.c_double_p), nn.ctypes.data_as(abi.c_int64_p), _stream()))
    return out, int(nn[0])


def histogram_features_device(feat_t, n_meas_dim, n_dim, weights, par_t, order_t=None, normalize=True, k_bin=1,
                              nbin_default=50, use_ecdf=True, skip_zero_weight=True):
    """Device-resident counterpart of `normalize(sol)` + the feature assembly of `distance_matrix(sol, prob, weights;
    histograms = true, k_bin)` (src/setup_mc_prob.jl:635-660, src/eval_clustering.jl:187-211, 587-642) for the per-dimension
    measures and the parameter columns: feat_t [n_meas_dim * n_dim, n] as mcbb_solve_features_dev wrote it, par_t
    [n_par, n]; order_t (int64 [n], optional) = sortperm of the runs (`sort!(sol, prob)`).  Per measure: extrema and
    quartile neighbours by one device sort (mcbb_order_stats_dev) -> edges on the host (a Julia range from four numbers) ->
    Float32 ECDF rows on the device (mcbb_hist_rows_dev).  The measures never leave the device.  Returns
    (X_t [F, n], group_len, group_weight, hist_edges, bin_widths): bit-identical to the host path (api._assemble_features),
    ready for dbscan_features_device / dbscan_features_sharded.  weights = one per measure, then one per parameter."""
    import torch

    from . import api
    n = feat_t.shape[1]
    n_par = par_t.shape[0]
    if len(weights) != n_meas_dim + n_par:
        raise api.MCBBError("Amount of weights not the same as Measures + Parameters")
    if order_t is not None:
        order_t = order_t.to(torch.int64).contiguous()
    blocks, glen, gw, hist_edges, bin_widths = [], [], [], [], []
    for mi in range(n_meas_dim):
        slab = feat_t[mi * n_dim:(mi + 1) * n_dim]
        assert slab.is_contiguous()
        count = slab.numel()
        l25, h25, g25 = api._quantile_neighbours(count, 0.25)
        l75, h75, g75 = api._quantile_neighbours(count, 0.75)
        st, n_nan = order_stats_device(slab, [0, count - 1, l25, h25, l75, h75])
        finite = n_nan == 0 and np.isfinite(st[0]) and np.isfinite(st[1])
        if skip_zero_weight and weights[mi] == 0 and finite:
            hist_edges.append(None)
            bin_widths.append(None)
            continue
        if n_nan:
            raise api.MCBBError("histogram features: measure %d holds NaN values" % (mi + 1))
        lo, rng = (st[0], st[1] - st[0]) if normalize else (0.0, 1.0)
        with np.errstate(invalid="ignore", divide="ignore"):
            v = (st - lo) / rng if normalize else st
        q25, q75 = api._lerp(v[2], v[3], g25), api._lerp(v[4], v[5], g75)
        edges, bw = api._hist_edges_from_stats(float(v[0]), float(v[1]), q25, q75, n_dim, k_bin, nbin_default)
        edges = np.ascontiguousarray(edges, dtype=np.float64)
        rows = torch.empty((len(edges) - 1, n), dtype=torch.float64, device=feat_t.device)
        check(lib().mcbb_hist_rows_dev(_ptr(slab), n, n_dim, _ptr(order_t) if order_t is not None else C.c_void_p(0),
                                       1 if normalize else 0, float(lo), float(rng), edges.ctypes.data_as(abi.c_double_p),
                                       len(edges), 1 if use_ecdf else 0, _ptr(rows), _stream()))
        blocks.append(rows)
        glen.append(len(edges) - 1)
        gw.append(weights[mi] * bw)
        hist_edges.append(edges)
        bin_widths.append(bw)
    par_s = par_t if order_t is None else par_t[:, order_t]
    for p in range(n_par):
        blocks.append(par_s[p:p + 1].to(torch.float64))
        glen.append(1)
        gw.append(weights[n_meas_dim + p])
    X_t = torch.cat(blocks, dim=0).contiguous()
    return X_t, np.asarray(glen, dtype=np.int32), np.asarray(gw, dtype=np.float64), hist_edges, bin_widths


_TILE = 64  # pair-tile edge of the stage kernels (csrc/pair_tiles.cuh)


def cyclic_order(m, world, device=None):
    """Permutation of range(m) that lists the 64-point tiles owned by rank 0, then rank 1, ... where tile t belongs
    to rank t % world.  Runs arrive sorted by parameter, and the cost of a row (how early it saturates, how many of
    its tiles exit early) varies strongly along that order: contiguous shards of the permuted order are balanced."""
    import torch
    idx = torch.arange(m, device=device)
    owner = (idx // _TILE) % world
    return torch.sort(owner, stable=True).indices


def triangle_range(m, rank, world):
    """Row block [lo, hi) (tile aligned) of the pair triangle {col > row} holding ~1/world of its pairs."""
    def edge(k):
        if k >= world:
            return m
        b = int(m * (1.0 - (1.0 - k / world) ** 0.5))
        return min(m, (b // _TILE) * _TILE)
    return edge(rank), edge(rank + 1)


def dbscan_features_sharded(X_t, group_len, group_weight, eps, minpts, rank, world, dist_mod=None, group=None,
                            backend_ops=None):
    """Row-sharded DBSCAN over `world` ranks; every rank holds the full feature matrix X_t [F, n] (after the
    feature all-gather) and returns identical (assignments, seeds, counts) tensors.

    Balance: every stage runs on a tile-cyclic permutation of its rows (`cyclic_order`); counts and border labels are
    order independent and are scattered back, the union stage only yields the COMPONENTS of the core graph in the
    permuted index space (which pairs a rank visits does not matter) and the seed of a component is then taken as the
    smallest ORIGINAL index among its members, as the sequential algorithm would.  The union stage's pair triangle is
    cut into equal-area row blocks (`triangle_range`).

    `backend_ops` lets the CPU gloo tests substitute the stage kernels with a checker implementation while the
    collectives / index bookkeeping under test stay the same code.
    """
    import torch
    ops = backend_ops or _CudaStageOps()
    F, n = X_t.shape
    gl = np.ascontiguousarray(group_len, dtype=np.int32)
    gw = np.ascontiguousarray(group_weight, dtype=np.float64)
    dev = X_t.device
    i32max = torch.iinfo(torch.int32).max
    import os
    import time
    trace = bool(os.environ.get("MCBB_TRACE")) and X_t.is_cuda
    t_last = [time.perf_counter()]

    def lap(what):
        if trace:
            torch.cuda.synchronize()
            now = time.perf_counter()
            print("[mcbb trace] rank %d sharded dbscan: %-28s %9.3f ms" % (rank, what, 1e3 * (now - t_last[0])),
                  file=__import__("sys").stderr, flush=True)
            t_last[0] = now
    # ---- stage 1: neighbour counts of the owned rows -------------------------------------------------------
    perm = cyclic_order(n, world, dev)
    Xp = ops.gather(X_t, perm.to(torch.int32))
    r0, r1 = shard_range(n, rank, world)
    lists = hasattr(ops, "rows_count_lists") and (r1 - r0) * minpts * 4 <= (512 << 20)
    if lists:  # also the first minpts neighbours of every owned row: the border stage becomes a lookup
        cnt_local, nbr_local, nfill_local = ops.rows_count_lists(Xp, gl, gw, eps, minpts, r0, r1)
    else:
        cnt_local = ops.rows_count(Xp, gl, gw, eps, minpts, r0, r1)
    lap("permute + count rows")
    cnt_perm = _all_gather_rows(cnt_local, shard_sizes(n, world), dist_mod, group)
    lap("all-gather counts")
    cnt = torch.empty_like(cnt_perm)
    cnt[perm] = cnt_perm
    del Xp
    core = cnt >= minpts
    core_idx = torch.nonzero(core).flatten().to(torch.int32)  # ascending = stable compaction
    non_idx = torch.nonzero(~core).flatten().to(torch.int32)
    n_core, n_non = core_idx.numel(), non_idx.numel()
    # ---- stage 2: components of the core graph, union-find merge across ranks -------------------------------
    root_label = torch.full((n,), -1, dtype=torch.int32, device=dev)
    if n_core > 0:
        lap("core flags, index lists")
        cp = cyclic_order(n_core, world, dev)
        core_orig = core_idx[cp].contiguous()  # original index of every permuted core position
        Xc = ops.gather(X_t, core_orig)
        lap("core flags + gathers")
        # the pair triangle {col > row} is cut into 4 * world equal-area row blocks, dealt out cyclically: how many
        # tiles a block can skip (points that already share a root) varies along the triangle
        parts = 4 if hasattr(ops, "rows_union_more") else 1
        c0, c1 = triangle_range(n_core, rank, world * parts)
        parent = ops.rows_union(Xc, gl, gw, eps, c0, c1)
        for k in range(1, parts):
            c0, c1 = triangle_range(n_core, rank + k * world, world * parts)
            ops.rows_union_more(Xc, gl, gw, eps, c0, c1, parent)
        lap("union rows")
        if world > 1:
            others = [torch.empty_like(parent) for _ in range(world)]
            dist_mod.all_gather(others, parent, group=group)
            for r in range(world):
                if r != rank:
                    ops.merge_parents(parent, others[r])
        lap("all-gather + merge parents")
        root = parent.to(torch.int64)
        while True:  # pointer jumping to the component representatives
            nxt = root[root]
            if torch.equal(nxt, root):
                break
            root = nxt
        # seed of a component = its smallest original index (sparse_clustering.jl:149-157 visits points in order)
        seed_of = torch.full((n_core,), i32max, dtype=torch.int32, device=dev)
        seed_of.scatter_reduce_(0, root, core_orig, reduce="amin", include_self=True)
        core_seed = seed_of[root].contiguous()  # per permuted core position
        root_label[core_orig.long()] = core_seed
        lap("roots + seeds")
        # ---- stage 3: border points ----------------------------------------------------------------------
        if n_non > 0 and lists:
            # a non-core point has fewer than minpts neighbours, all of them recorded (as permuted column indices)
            # by the rank that owned its row in stage 1: its label is the smallest seed among the core ones
            seed_of_point = torch.full((n,), i32max, dtype=torch.int32, device=dev)
            seed_of_point[core_orig.long()] = core_seed
            valid = torch.arange(minpts, device=dev)[None, :] < nfill_local[:, None]
            q = perm[nbr_local.long().clamp_(0, n - 1)]  # original index of every recorded neighbour
            best_local = torch.where(valid, seed_of_point[q], torch.full_like(nbr_local, i32max)).amin(dim=1)
            best_perm = _all_gather_rows(best_local, shard_sizes(n, world), dist_mod, group)
            best = torch.empty_like(best_perm)
            best[perm] = best_perm  # by original point index
            lab = torch.where(best == i32max, torch.full_like(best, -1), best)
            root_label[non_idx.long()] = lab[non_idx.long()]
        elif n_non > 0:
            bp = cyclic_order(n_non, world, dev)
            non_orig = non_idx[bp].contiguous()
            Xn = ops.gather(X_t, non_orig)
            b0, b1 = shard_range(n_non, rank, world)
            best_local = ops.rows_border(Xn, b0, b1, Xc, core_seed, gl, gw, eps)
            best = _all_gather_rows(best_local, shard_sizes(n_non, world), dist_mod, group)
            lab = torch.where(best == i32max, torch.full_like(best, -1), best)
            root_label[non_orig.long()] = lab
    lap("border")
    out = ops.labels(root_label)
    lap("labels")
    return out


class _CudaStageOps:
    """Stage kernels through the C-ABI (device pointers of torch CUDA tensors)."""

    def rows_count(self, X, gl, gw, eps, minpts, r0, r1):
        import torch
        out = torch.empty((r1 - r0,), dtype=torch.int32, device=X.device)
        check(lib().mcbb_dbscan_rows_count_dev(_ptr(X), X.shape[1], len(gl), gl.ctypes.data_as(abi.c_int32_p),
                                               gw.ctypes.data_as(abi.c_double_p), float(eps), int(minpts), r0, r1,
                                               _ptr(out), _stream()))
        return out

    def rows_count_lists(self, X, gl, gw, eps, minpts, r0, r1):
        import torch
        out = torch.empty((r1 - r0,), dtype=torch.int32, device=X.device)
        nbr = torch.zeros((r1 - r0, minpts), dtype=torch.int32, device=X.device)
        nfill = torch.empty((r1 - r0,), dtype=torch.int32, device=X.device)
        check(lib().mcbb_dbscan_rows_count_lists_dev(_ptr(X), X.shape[1], len(gl), gl.ctypes.data_as(abi.c_int32_p),
                                                     gw.ctypes.data_as(abi.c_double_p), float(eps), int(minpts), r0, r1,
                                                     _ptr(out), _ptr(nbr), _ptr(nfill), _stream()))
        return out, nbr, nfill

    def gather(self, X, idx):
        import torch
        out = torch.empty((X.shape[0], idx.numel()), dtype=torch.float64, device=X.device)
        check(lib().mcbb_gather_columns_dev(_ptr(X), X.shape[1], _ptr(idx), idx.numel(), X.shape[0], _ptr(out),
                                            _stream()))
        return out

    def rows_union(self, Xc, gl, gw, eps, c0, c1):
        import torch
        parent = torch.empty((Xc.shape[1],), dtype=torch.int32, device=Xc.device)
        check(lib().mcbb_dbscan_rows_union_dev(_ptr(Xc), Xc.shape[1], len(gl), gl.ctypes.data_as(abi.c_int32_p),
                                               gw.ctypes.data_as(abi.c_double_p), float(eps), c0, c1, _ptr(parent),
                                               _stream()))
        return parent

    def rows_union_more(self, Xc, gl, gw, eps, c0, c1, parent):
        check(lib().mcbb_dbscan_rows_union_more_dev(_ptr(Xc), Xc.shape[1], len(gl), gl.ctypes.data_as(abi.c_int32_p),
                                                    gw.ctypes.data_as(abi.c_double_p), float(eps), c0, c1, _ptr(parent),
                                                    _stream()))

    def merge_parents(self, parent, other):
        check(lib().mcbb_dbscan_merge_parents_dev(_ptr(parent), _ptr(other), parent.numel(), _stream()))

    def rows_border(self, Xn, b0, b1, Xc, core_seed, gl, gw, eps):
        import torch
        out = torch.empty((b1 - b0,), dtype=torch.int32, device=Xn.device)
        check(lib().mcbb_dbscan_rows_border_dev(_ptr(Xn), Xn.shape[1], b0, b1, _ptr(Xc), Xc.shape[1], _ptr(core_seed),
                                                len(gl), gl.ctypes.data_as(abi.c_int32_p),
                                                gw.ctypes.data_as(abi.c_double_p), float(eps), _ptr(out), _stream()))
        return out

    def labels(self, root_label):
        import torch
        n = root_label.numel()
        a = torch.empty((n,), dtype=torch.int64, device=root_label.device)
        s = torch.empty((n,), dtype=torch.int64, device=root_label.device)
        c = torch.empty((n,), dtype=torch.int64, device=root_label.device)
        k = np.zeros(1, dtype=np.int64)
        check(lib().mcbb_dbscan_labels_dev(_ptr(root_label), n, _ptr(a), _ptr(s), _ptr(c),
                                           k.ctypes.data_as(abi.c_int64_p), _stream()))
        return a, s[: k[0]], c[: k[0]]
