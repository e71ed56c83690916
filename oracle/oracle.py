"""ctypes wrapper of the CPU oracle (oracle/mcbb_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (mcbb.jl_b200/) must never import this module.  PARITY UNPINNED — see the header of
mcbb_oracle.c for what the oracle is and is not checked against.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def _load_abi():
    """The declaration module mcbb.jl_b200/_abi.py (structs / enums only, loads no library), registered under its
    canonical name so that the product package — whichever of the two is imported first — shares the SAME ctypes
    classes (a second copy would make `expected LP_SystemDesc instance` errors depend on the import order)."""
    name = "mcbb_jl_b200._abi"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(_ROOT, "mcbb.jl_b200", "_abi.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


abi = _load_abi()
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "_ref", "libmcbb_oracle.so")
    src = os.path.join(_HERE, "mcbb_oracle.c")
    hdr = os.path.join(_ROOT, "include", "mcbb_b200.h")
    stale = (not os.path.exists(so)) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip, lp = abi.c_double_p, abi.c_int32_p, abi.c_int64_p
        L.orc_tsave_array.argtypes = [C.POINTER(abi.SolverOpts), C.c_int32, C.c_double, dp, ip]
        L.orc_trajectory.argtypes = [C.POINTER(abi.SystemDesc), C.POINTER(abi.SolverOpts), dp, dp, dp,
                                     C.c_int32, dp, lp, ip]
        L.orc_solve_features.argtypes = [C.POINTER(abi.SystemDesc), C.POINTER(abi.SolverOpts),
                                         C.POINTER(abi.FeatOpts), C.c_int64, dp, dp, dp, C.c_int32, dp, dp, dp,
                                         ip, lp, C.c_int32, C.c_int32]
        L.orc_featurize_saved.argtypes = [C.POINTER(abi.SystemDesc), C.POINTER(abi.FeatOpts), C.c_int64, dp,
                                          C.c_int32, dp, dp, dp, C.c_int32]
        L.orc_distance_matrix.argtypes = [dp, C.c_int64, C.c_int32, ip, dp, dp, C.c_int32]
        L.orc_dbscan.argtypes = [dp, C.c_int64, C.c_double, C.c_int32, lp, lp, lp, lp]
        L.orc_cluster_membership.argtypes = [dp, lp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int32,
                                             dp, dp, ip]
        L.orc_kl_hist.argtypes = [dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orc_kl_hist.restype = C.c_double
        L.orc_distance_matrix_hist.argtypes = [dp, C.c_int64, C.c_int32, C.c_int32, dp, C.c_int32, dp, C.c_double, C.c_int32, dp]
        L.orc_hist_edges.argtypes = [dp, C.c_int64, C.c_int32, C.c_double, C.c_int32, dp, dp]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(abi.c_double_p) if a is not None else abi.c_double_p()


def num_threads():
    return int(lib().orc_num_threads())


def tsave_array(opts, n_t, rel_transient_time):
    n = C.c_int32(0)
    lib().orc_tsave_array(C.byref(opts), n_t, rel_transient_time, abi.c_double_p(), C.byref(n))
    out = np.zeros(n.value)
    lib().orc_tsave_array(C.byref(opts), n_t, rel_transient_time, _dp(out), C.byref(n))
    return out


def trajectory(psys, opts, u0, par_row, t_save):
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    par_row = np.ascontiguousarray(np.atleast_1d(par_row), dtype=np.float64)
    t_save = np.ascontiguousarray(t_save, dtype=np.float64)
    saved = np.full((psys.n_dim, len(t_save)), np.nan, order="F")
    ns = np.zeros(2, dtype=np.int64)
    nsv = C.c_int32(0)
    ret = lib().orc_trajectory(C.byref(psys.desc), C.byref(opts), _dp(u0), _dp(par_row), _dp(t_save),
                               len(t_save), _dp(saved), ns.ctypes.data_as(abi.c_int64_p), C.byref(nsv))
    return saved, ret, ns, nsv.value


def solve_features(psys, opts, pfeat, ic, par, t_save, n_threads=0, kl_force_len=0):
    """Returns dict(feat[n_mc, n_filter, n_meas], matrix, glob, retcode, nsteps[n_mc,2])."""
    ic = np.asfortranarray(ic, dtype=np.float64)
    par = np.asfortranarray(np.asarray(par, dtype=np.float64).reshape(ic.shape[0], -1))
    t_save = np.ascontiguousarray(t_save, dtype=np.float64)
    n_mc = ic.shape[0]
    nf = pfeat.n_filter(psys)
    feat = np.full((n_mc, nf, pfeat.n_meas), np.nan, order="F")
    mat = np.full((n_mc, pfeat.corr_nbins), np.nan, order="F") if pfeat.matrix_meas else None
    glob = np.full((n_mc, pfeat.n_global), np.nan, order="F") if pfeat.n_global else None
    ret = np.zeros(n_mc, dtype=np.int32)
    nsteps = np.zeros((n_mc, 2), dtype=np.int64, order="F")
    st = lib().orc_solve_features(C.byref(psys.desc), C.byref(opts), C.byref(pfeat.opts), n_mc, _dp(ic),
                                  _dp(par), _dp(t_save), len(t_save), _dp(feat), _dp(mat), _dp(glob),
                                  ret.ctypes.data_as(abi.c_int32_p), nsteps.ctypes.data_as(abi.c_int64_p),
                                  int(n_threads), int(kl_force_len))
    if st != 0:
        raise RuntimeError("oracle solve_features failed")
    return dict(feat=feat, matrix=mat, glob=glob, retcode=ret, nsteps=nsteps)


def featurize_saved(psys, pfeat, saved_all, kl_force_len=0):
    """saved_all: [n_mc, n_dim, n_t] (C order over runs; each run column-major n_dim x n_t)."""
    n_mc, n_dim, n_t = saved_all.shape
    buf = np.ascontiguousarray(np.transpose(saved_all, (0, 2, 1)), dtype=np.float64)  # run, t, dim
    nf = pfeat.n_filter(psys)
    feat = np.full((n_mc, nf, pfeat.n_meas), np.nan, order="F")
    mat = np.full((n_mc, pfeat.corr_nbins), np.nan, order="F") if pfeat.matrix_meas else None
    glob = np.full((n_mc, pfeat.n_global), np.nan, order="F") if pfeat.n_global else None
    lib().orc_featurize_saved(C.byref(psys.desc), C.byref(pfeat.opts), n_mc, _dp(buf), n_t, _dp(feat), _dp(mat),
                              _dp(glob), int(kl_force_len))
    return dict(feat=feat, matrix=mat, glob=glob)


def distance_matrix(X, group_len, group_weight, n_threads=0):
    X = np.asfortranarray(X, dtype=np.float64)
    n = X.shape[0]
    gl = np.ascontiguousarray(group_len, dtype=np.int32)
    gw = np.ascontiguousarray(group_weight, dtype=np.float64)
    assert gl.sum() == X.shape[1]
    D = np.zeros((n, n), order="F")
    lib().orc_distance_matrix(_dp(X), n, len(gl), gl.ctypes.data_as(abi.c_int32_p), _dp(gw), _dp(D),
                              int(n_threads))
    return D


def dbscan(D, eps, minpts):
    D = np.asfortranarray(D, dtype=np.float64)
    n = D.shape[0]
    a = np.zeros(n, dtype=np.int64)
    s = np.zeros(n, dtype=np.int64)
    c = np.zeros(n, dtype=np.int64)
    k = np.zeros(1, dtype=np.int64)
    lp = abi.c_int64_p
    st = lib().orc_dbscan(_dp(D), n, float(eps), int(minpts), a.ctypes.data_as(lp), s.ctypes.data_as(lp),
                          c.ctypes.data_as(lp), k.ctypes.data_as(lp))
    if st != 0:
        raise ValueError("dbscan: invalid arguments")
    return dict(assignments=a, seeds=s[:k[0]].copy(), counts=c[:k[0]].copy())


def cluster_membership(par, assignments, k, window, offset, normalize=True):
    par = np.ascontiguousarray(par, dtype=np.float64)
    a = np.ascontiguousarray(assignments, dtype=np.int64)
    nw = C.c_int32(0)
    L = lib()
    L.orc_cluster_membership(_dp(par), a.ctypes.data_as(abi.c_int64_p), len(par), int(k), window, offset,
                             int(normalize), abi.c_double_p(), abi.c_double_p(), C.byref(nw))
    out = np.zeros((nw.value, k + 1), order="F")
    mesh = np.zeros(nw.value)
    L.orc_cluster_membership(_dp(par), a.ctypes.data_as(abi.c_int64_p), len(par), int(k), window, offset,
                             int(normalize), _dp(out), _dp(mesh), C.byref(nw))
    return mesh, out


def kl_hist(u, mu, sig, hist_bins=31, n_stds=3.0, sig_tol=1e-4, force_len=0):
    u = np.ascontiguousarray(u, dtype=np.float64)
    return float(lib().orc_kl_hist(_dp(u), len(u), mu, sig, hist_bins, n_stds, sig_tol, force_len))


def distance_matrix_hist(feat, par, weights, k_bin=1.0, nbin_default=50):
    """distance_matrix(...; histograms=true) on per-dimension measures feat[n_mc, n_dim, n_meas] + parameters."""
    feat = np.asfortranarray(feat, dtype=np.float64)
    n_mc, n_dim, n_meas = feat.shape
    par = np.asfortranarray(np.asarray(par, dtype=np.float64).reshape(n_mc, -1))
    w = np.ascontiguousarray(weights, dtype=np.float64)
    D = np.zeros((n_mc, n_mc), order="F")
    lib().orc_distance_matrix_hist(_dp(feat), n_mc, n_dim, n_meas, _dp(par), par.shape[1], _dp(w), float(k_bin),
                                   int(nbin_default), _dp(D))
    return D


def hist_edges(data, n_dim, k_bin=1.0, nbin_default=50):
    data = np.ascontiguousarray(np.asarray(data, dtype=np.float64).ravel())
    edges = np.zeros(nbin_default + 8)
    bw = C.c_double(0)
    ne = lib().orc_hist_edges(_dp(data), len(data), int(n_dim), float(k_bin), int(nbin_default), _dp(edges),
                              C.cast(C.byref(bw), abi.c_double_p))
    return edges[:ne].copy(), bw.value
