"""ctypes mirror of include/mcbb_b200.h (structs, enums, argument packing).

Pure declarations: no library is loaded here, so the oracle wrapper under oracle/ (test infrastructure) can
share the exact same descriptors the product library receives.
"""
import ctypes as C

import numpy as np

MCBB_MAX_MEAS = 8
MCBB_MAX_PAR = 4
MCBB_MAX_GLOBAL = 4
MCBB_N_SCALARS = 8

# enum mcbb_system
SYS_LOGISTIC = 0
SYS_KURAMOTO_NETWORK = 1
SYS_SECOND_ORDER_KURAMOTO = 2
SYS_NONLOCAL_KURAMOTO_RING = 3
SYS_STUART_LANDAU = 4
SYS_ROESSLER_NETWORK = 5
SYS_KURAMOTO = 6

# enum mcbb_alg
ALG_TSIT5 = 0
ALG_RK4 = 1
ALG_FUNCTIONMAP = 2

# enum mcbb_retcode
RET_SUCCESS = 0
RET_MAXITERS = 1
RET_DTLESSTHANMIN = 2
RET_UNSTABLE = 3

# enum mcbb_measure / matrix / global
MEAS_MEAN = 0
MEAS_STD = 1
MEAS_KL_HIST = 2
MAT_NONE = 0
MAT_CORR_ECDF = 1
MAT_CORR_HIST = 2
SAVE_NONE, SAVE_ALL, SAVE_LAST = 0, 1, 2
SOLVER_FIELDS = {"abstol": 1, "reltol": 2, "dt": 3, "dtmax": 4, "dtmin": 5}
GLOB_SUM = 0

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class SystemDesc(C.Structure):
    _fields_ = [
        ("system", C.c_int32),
        ("n_osc", C.c_int32),
        ("n_dim", C.c_int32),
        ("n_edges", C.c_int32),
        ("mat_a", c_double_p),
        ("mat_b", c_double_p),
        ("vec_a", c_double_p),
        ("vec_b", c_double_p),
        ("vec_c", c_double_p),
        ("scalars", C.c_double * MCBB_N_SCALARS),
        ("n_par", C.c_int32),
        ("par_slot", C.c_int32 * MCBB_MAX_PAR),
    ]


class SolverOpts(C.Structure):
    _fields_ = [
        ("alg", C.c_int32),
        ("reserved", C.c_int32),
        ("t0", C.c_double),
        ("t1", C.c_double),
        ("abstol", C.c_double),
        ("reltol", C.c_double),
        ("dt", C.c_double),
        ("dtmin", C.c_double),
        ("dtmax", C.c_double),
        ("maxiters", C.c_int64),
        ("qmin", C.c_double),
        ("qmax", C.c_double),
        ("gamma", C.c_double),
        ("beta1", C.c_double),
        ("beta2", C.c_double),
        ("qoldinit", C.c_double),
    ]


class FeatOpts(C.Structure):
    _fields_ = [
        ("n_meas", C.c_int32),
        ("meas", C.c_int32 * MCBB_MAX_MEAS),
        ("meas_comp", C.c_int32 * MCBB_MAX_MEAS),
        ("n_filter", C.c_int32),
        ("state_filter", c_int32_p),
        ("cyclic_setback", C.c_int32),
        ("matrix_meas", C.c_int32),
        ("corr_nbins", C.c_int32),
        ("n_global", C.c_int32),
        ("global_meas", C.c_int32 * MCBB_MAX_GLOBAL),
        ("kl_bins", C.c_int32),
        ("kl_n_stds", C.c_double),
        ("kl_sig_tol", C.c_double),
    ]


def _f64(a):
    """Contiguous float64 copy in Fortran (Julia) order, or None."""
    if a is None:
        return None
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _ptr(a, typ=c_double_p):
    if a is None:
        return typ()
    return a.ctypes.data_as(typ)


class PackedSystem:
    """A SystemDesc plus the numpy arrays it points into (kept alive together)."""

    def __init__(self, system, n_osc, n_dim, scalars, par_slot, mat_a=None, mat_b=None, vec_a=None,
                 vec_b=None, vec_c=None, n_edges=0):
        self._keep = [_f64(mat_a), _f64(mat_b), _f64(vec_a), _f64(vec_b), _f64(vec_c)]
        d = SystemDesc()
        d.system = system
        d.n_osc = n_osc
        d.n_dim = n_dim
        d.n_edges = n_edges
        d.mat_a, d.mat_b, d.vec_a, d.vec_b, d.vec_c = (_ptr(a) for a in self._keep)
        for i, s in enumerate(list(scalars) + [0.0] * (MCBB_N_SCALARS - len(scalars))):
            d.scalars[i] = float(s)
        if len(par_slot) > MCBB_MAX_PAR:
            raise ValueError("at most %d swept parameters" % MCBB_MAX_PAR)
        d.n_par = len(par_slot)
        for i, s in enumerate(par_slot):
            d.par_slot[i] = int(s)
        self.desc = d
        self.system = system
        self.n_osc = n_osc
        self.n_dim = n_dim


class PackedFeat:
    def __init__(self, meas, meas_comp=None, state_filter=None, cyclic_setback=False, matrix_meas=MAT_NONE,
                 corr_nbins=30, global_meas=(), kl_bins=31, kl_n_stds=3.0, kl_sig_tol=1e-4):
        if not 1 <= len(meas) <= MCBB_MAX_MEAS:
            raise ValueError("No per dimension measures" if len(meas) < 1 else "too many measures")
        f = FeatOpts()
        f.n_meas = len(meas)
        comps = list(meas_comp) if meas_comp is not None else [0] * len(meas)
        for i, m in enumerate(meas):
            f.meas[i] = int(m)
            f.meas_comp[i] = int(comps[i])
        self._filter = None
        if state_filter is not None:
            self._filter = np.ascontiguousarray(np.asarray(state_filter, dtype=np.int32))
            f.n_filter = len(self._filter)
            f.state_filter = self._filter.ctypes.data_as(c_int32_p)
        else:
            f.n_filter = 0
        f.cyclic_setback = int(bool(cyclic_setback))
        f.matrix_meas = int(matrix_meas)
        f.corr_nbins = int(corr_nbins)
        f.n_global = len(global_meas)
        for i, g in enumerate(global_meas):
            f.global_meas[i] = int(g)
        f.kl_bins = int(kl_bins)
        f.kl_n_stds = float(kl_n_stds)
        f.kl_sig_tol = float(kl_sig_tol)
        self.opts = f
        self.n_meas = len(meas)
        self.n_global = len(global_meas)
        self.matrix_meas = int(matrix_meas)
        self.corr_nbins = int(corr_nbins)

    def n_filter(self, packed_sys):
        if self._filter is not None:
            return len(self._filter)
        return packed_sys.n_dim // 2 if packed_sys.system == SYS_STUART_LANDAU else packed_sys.n_dim


def solver_opts(alg, tspan, abstol=0.0, reltol=0.0, dt=0.0, dtmin=0.0, dtmax=0.0, maxiters=0, qmin=0.0,
                qmax=0.0, gamma=0.0, beta1=0.0, beta2=0.0, qoldinit=0.0):
    o = SolverOpts()
    o.alg = int(alg)
    o.t0, o.t1 = float(tspan[0]), float(tspan[1])
    o.abstol, o.reltol, o.dt, o.dtmin, o.dtmax = abstol, reltol, dt, dtmin, dtmax
    o.maxiters = int(maxiters)
    o.qmin, o.qmax, o.gamma, o.beta1, o.beta2, o.qoldinit = qmin, qmax, gamma, beta1, beta2, qoldinit
    return o
