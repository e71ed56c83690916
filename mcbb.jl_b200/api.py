"""Host-side mirror of the MCBB.jl interface for the hot path, on top of the C-ABI library.

The reference's host language (Julia) is not available in this image, so this module plays the role of the
Julia glue (mcbb.jl_b200/julia/MCBBB200.jl): same names, argument meaning and error behaviour as

    DEMCBBProblem / solve / DEMCBBSol      src/setup_mc_prob.jl:334-376, 462-472, 1098-1128
    eval_ode_run                           src/eval_mc_prob.jl:75-198
    distance_matrix / DistanceMatrix       src/eval_clustering.jl:32-38, 164-241
    dbscan / DbscanResult                  src/eval_clustering.jl:84, src/sparse_clustering.jl:126-209
    cluster_membership, cluster_n_noise    src/eval_clustering.jl:1154-1265, 1434-1457

Problem setup, sorting and cluster_membership stay host code exactly as in the reference (SURVEY §8(a) A1-A4,
A14); integration+featurization, the distance matrix and DBSCAN go through libmcbb_b200.so.
Indices exposed to the user are 1-based where the reference's are (seeds, state_filter).
"""
import ctypes as C
import math

import numpy as np

from . import _abi as abi
from ._lib import check, lib, MCBBError

# ------------------------------------------------------------------------------------------------------------
# systems (src/systems.jl)
# ------------------------------------------------------------------------------------------------------------


class DEParameters:
    """Abstract parameter type (systems.jl:15)."""
    _system = None
    _fields = ()

    def reconstruct(self, **kw):
        """Parameters.jl `reconstruct(pars; name=value)` (setup_mc_prob.jl:65, 94)."""
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        for k, v in kw.items():
            if k not in self._fields:
                raise KeyError("type %s has no field %s" % (type(self).__name__, k))
            setattr(new, k, v)
        return new


class logistic_parameters(DEParameters):
    _system = abi.SYS_LOGISTIC
    _fields = ("r",)
    _slots = {"r": 0}

    def __init__(self, r):
        self.r = float(r)

    def n_dim(self):
        return 1

    def pack(self, par_names):
        return abi.PackedSystem(self._system, 1, 1, [self.r], [self._slots[p] for p in par_names])


class kuramoto_parameters(DEParameters):
    _system = abi.SYS_KURAMOTO
    _fields = ("K", "w", "N")
    _slots = {"K": 0}

    def __init__(self, K, w, N):
        self.K, self.w, self.N = float(K), np.asarray(w, dtype=np.float64), int(N)

    def n_dim(self):
        return self.N

    def pack(self, par_names):
        return abi.PackedSystem(self._system, self.N, self.N, [self.K], [self._slots[p] for p in par_names],
                                vec_a=self.w)


class kuramoto_network_parameters(DEParameters):
    _system = abi.SYS_KURAMOTO_NETWORK
    _fields = ("K", "w", "N", "A")
    _slots = {"K": 0}

    def __init__(self, K, w, N, A):
        self.K, self.w, self.N = float(K), np.asarray(w, dtype=np.float64), int(N)
        self.A = np.asarray(A, dtype=np.float64)
        if self.A.shape != (self.N, self.N) or self.w.shape != (self.N,):
            raise ValueError("kuramoto_network_parameters: A must be N x N and w of length N")

    def n_dim(self):
        return self.N

    def pack(self, par_names):
        return abi.PackedSystem(self._system, self.N, self.N, [self.K], [self._slots[p] for p in par_names],
                                mat_a=self.A, vec_a=self.w)


class second_order_kuramoto_parameters(DEParameters):
    _system = abi.SYS_SECOND_ORDER_KURAMOTO
    _fields = ("systemsize", "damping", "coupling", "incidence", "drive")
    _slots = {"damping": 0, "coupling": 1}

    def __init__(self, systemsize, damping=0.1, coupling=8.0, incidence=None, drive=None):
        self.systemsize = int(systemsize)
        self.damping, self.coupling = float(damping), float(coupling)
        self.incidence = np.asarray(incidence, dtype=np.float64)
        self.drive = np.asarray(drive, dtype=np.float64)
        if self.incidence.shape[0] != self.systemsize:
            raise ValueError("incidence must be systemsize x n_edges")

    def n_dim(self):
        return 2 * self.systemsize

    def pack(self, par_names):
        return abi.PackedSystem(self._system, self.systemsize, 2 * self.systemsize,
                                [self.damping, self.coupling], [self._slots[p] for p in par_names],
                                mat_a=self.incidence, vec_a=self.drive, n_edges=self.incidence.shape[1])


class non_local_kuramoto_ring_parameters(DEParameters):
    _system = abi.SYS_NONLOCAL_KURAMOTO_RING
    _fields = ("N", "fak", "omega_0", "phase_delay", "coupling_function")
    _slots = {"omega_0": 0, "phase_delay": 1}

    def __init__(self, N, omega_0, phase_delay, coupling_function, fak=None):
        self.N = int(N)
        self.fak = (2 * math.pi) / self.N if fak is None else float(fak)
        self.omega_0, self.phase_delay = float(omega_0), float(phase_delay)
        self.coupling_function = coupling_function

    def n_dim(self):
        return self.N

    def pack(self, par_names):
        # the Julia closure cannot cross the ABI: tabulate G(fak*(k-j)) for k-j = -(N-1)..(N-1)
        G = np.array([self.coupling_function(self.fak * d) for d in range(-(self.N - 1), self.N)],
                     dtype=np.float64)
        return abi.PackedSystem(self._system, self.N, self.N, [self.omega_0, self.phase_delay, self.fak],
                                [self._slots[p] for p in par_names], mat_a=G)


class roessler_parameters(DEParameters):
    _system = abi.SYS_ROESSLER_NETWORK
    _fields = ("a", "b", "c", "K", "L", "N")
    _slots = {"K": 0}

    def __init__(self, a, b, c, K, L, N):
        self.a, self.b, self.c = (np.asarray(v, dtype=np.float64) for v in (a, b, c))
        self.K, self.L, self.N = float(K), np.asarray(L, dtype=np.float64), int(N)

    def n_dim(self):
        return 3 * self.N

    def pack(self, par_names):
        return abi.PackedSystem(self._system, self.N, 3 * self.N, [self.K], [self._slots[p] for p in par_names],
                                mat_a=self.L, vec_a=self.a, vec_b=self.b, vec_c=self.c)


class stuart_landau_sathiyadevi_pars(DEParameters):
    """paper/stuartlandau/stuart_landau_sathiyadevi-standard-config.ipynb cell 1."""
    _system = abi.SYS_STUART_LANDAU
    _fields = ("lambda_", "omega", "eps", "N", "P_1", "P_2", "A_1", "A_2")
    _slots = {"eps": 0, "lambda_": 1, "omega": 2}

    def __init__(self, lambda_, omega, eps, N, P_1, P_2, A_1, A_2):
        self.lambda_, self.omega, self.eps = float(lambda_), float(omega), float(eps)
        self.N, self.P_1, self.P_2 = int(N), int(P_1), int(P_2)
        self.A_1 = np.asarray(A_1, dtype=np.float64)
        self.A_2 = np.asarray(A_2, dtype=np.float64)

    def n_dim(self):
        return 2 * self.N

    def pack(self, par_names):
        return abi.PackedSystem(self._system, self.N, 2 * self.N,
                                [self.eps, self.lambda_, self.omega, float(self.P_1), float(self.P_2)],
                                [self._slots[p] for p in par_names], mat_a=self.A_1, mat_b=self.A_2)


def _sysfun(name, system):
    def f(du, u, p, t):  # the RHS itself runs on the device; this symbol only names the system
        raise MCBBError("%s is evaluated inside libmcbb_b200.so; call solve(DEMCBBProblem(...))" % name)

    f.__name__ = name
    f.system = system
    return f


logistic = _sysfun("logistic", abi.SYS_LOGISTIC)
kuramoto = _sysfun("kuramoto", abi.SYS_KURAMOTO)
kuramoto_network = _sysfun("kuramoto_network", abi.SYS_KURAMOTO_NETWORK)
second_order_kuramoto = _sysfun("second_order_kuramoto", abi.SYS_SECOND_ORDER_KURAMOTO)
non_local_kuramoto_ring = _sysfun("non_local_kuramoto_ring", abi.SYS_NONLOCAL_KURAMOTO_RING)
roessler_network = _sysfun("roessler_network", abi.SYS_ROESSLER_NETWORK)
stuart_landau_sathiyadevi = _sysfun("stuart_landau_sathiyadevi!", abi.SYS_STUART_LANDAU)


def order_parameter(u, N):
    """systems.jl:246-248"""
    u = np.asarray(u)
    return 1.0 / N * math.sqrt(np.sum(np.sin(u)) ** 2 + np.sum(np.cos(u)) ** 2)


# ------------------------------------------------------------------------------------------------------------
# DifferentialEquations.jl problem / algorithm stand-ins (only what solve(prob::DEMCBBProblem) reads)
# ------------------------------------------------------------------------------------------------------------
class ODEProblem:
    discrete = False

    def __init__(self, f, u0, tspan, p):
        if p._system != f.system:
            raise MCBBError("parameter struct does not belong to system %s" % f.__name__)
        self.f, self.u0, self.tspan, self.p = f, np.asarray(u0), (float(tspan[0]), float(tspan[1])), p


class DiscreteProblem(ODEProblem):
    discrete = True


class Tsit5:
    alg = abi.ALG_TSIT5


class RK4:
    alg = abi.ALG_RK4


class FunctionMap:
    alg = abi.ALG_FUNCTIONMAP


# ------------------------------------------------------------------------------------------------------------
# eval_ode_run (src/eval_mc_prob.jl:75-198) as a descriptor: the closure cannot cross the ABI
# ------------------------------------------------------------------------------------------------------------
mean = "mean"
std = "std"
empirical_1D_KL_divergence_hist = "kl_hist"
correlation_ecdf = "correlation_ecdf"
correlation_hist = "correlation_hist"  # the raw bin weights, eval_mc_prob.jl:472-481
_MAT = {"correlation_ecdf": abi.MAT_CORR_ECDF, "correlation_hist": abi.MAT_CORR_HIST}
_MEAS = {"mean": abi.MEAS_MEAN, "std": abi.MEAS_STD, "kl_hist": abi.MEAS_KL_HIST, "mean_real": abi.MEAS_MEAN,
         "std_real": abi.MEAS_STD, "kl_real": abi.MEAS_KL_HIST, "mean_imag": abi.MEAS_MEAN,
         "std_imag": abi.MEAS_STD, "kl_imag": abi.MEAS_KL_HIST}


class EvalOdeRun:
    """eval_ode_run(sol, i, state_filter, eval_funcs, matrix_eval_funcs, global_eval_funcs; cyclic_setback=...)"""

    def __init__(self, state_filter=None, eval_funcs=(mean, std, empirical_1D_KL_divergence_hist),
                 matrix_eval_funcs=None, global_eval_funcs=None, failure_handling="None", cyclic_setback=False,
                 flag_past_measures=True, continuation=None):
        if len(eval_funcs) < 1:
            raise MCBBError("No per dimension measures")  # eval_mc_prob.jl:79-81
        if failure_handling not in ("None", "Inf", "Repeat"):
            raise MCBBError("failure_handling symbol not known")  # eval_mc_prob.jl:124-126
        self.state_filter = None if state_filter is None else [int(i) for i in state_filter]
        self.eval_funcs = list(eval_funcs)
        self.matrix_eval_funcs = list(matrix_eval_funcs or [])
        self.global_eval_funcs = list(global_eval_funcs or [])
        self.failure_handling = failure_handling
        self.cyclic_setback = bool(cyclic_setback)
        self.continuation = continuation  # ContinuationResponse or None

    def pack(self):
        meas = [_MEAS[f] for f in self.eval_funcs]
        comp = [1 if f.endswith("_imag") else 0 for f in self.eval_funcs]
        if len(self.matrix_eval_funcs) > 1:
            raise MCBBError("one matrix measure per problem (correlation_ecdf or correlation_hist)")
        if self.matrix_eval_funcs and self.matrix_eval_funcs[0] not in _MAT:
            raise MCBBError("matrix measure not known: %r" % (self.matrix_eval_funcs[0],))
        mat = _MAT[self.matrix_eval_funcs[0]] if self.matrix_eval_funcs else abi.MAT_NONE
        glob = [abi.GLOB_SUM for _ in self.global_eval_funcs]
        return abi.PackedFeat(meas, comp, self.state_filter, self.cyclic_setback, mat, 30, glob)


class ContinuationResponse:
    """The extra arguments of the response-analysis method of eval_ode_run (eval_mc_prob.jl:202-257):
    `eval_ode_run(sol, i, state_filter, eval_funcs, global_eval_funcs, matrix_eval_funcs, par_var, eps, par_bounds,
    distance_matrix_func; N_t=200, alg=nothing, return_pm=true, new_tspan=nothing, kwargs...)`.

    After the main run, three short runs restart from its end state `sol[end]` at p - eps, p, p + eps (clipped to
    par_bounds) over `new_tspan` (default: the first 15 % of the main time span) with rel_transient_time 0.1; the
    distances D[1,2] and D[2,3] between their measures (or their mean when return_pm is false) are appended to the
    measures as scalars.  `distance_matrix_func` is fixed to the default L1 `distance_matrix(sol, prob, weights)`;
    it crosses the ABI as its `weights`.  `solve_kwargs` are the kwargs the reference hands to the inner solve."""

    def __init__(self, eps, par_bounds, weights, N_t=200, return_pm=True, new_tspan=None, alg=None, **solve_kwargs):
        self.eps = float(eps)
        self.par_bounds = (float(par_bounds[0]), float(par_bounds[1]))
        self.weights = [float(w) for w in weights]
        self.N_t = int(N_t)
        self.return_pm = bool(return_pm)
        self.new_tspan = None if new_tspan is None else (float(new_tspan[0]), float(new_tspan[1]))
        self.alg = alg
        self.solve_kwargs = solve_kwargs


eval_ode_run = EvalOdeRun()  # default: mean, std (corrected, given mean), KL-hist — eval_mc_prob.jl:191-198

# ------------------------------------------------------------------------------------------------------------
# DEMCBBProblem (src/setup_mc_prob.jl:334-376, 686-708, 850-969)
# ------------------------------------------------------------------------------------------------------------


class SolverParameterVar:
    """SolverParameterVar(name, new_val) (src/solver_parametervar.jl:5-35): a SOLVER option instead of a system
    parameter is varied over the ensemble — `solve(prob_i, alg; name => par[i], kwargs...)` (:75-90).  `new_val` is
    a generator `()->value` or an array of values; `name` one of abstol, reltol, dt, dtmax, dtmin."""

    def __init__(self, name, new_val):
        if name not in abi.SOLVER_FIELDS:
            raise MCBBError("solver option %r cannot be varied per run on the B200 integrator (known: %s)"
                            % (name, ", ".join(sorted(abi.SOLVER_FIELDS))))
        self.name, self.new_val = name, new_val


class MultiDimParameterVar:
    """MultiDimParameterVar(data) (setup_mc_prob.jl:176-236): several parameters varied simultaneously; `data` is a
    list of `(name, range | generator)` tuples, all ranges or all generators.  The new-parameter function is
    Parameters.reconstruct (the default, :233), i.e. the named fields are overwritten per run."""

    def __init__(self, data):
        data = [tuple(d[:2]) for d in data]
        if len(data) < 1:
            raise MCBBError("MultiDimParameterVar needs at least one parameter")
        kinds = {callable(d[1]) for d in data}
        if len(kinds) != 1:
            raise MCBBError("MultiDimParameterVar: either all parameters are ranges or all are generators")
        if len(data) > abi.MCBB_MAX_PAR:
            raise MCBBError("at most %d parameters can be varied simultaneously" % abi.MCBB_MAX_PAR)
        self.data = data
        self.N = len(data)
        self.is_func = kinds.pop()

    def __len__(self):
        return self.N

    def __getitem__(self, i):
        return self.data[i]


class DEMCBBProblem:
    """DEMCBBProblem(p, ic_gens | ic_ranges, [N_ic,] pars, (name, range | generator), eval_ode_func, tail_frac)
    DEMCBBProblem(p, ic_gens | ic_ranges, [N_ic,] pars, SolverParameterVar(name, ...), eval_ode_func, tail_frac)
    DEMCBBProblem(p, ic_gens | ic_ranges, [N_ic,] pars, MultiDimParameterVar([...]) | [(name, ...), ...], ...)

    * ic_ranges: list of arrays + parameter array      -> cartesian product, IC index fastest (:850-882)
    * ic generator(s), N_ic, parameter generator       -> N_ic random runs                     (:888-923)
    * ic generator(s), N_ic, parameter array           -> identical ICs per parameter value    (:929-969)
    Generators are Python callables `()->value`; a single generator returning an array sets all dimensions.
    """

    def __init__(self, p, ic, *args):
        if len(args) >= 2 and isinstance(args[0], (int, np.integer)) and not isinstance(args[0], bool) \
                and isinstance(args[1], DEParameters):
            N_ic, pars, par_var = int(args[0]), args[1], args[2]
            rest = args[3:]
        else:
            N_ic, pars, par_var = None, args[0], args[1]
            rest = args[2:]
        eval_func = rest[0] if len(rest) > 0 else eval_ode_run
        tail_frac = rest[1] if len(rest) > 1 else 0.9
        self.p = p
        self.pars = pars
        self.eval_ode_func = eval_func
        self.rel_transient_time = float(tail_frac)
        self.solver_par = None
        if isinstance(par_var, SolverParameterVar):  # setup_mc_prob.jl:354-364: the system parameters stay fixed
            name, values = par_var.name, par_var.new_val
            self.solver_par = name
            self.par_names = []
        elif isinstance(par_var, MultiDimParameterVar) or (isinstance(par_var, list) and len(par_var) > 0
                                                          and isinstance(par_var[0], (tuple, list))):
            mv = par_var if isinstance(par_var, MultiDimParameterVar) else MultiDimParameterVar(par_var)
            self.par_names = [d[0] for d in mv.data]
            self.par_var = mv
            self._init_multidim(p, ic, N_ic, pars, mv)
            return
        else:
            name, values = par_var[0], par_var[1]
            self.par_names = [name]
        self.par_var = (name, values)
        n_dim = pars.n_dim() if not np.iscomplexobj(p.u0) else len(p.u0)
        self.ic_gens = None
        if callable(ic) or (isinstance(ic, (list, tuple)) and len(ic) > 0 and callable(ic[0])):
            gens = [ic] if callable(ic) else list(ic)
            self.ic_gens = gens
            if N_ic is None:
                raise MCBBError("random initial conditions need N_ic")
            ics = self._draw_ics(gens, N_ic, n_dim)
            if callable(values):  # random x random, :888-923
                self.ic = ics
                self.par = np.array([[values()] for _ in range(N_ic)], dtype=np.float64)
            else:  # identical ICs per parameter value, :929-969 (IC index fastest)
                vals = np.asarray(list(values), dtype=np.float64)
                self.ic = np.tile(ics, (len(vals), 1))
                self.par = np.repeat(vals, N_ic)[:, None]
        else:  # ranges x range, cartesian product, first IC dimension fastest (:850-882)
            ranges = [np.asarray(list(r)) for r in ic]
            vals = np.asarray(list(values), dtype=np.float64)
            grids = np.meshgrid(*ranges, vals, indexing="ij")
            cols = [g.reshape(-1, order="F") for g in grids]
            self.ic = np.stack(cols[:-1], axis=1)
            self.par = cols[-1][:, None].astype(np.float64)
        self.ic = np.asfortranarray(self.ic)
        self.par = np.asfortranarray(self.par)
        self.N_mc = self.ic.shape[0]
        if self.N_mc == 0:
            raise MCBBError("Zero inital conditions. Either at least one of the ranges has length 0 or an "
                            "overflow occured")

    def _init_multidim(self, p, ic, N_ic, pars, mv):
        """_ic_par_matrix for several varied parameters (setup_mc_prob.jl:850-969)."""
        n_dim = pars.n_dim() if not np.iscomplexobj(p.u0) else len(p.u0)
        self.ic_gens = None
        if callable(ic) or (isinstance(ic, (list, tuple)) and len(ic) > 0 and callable(ic[0])):
            gens = [ic] if callable(ic) else list(ic)
            self.ic_gens = gens
            if N_ic is None:
                raise MCBBError("random initial conditions need N_ic")
            if not mv.is_func:
                if mv.N > 1:  # :931-933
                    raise MCBBError("Random IC and Parameter with Ranges is so far only supported for N_par=1")
                ics = self._draw_ics(gens, N_ic, n_dim)
                vals = np.asarray(list(mv[0][1]), dtype=np.float64)
                self.ic = np.tile(ics, (len(vals), 1))
                self.par = np.repeat(vals, N_ic)[:, None]
            else:  # random x random, :888-923
                self.ic = self._draw_ics(gens, N_ic, n_dim)
                self.par = np.array([[d[1]() for d in mv.data] for _ in range(N_ic)], dtype=np.float64)
        else:  # ranges x ranges: cartesian product, IC dimensions first, first index fastest (:850-882)
            if mv.is_func:
                raise MCBBError("initial condition ranges need parameter ranges, not generators")
            ranges = [np.asarray(list(r)) for r in ic]
            pvals = [np.asarray(list(d[1]), dtype=np.float64) for d in mv.data]
            grids = np.meshgrid(*ranges, *pvals, indexing="ij")
            cols = [g.reshape(-1, order="F") for g in grids]
            self.ic = np.stack(cols[:len(ranges)], axis=1)
            self.par = np.stack(cols[len(ranges):], axis=1).astype(np.float64)
        self.ic = np.asfortranarray(self.ic)
        self.par = np.asfortranarray(self.par)
        self.N_mc = self.ic.shape[0]
        if self.N_mc == 0:
            raise MCBBError("Zero inital conditions. Either at least one of the ranges has length 0 or an "
                            "overflow occured")

    def redraw(self, idx):
        """A sub-problem over runs `idx` with freshly drawn initial conditions (failure_handling=:Repeat)."""
        import copy
        sub = copy.copy(self)
        sub.ic = np.asfortranarray(self._draw_ics(self.ic_gens, len(idx), self.ic.shape[1]))
        sub.par = np.asfortranarray(self.par[idx])
        sub.N_mc = len(idx)
        return sub

    @staticmethod
    def _draw_ics(gens, N_ic, n_dim):
        first = gens[0]()
        if len(gens) == 1 and np.ndim(first) > 0:
            rows = [np.asarray(first)] + [np.asarray(gens[0]()) for _ in range(N_ic - 1)]
            return np.stack(rows, axis=0)
        if n_dim % len(gens) != 0:
            raise MCBBError("Number of initial cond. genators and Number of initial cond. doesn't fit together")
        out = np.zeros((N_ic, n_dim), dtype=np.result_type(type(first), np.float64))
        for i in range(N_ic):
            for d in range(n_dim):
                out[i, d] = first if (i == 0 and d == 0) else gens[d % len(gens)]()
        return out


def parameter(prob, i=None):
    """parameter(prob) — vector for one parameter, matrix otherwise (setup_mc_prob.jl:424-436)."""
    if i is not None:
        return prob.par[:, i - 1]
    return prob.par[:, 0] if prob.par.shape[1] == 1 else prob.par


def tsave_array(prob, N_t, rel_transient_time=0.7):
    """tsave_array (setup_mc_prob.jl:1166-1180) — computed by the library so every caller sees the same grid."""
    alg = abi.ALG_FUNCTIONMAP if prob.discrete else abi.ALG_TSIT5
    o = abi.solver_opts(alg, prob.tspan)
    n = C.c_int32(0)
    check(lib().mcbb_tsave_array(C.byref(o), int(N_t), float(rel_transient_time), abi.c_double_p(), C.byref(n)))
    out = np.zeros(n.value)
    check(lib().mcbb_tsave_array(C.byref(o), int(N_t), float(rel_transient_time),
                                 out.ctypes.data_as(abi.c_double_p), C.byref(n)))
    return out


class DEMCBBSol:
    """DEMCBBSol (setup_mc_prob.jl:462-472).  `sol[i]` is the tuple of measures of run i (0-based here),
    `get_measure(sol, k)` the N_mc x N_dim matrix of measure k (1-based k as in the reference)."""

    def __init__(self, feat, matrix, glob, retcode, nsteps, N_t, solve_command):
        self.feat = feat  # [N_mc, n_filter, n_meas]
        self.matrix = matrix
        self.glob = glob
        self.retcode = retcode
        self.nsteps = nsteps
        self.N_mc, self.N_dim, self.N_meas_dim = feat.shape
        self.N_meas_matrix = 0 if matrix is None else 1
        self.N_meas_global = 0 if glob is None else glob.shape[1]
        self.N_meas = self.N_meas_dim + self.N_meas_matrix + self.N_meas_global
        self.N_t = N_t
        self.solve_command = solve_command

    def __getitem__(self, i):
        out = [self.feat[i, :, m] for m in range(self.N_meas_dim)]
        if self.matrix is not None:
            out.append(self.matrix[i])
        if self.glob is not None:
            out.extend(self.glob[i, g] for g in range(self.glob.shape[1]))
        return tuple(out)

    @property
    def sol(self):
        return [self[i] for i in range(self.N_mc)]


def get_measure(sol, k, state_filter=None):
    """get_measure(sol, k; state_filter) (setup_mc_prob.jl:601-624), k 1-based; `state_filter` (1-based indices into
    the measured dimensions) applies to per-dimension measures only."""
    k -= 1
    if k < sol.N_meas_dim:
        if state_filter is not None:
            return sol.feat[:, np.asarray(list(state_filter), dtype=np.int64) - 1, k]
        return sol.feat[:, :, k]
    k -= sol.N_meas_dim
    if k < sol.N_meas_matrix:
        return sol.matrix
    k -= sol.N_meas_matrix
    return sol.glob[:, k]


def check_inf_nan(sol):
    """check_inf_nan (eval_mc_prob.jl:264-279): indices [run, measure] (1-based) with Inf / NaN entries."""
    res = {"Inf": [], "NaN": []}
    for m in range(sol.N_meas):
        a = np.asarray(get_measure(sol, m + 1)).reshape(sol.N_mc, -1)
        for i in np.nonzero(np.isnan(a).any(axis=1))[0]:
            res["NaN"].append([int(i) + 1, m + 1])
        for i in np.nonzero(np.isinf(a).any(axis=1))[0]:
            res["Inf"].append([int(i) + 1, m + 1])
    return res


def _solver_opts(prob, alg, kwargs):
    a = alg.alg if alg is not None else (abi.ALG_FUNCTIONMAP if prob.p.discrete else abi.ALG_TSIT5)
    known = ("abstol", "reltol", "dt", "dtmin", "dtmax", "maxiters", "qmin", "qmax", "gamma", "beta1", "beta2",
             "qoldinit")
    for k in kwargs:
        if k not in known:
            raise MCBBError("solve: unsupported keyword %r for the B200 integrator" % k)
    return abi.solver_opts(a, prob.p.tspan, **kwargs)


def custom_solve_b200(prob, t_save, alg=None, **kwargs):
    """The `custom_solve(prob, t_save)` plug-in (setup_mc_prob.jl:1102-1104): integrates and featurizes the
    whole ensemble on the GPU and returns the per-run measures in the caller's run order."""
    return _solve_b200(prob, t_save, alg, abi.SAVE_NONE, kwargs)[:5]


def _solve_b200(prob, t_save, alg, save_mode, kwargs):
    """custom_solve_b200 plus the optional sample export of mcbb_solve_features_saved: returns
    (feat, mat, glob, retcode, nsteps, saved) with saved = None, [N_mc, n_dim, n_t] (SAVE_ALL) or the end states
    [N_mc, n_dim] (SAVE_LAST); complex systems come back as complex arrays."""
    names = prob.par_names
    psys = prob.pars.pack(names)
    pfeat = prob.eval_ode_func.pack()
    o = _solver_opts(prob, alg, kwargs)
    ic = prob.ic
    if np.iscomplexobj(ic):  # Matrix{Complex{Float64}} -> interleaved (re, im) per oscillator
        ic = np.stack([ic.real, ic.imag], axis=2).reshape(ic.shape[0], -1)
    ic = np.asfortranarray(ic, dtype=np.float64)
    par = np.asfortranarray(prob.par, dtype=np.float64)
    t_save = np.ascontiguousarray(t_save, dtype=np.float64)
    n_mc = ic.shape[0]
    if ic.shape[1] != psys.n_dim:
        raise MCBBError("initial conditions have %d dimensions, the system has %d" % (ic.shape[1], psys.n_dim))
    nf = pfeat.n_filter(psys)
    feat = np.full((n_mc, nf, pfeat.n_meas), np.nan, order="F")
    mat = np.full((n_mc, pfeat.corr_nbins), np.nan, order="F") if pfeat.matrix_meas else None
    glob = np.full((n_mc, pfeat.n_global), np.nan, order="F") if pfeat.n_global else None
    ret = np.zeros(n_mc, dtype=np.int32)
    nsteps = np.zeros((n_mc, 2), dtype=np.int64, order="F")
    dp = abi.c_double_p
    if getattr(prob, "solver_par", None) is not None:  # the parameter column holds the per-run solver option
        if prob.solver_par in kwargs:
            raise MCBBError("solver option %r is both varied per run and given as a keyword" % prob.solver_par)
        sv = np.ascontiguousarray(prob.par[:, 0], dtype=np.float64)
        check(lib().mcbb_set_solver_parameter_var(abi.SOLVER_FIELDS[prob.solver_par], sv.ctypes.data_as(dp), n_mc))
    saved = None
    if save_mode == abi.SAVE_ALL:
        saved = np.empty((n_mc, len(t_save), psys.n_dim))  # per run: n_dim x n_t column-major
    elif save_mode == abi.SAVE_LAST:
        saved = np.empty((n_mc, psys.n_dim), order="F")
    check(lib().mcbb_solve_features_saved(
        C.byref(psys.desc), C.byref(o), C.byref(pfeat.opts), n_mc, ic.ctypes.data_as(dp), par.ctypes.data_as(dp),
        t_save.ctypes.data_as(dp), len(t_save), feat.ctypes.data_as(dp),
        mat.ctypes.data_as(dp) if mat is not None else dp(), glob.ctypes.data_as(dp) if glob is not None else dp(),
        ret.ctypes.data_as(abi.c_int32_p), nsteps.ctypes.data_as(abi.c_int64_p), int(save_mode),
        saved.ctypes.data_as(dp) if saved is not None else dp()))
    if saved is not None:
        if save_mode == abi.SAVE_ALL:
            saved = np.transpose(saved, (0, 2, 1))  # [run, dim, sample]
        if np.iscomplexobj(prob.ic):
            saved = saved[:, 0::2] + 1j * saved[:, 1::2]
    if prob.eval_ode_func.failure_handling == "Inf":  # eval_mc_prob.jl:97-119
        # A DtLessThanMin / Unstable run becomes Inf only if its solution is EMPTY (it died before the first saved sample)
        # or its last saved state exceeds 1e11 in some dimension; otherwise the reference warns and evaluates what it has.
        # The last saved states of the (few) failed runs come from a second, sample-exporting solve of just those runs.
        bad = np.nonzero((ret == abi.RET_DTLESSTHANMIN) | (ret == abi.RET_UNSTABLE))[0]
        if len(bad):
            smp = np.full((len(bad), len(t_save), ic.shape[1]), np.nan)  # SAVE_ALL: [run][sample][dim]; NaN = never reached
            f2 = np.full((len(bad), nf, pfeat.n_meas), np.nan, order="F")
            m2 = np.full((len(bad), pfeat.corr_nbins), np.nan, order="F") if mat is not None else None
            g2 = np.full((len(bad), pfeat.n_global), np.nan, order="F") if glob is not None else None
            r2 = np.zeros(len(bad), dtype=np.int32)
            n2 = np.zeros((len(bad), 2), dtype=np.int64, order="F")
            ic_b, par_b = np.asfortranarray(ic[bad]), np.asfortranarray(par[bad])
            check(lib().mcbb_solve_features_saved(
                C.byref(psys.desc), C.byref(o), C.byref(pfeat.opts), len(bad), ic_b.ctypes.data_as(dp),
                par_b.ctypes.data_as(dp), t_save.ctypes.data_as(dp), len(t_save), f2.ctypes.data_as(dp),
                m2.ctypes.data_as(dp) if m2 is not None else dp(), g2.ctypes.data_as(dp) if g2 is not None else dp(),
                r2.ctypes.data_as(abi.c_int32_p), n2.ctypes.data_as(abi.c_int64_p), abi.SAVE_ALL, smp.ctypes.data_as(dp)))
            reached = ~np.isnan(smp).all(axis=2)                      # [run][sample]: was this sample saved?
            n_saved = reached.sum(axis=1)
            last = smp[np.arange(len(bad)), np.maximum(n_saved - 1, 0)]  # sol.u[end] of the truncated solution
            with np.errstate(invalid="ignore"):
                empty = n_saved == 0  # length(sol) == 0
                diverged = empty | (np.abs(last) > 1e11).any(axis=1) | np.isinf(last).any(axis=1)
            feat[bad[diverged]] = np.inf
            if mat is not None:
                mat[bad[diverged]] = np.inf
            if glob is not None:
                glob[bad[diverged]] = np.inf
            if (~diverged).any():
                import warnings
                warnings.warn("Failure Handling Warning, DtLessThanMin but Solution not diverging.")  # :117
    return feat, mat, glob, ret, nsteps, saved


def solve(prob, alg=None, N_t=400, parallel_type=None, flag_check_inf_nan=True, custom_solve=None,
          sort_results=True, **kwargs):
    """solve(prob::DEMCBBProblem, alg, N_t, parallel_type; ...) -> DEMCBBSol (setup_mc_prob.jl:1098-1128).

    Integrates N_mc trajectories, keeps only the post-transient statistics and sorts by parameter value.
    """
    t_save = tsave_array(prob.p, N_t, prob.rel_transient_time)
    cont = getattr(prob.eval_ode_func, "continuation", None)
    end = None
    if cont is not None:
        if custom_solve is not None:
            raise MCBBError("the continuation / response analysis needs the end states: not available with a "
                            "user custom_solve")

        def custom_solve(sub, ts, alg=None, **kw):  # keeps the end states of the runs it solved last
            nonlocal last_end
            *out, last_end = _solve_b200(sub, ts, alg, abi.SAVE_LAST, kw)
            return tuple(out)
        last_end = None
    if custom_solve is None:
        custom_solve = custom_solve_b200
    feat, mat, glob, ret, nsteps = custom_solve(prob, t_save, alg=alg, **kwargs)
    if cont is not None:
        end = last_end
    if prob.eval_ode_func.failure_handling == "Repeat":
        # eval_ode_run returns ((), true) for a failed run and the ensemble loop redraws its initial condition,
        # at most 10 times (eval_mc_prob.jl:120-123, setup_mc_prob.jl:790-813)
        for rep in range(10):
            bad = np.nonzero(ret != abi.RET_SUCCESS)[0]
            if len(bad) == 0:
                break
            if prob.ic_gens is None:
                raise MCBBError("Integration failed and repeated with new IC, but the IC are not random, so the "
                                "result will not change")
            sub = prob.redraw(bad)
            f2, m2, g2, r2, n2 = custom_solve(sub, t_save, alg=alg, **kwargs)
            prob.ic[bad] = sub.ic
            feat[bad], ret[bad], nsteps[bad] = f2, r2, n2
            if end is not None:
                end[bad] = last_end
            if mat is not None:
                mat[bad] = m2
            if glob is not None:
                glob[bad] = g2
        else:
            if np.any(ret != abi.RET_SUCCESS):
                raise MCBBError("More than 10 Repeats of a Problem in the Monte Carlo Run, there might me something "
                                "wrong here!")
    if cont is not None:  # (meas_1..., dD...): the responses follow the global measures, eval_mc_prob.jl:253-256
        dD = _continuation_response(prob, end, cont, alg, kwargs)
        glob = dD if glob is None else np.asfortranarray(np.hstack([glob, dD]))
    sol = DEMCBBSol(feat, mat, glob, ret, nsteps, N_t, custom_solve)
    if flag_check_inf_nan:
        inf_nan = check_inf_nan(sol)
        if len(inf_nan["Inf"]) > 0 or len(inf_nan["NaN"]) > 0:
            import warnings
            warnings.warn("Some elements of the solution are Inf or NaN, check the solution with check_inf_nan "
                          "again!")
    if sort_results:
        sort_(sol, prob)
    return sol


def _continuation_response(prob, end, cont, alg, kwargs):
    """The three restarted runs per trajectory of eval_mc_prob.jl:209-244, for all N_mc trajectories in ONE
    ensemble of 3 N_mc runs, and their distances D[1,2], D[2,3] (eval_mc_prob.jl:246-251)."""
    import copy
    n = prob.N_mc
    p0 = prob.par[:, 0]
    par3 = np.stack([p0 - cont.eps, p0, p0 + cont.eps], axis=1)
    par3[par3[:, 0] < cont.par_bounds[0], 0] = cont.par_bounds[0]
    par3[par3[:, 2] > cont.par_bounds[1], 2] = cont.par_bounds[1]
    t0, t1 = prob.p.tspan
    tspan = cont.new_tspan if cont.new_tspan is not None else (t0, t0 + 0.15 * (t1 - t0))
    sub = copy.copy(prob)
    sub.p = copy.copy(prob.p)
    sub.p.tspan = (float(tspan[0]), float(tspan[1]))
    sub.ic = np.asfortranarray(np.repeat(end, 3, axis=0))  # runs 3i, 3i+1, 3i+2 restart from sol_i[end]
    sub.par = np.asfortranarray(par3.reshape(-1, 1))
    sub.N_mc = 3 * n
    sub.rel_transient_time = 0.1  # DEMCBBProblem(mcp, 3, 0.1, ic, par, par_var), :239
    sub.eval_ode_func = copy.copy(prob.eval_ode_func)
    sub.eval_ode_func.continuation = None
    ts = tsave_array(sub.p, cont.N_t, sub.rel_transient_time)
    kw = cont.solve_kwargs if cont.solve_kwargs else kwargs
    f, mt, g, r, ns = custom_solve_b200(sub, ts, alg=cont.alg if cont.alg is not None else alg, **kw)
    sol3 = DEMCBBSol(f, mt, g, r, ns, cont.N_t, custom_solve_b200)
    # the inner solve's sort by parameter keeps (p - eps, p, p + eps) in place: ascending, ties stable
    X, glen, gw = feature_groups(sol3, sub, cont.weights)
    d12, d23 = np.zeros(n), np.zeros(n)
    c = 0
    for L, w in zip(glen, gw):  # distance_matrix: per-group L1 sums, weighted, accumulated in group order
        blk = X[:, c:c + L]
        d12 += w * np.abs(blk[0::3] - blk[1::3]).sum(axis=1)
        d23 += w * np.abs(blk[1::3] - blk[2::3]).sum(axis=1)
        c += L
    if cont.return_pm:
        return np.asfortranarray(np.stack([d12, d23], axis=1))
    return np.asfortranarray(((d12 + d23) / 2.0)[:, None])


class Trajectory:
    """What `sol.solve_command(prob_i)` returns to get_trajectory: the samples at t_save (savestart = false,
    save_everystep = false; setup_mc_prob.jl:1108-1110).  `np.asarray(tr)` is the N_dim x N_t matrix."""

    def __init__(self, t, u, retcode, i_run):
        self.t, self.u, self.retcode, self.i_run = t, u, int(retcode), int(i_run)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.u, dtype=dtype)

    def __getitem__(self, k):
        return self.u[:, k]

    def __len__(self):
        return self.u.shape[1]


def get_trajectory(prob, sol, clusters_or_i, i=None, only_sol=True, alg=None, rng=None, **kwargs):
    """get_trajectory(prob, sol, clusters, i; only_sol) / get_trajectory(prob, sol, i_run; only_sol)
    (eval_clustering.jl:1484-1500): re-solves ONE run and returns its saved trajectory.  With a clustering result a
    random member of cluster `i` (1-based, 1 = the noise "cluster", as in the reference) is picked; `i_run` is
    0-based like `sol[i_run]`.  `rng` (numpy Generator) makes the pick reproducible."""
    import copy
    if i is not None:
        members = np.nonzero(np.asarray(clusters_or_i.assignments) == (int(i) - 1))[0]
        if len(members) == 0:
            raise MCBBError("cluster %d has no members" % int(i))
        i_run = int((rng or np.random.default_rng()).choice(members))
    else:
        i_run = int(clusters_or_i)
    if not 0 <= i_run < prob.N_mc:
        raise MCBBError("run index out of range")
    sub = copy.copy(prob)
    sub.ic = np.asfortranarray(prob.ic[i_run:i_run + 1])
    sub.par = np.asfortranarray(prob.par[i_run:i_run + 1])
    sub.N_mc = 1
    t_save = tsave_array(prob.p, sol.N_t, prob.rel_transient_time)
    *_, ret, _ns, saved = _solve_b200(sub, t_save, alg, abi.SAVE_ALL, kwargs)
    tr = Trajectory(t_save, saved[0], ret[0], i_run)
    return tr if only_sol else (tr, sub, i_run)


def sort_(sol, prob, i=1):
    """sort!(sol, prob, i) (setup_mc_prob.jl:520-536): stable sortperm on parameter i, permutes sol, prob.ic
    and prob.par in place."""
    p = parameter(prob, i)
    key = np.abs(p) if np.iscomplexobj(p) else p
    idx = np.argsort(key, kind="stable")
    prob.ic[:, :] = prob.ic[idx, :]
    prob.par[:, :] = prob.par[idx, :]
    sol.feat = np.asfortranarray(sol.feat[idx])
    if sol.matrix is not None:
        sol.matrix = np.asfortranarray(sol.matrix[idx])
    if sol.glob is not None:
        sol.glob = np.asfortranarray(sol.glob[idx])
    sol.retcode = sol.retcode[idx]
    sol.nsteps = sol.nsteps[idx]
    return idx


def sort_prob_(prob, i=1):
    """sort!(prob, i) (setup_mc_prob.jl:558-573): the problem alone."""
    p = parameter(prob, i)
    idx = np.argsort(np.abs(p) if np.iscomplexobj(p) else p, kind="stable")
    prob.ic[:, :] = prob.ic[idx, :]
    prob.par[:, :] = prob.par[idx, :]
    return idx


def sort(sol, prob, i=1):
    """sort(sol, prob, i) (setup_mc_prob.jl:505-511): sorted COPIES of both."""
    import copy
    sol_c, prob_c = copy.copy(sol), copy.copy(prob)
    prob_c.ic, prob_c.par = prob.ic.copy(order="F"), prob.par.copy(order="F")
    sort_(sol_c, prob_c, i)
    return sol_c, prob_c


def show_results(sol, prob, min_par, max_par, i=1, sorted=False):
    """show_results (setup_mc_prob.jl:583-593): the measure tuples of the runs whose i-th parameter lies in the OPEN
    interval (min_par, max_par), in sorted order."""
    if not sorted:
        sol, prob = sort(sol, prob, i)
    p = parameter(prob, i)
    return [sol[int(r)] for r in np.nonzero((p > min_par) & (p < max_par))[0]]


def normalize(sol, k=None):
    """normalize(sol, k) (setup_mc_prob.jl:635-660): a copy of `sol` whose measures k (1-based; default all) are
    rescaled to [0, 1] by the minimum and range over ALL runs and dimensions of that measure.  Global measures are
    left alone with the reference's warning."""
    import copy
    import warnings
    ks = list(range(1, sol.N_meas + 1)) if k is None else [int(q) for q in k]
    out = copy.copy(sol)
    out.feat = sol.feat.copy(order="F")
    out.matrix = None if sol.matrix is None else sol.matrix.copy(order="F")
    out.glob = None if sol.glob is None else sol.glob.copy(order="F")
    for q in ks:
        if not 1 <= q <= sol.N_meas:
            raise MCBBError("normalize: measure index %d out of range" % q)
        if q > sol.N_meas_dim + sol.N_meas_matrix:
            warnings.warn("Global measures are not normlized.")
            continue
        src = np.asarray(get_measure(sol, q))
        lo, rng = src.min(), src.max() - src.min()
        with np.errstate(invalid="ignore", divide="ignore"):
            if q <= sol.N_meas_dim:
                out.feat[:, :, q - 1] = (src - lo) / rng
            else:
                out.matrix[:, :] = (src - lo) / rng
    return out


# ------------------------------------------------------------------------------------------------------------
# distance_matrix / dbscan (src/eval_clustering.jl:164-241, 84)
# ------------------------------------------------------------------------------------------------------------
def _compute_hist_edges(data, N_dim, k_bin, nbin_default=50, nbin=None, bin_edges=None):
    """_compute_hist_edges (eval_clustering.jl:607-642): common edges of one measure from the
    Freedman-Draconis rule 0.5*k_bin*IQR/N_dim^(1/3); falls back to `nbin_default` equidistant edges over
    [min-0.1min, max+0.1max] when the IQR is 0 or the rule yields more than nbin_default edges.  `nbin` (edges
    (min-bw):bw:(max+bw) with bw = (max-min)/(nbin-1)) or explicit equidistant `bin_edges` switch the rule off."""
    data = np.asarray(data, dtype=np.float64)
    minval, maxval = float(data.min()), float(data.max())
    q75 = q25 = None
    if bin_edges is None and nbin is None:
        q75, q25 = np.percentile(data, [75, 25])  # type-7 quantiles, like StatsBase.iqr
    return _hist_edges_from_stats(minval, maxval, q25, q75, N_dim, k_bin, nbin_default, nbin, bin_edges)


def _quantile_neighbours(count, q):
    """Type-7 quantile q of `count` sorted values = lerp(x[lo], x[lo + 1], gamma): returns (lo, hi, gamma)."""
    v = (count - 1) * q
    lo = int(math.floor(v))
    return lo, min(lo + 1, count - 1), v - lo


def _lerp(a, b, t):
    """numpy's quantile interpolation (numpy/lib/_function_base_impl.py `_lerp`)."""
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def _hist_edges_from_stats(minval, maxval, q25, q75, N_dim, k_bin, nbin_default=50, nbin=None, bin_edges=None):
    """The rule of _compute_hist_edges on the four numbers it needs (also fed from the device's order statistics)."""
    if bin_edges is not None and nbin is not None:
        raise MCBBError("bin_edges and nbin in kwargs. Please choose only one.")
    if nbin is not None:
        if not nbin > 1:
            raise MCBBError("nbin must be larger than 1")
        bw = (maxval - minval) / (nbin - 1)
        return _julia_range(minval - bw, bw, maxval + bw), bw
    if bin_edges is not None:
        edges = np.asarray(list(bin_edges), dtype=np.float64)
        if len(edges) < 2:
            raise MCBBError("bin_edges needs all edges of the histogram (one more element than bins)")
        return edges, float(edges[1] - edges[0])
    bw = 0.5 * k_bin * float(q75 - q25) / N_dim ** (1 / 3.0)
    if bw != 0:
        edges = _julia_range(minval - bw, bw, maxval + bw)
        if len(edges) <= nbin_default:
            return edges, bw
    import warnings
    warnings.warn("Bin width from the IQR is 0 or gives too many bins; the number of bins is set to %d" % nbin_default)
    a, b = np.longdouble(minval - 0.1 * minval), np.longdouble(maxval + 0.1 * maxval)
    edges = np.array([float(a + (b - a) * np.longdouble(i) / np.longdouble(nbin_default - 1))
                      for i in range(nbin_default)])
    return edges, float(edges[1] - edges[0])


def _compute_hist_weights(meas, edges, use_ecdf=True):
    """_compute_hist_weights (eval_clustering.jl:587-598): per-run Float32 histogram of the N_dim values over the
    common edges (closed=:left, out-of-range dropped), then ecdf_hist (:759-762) - Float32 like the reference."""
    meas = np.asarray(meas, dtype=np.float64)
    n_mc = meas.shape[0]
    ne, nb = len(edges), len(edges) - 1
    first, step = edges[0], edges[1] - edges[0]
    with np.errstate(invalid="ignore"):
        f = np.clip((meas - first) / step + 1.0, 1.0, float(ne))  # Base.searchsortedlast on a range
        k = np.rint(np.where(np.isnan(f), 1.0, f)).astype(np.int64)
        k -= (meas < edges[k - 1]).astype(np.int64)
        ok = (k >= 1) & (k <= nb) & ~np.isnan(meas)
    # counts are integers < 2^24: exact in Float32 whichever way they are accumulated
    flat = (np.arange(n_mc, dtype=np.int64)[:, None] * nb + (k - 1))[ok]
    H = np.bincount(flat, minlength=n_mc * nb).reshape(n_mc, nb).astype(np.float32)
    if use_ecdf:
        H = np.cumsum(H, axis=1, dtype=np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            H = H / H[:, -1:]
    return H


_HIST_ROWS_ON_DEVICE = 1 << 18  # values of one measure from which the rows are built by the library (mcbb_hist_rows)


def _hist_rows(meas, edges, use_ecdf=True):
    """_compute_hist_weights for one measure [N_mc, N_dim]: on the device (mcbb_hist_rows, bit-identical) for anything
    but toy sizes, where the H2D / D2H round trip costs more than the few numpy passes."""
    meas = np.asarray(meas, dtype=np.float64)
    if meas.size < _HIST_ROWS_ON_DEVICE:
        return _compute_hist_weights(meas, edges, use_ecdf)
    n_mc, n_dim = meas.shape
    slab = np.asfortranarray(meas)  # [n_dim][n_mc] in memory
    e = np.ascontiguousarray(edges, dtype=np.float64)
    rows = np.empty((n_mc, len(e) - 1), dtype=np.float64, order="F")
    check(lib().mcbb_hist_rows(slab.ctypes.data_as(abi.c_double_p), n_mc, n_dim, None, 0, 0.0, 1.0,
                               e.ctypes.data_as(abi.c_double_p), len(e), 1 if use_ecdf else 0,
                               rows.ctypes.data_as(abi.c_double_p)))
    return rows.astype(np.float32)


class DistanceMatrixHist:
    """DistanceMatrixHist (eval_clustering.jl:62-73)."""

    def __init__(self, data, weights, relative_parameter, hist_edges, bin_width, ecdf, k_bin):
        self.data, self.weights, self.relative_parameter = data, list(weights), relative_parameter
        self.hist_edges, self.bin_width, self.ecdf, self.k_bin = hist_edges, bin_width, ecdf, k_bin

    @property
    def shape(self):
        return self.data.shape


class DistanceMatrix:
    """DistanceMatrix (eval_clustering.jl:32-38): `.data` is the dense N x N Float64 matrix."""

    def __init__(self, data, weights, relative_parameter):
        self.data, self.weights, self.relative_parameter = data, list(weights), relative_parameter

    @property
    def shape(self):
        return self.data.shape

    def __getitem__(self, ij):
        return self.data[ij]


class B200FeatureMatrix:
    """AbstractDistanceMatrix stand-in that holds the feature matrix instead of the N x N distances; dbscan on it
    recomputes distances tile by tile on the device (the reference's large-N answers are distance_matrix_mmap /
    distance_matrix_sparse, eval_clustering.jl:317, 451)."""

    def __init__(self, X, group_len, group_weight, weights, relative_parameter):
        self.X, self.group_len, self.group_weight = X, group_len, group_weight
        self.weights, self.relative_parameter = list(weights), relative_parameter

    @property
    def shape(self):
        return (self.X.shape[0], self.X.shape[0])


def _relative_parameter(prob):
    """_relative_parameter (eval_clustering.jl:568-580): (par - min) / max   (sic: divides by max)."""
    par = np.array(prob.par, dtype=np.float64, copy=True)
    for i in range(par.shape[1]):
        par[:, i] = (par[:, i] - par[:, i].min()) / par[:, i].max()
    return par


def feature_groups(sol, prob, weights, relative_parameter=False):
    """Packs measures and parameters into the [N_mc, F] feature matrix + (group_len, group_weight) the C-ABI
    takes; group order = measure order of distance_matrix (per-dim, matrix, global, then parameters)."""
    n_par = prob.par.shape[1]
    if len(weights) != sol.N_meas + n_par:
        raise MCBBError("Amount of weights not the same as Measures + Parameters")
    cols, glen = [], []
    for m in range(sol.N_meas_dim):
        cols.append(sol.feat[:, :, m])
        glen.append(sol.N_dim)
    if sol.matrix is not None:
        cols.append(sol.matrix)
        glen.append(sol.matrix.shape[1])
    if sol.glob is not None:
        for g in range(sol.glob.shape[1]):
            cols.append(sol.glob[:, g:g + 1])
            glen.append(1)
    par = _relative_parameter(prob) if relative_parameter else prob.par
    for p in range(n_par):
        cols.append(np.asarray(par[:, p:p + 1], dtype=np.float64))
        glen.append(1)
    X = np.asfortranarray(np.concatenate(cols, axis=1), dtype=np.float64)
    return X, np.asarray(glen, dtype=np.int32), np.asarray(weights, dtype=np.float64)


def _assemble_features(sol, prob, weights, relative_parameter, histograms, use_ecdf, k_bin, nbin_default, nbin=None,
                       bin_edges=None, skip_zero_weight=False):
    """Feature matrix X [N_mc, F] + L1 groups (lengths, weights) whose weighted group sums are the reference distance
    (eval_clustering.jl:164-241), for the plain and the histogram variant; also the histogram edges / bin widths.
    skip_zero_weight (the on-the-fly DBSCAN path only): a per-dimension measure of weight 0 whose values are all finite
    contributes exactly 0 to every distance, so its histogram is not built at all (its edges come back as None)."""
    hist_edges, bin_widths = [], []
    if histograms:
        # per-dimension measures are replaced by per-run Float32 ECDF rows over common edges; the distance of two
        # runs is sum|ECDF_i - ECDF_j| * bin_width (:261-268, 746-752) = an L1 group with weight w * bin_width
        n_par = prob.par.shape[1]
        if len(weights) != sol.N_meas + n_par:
            raise MCBBError("Amount of weights not the same as Measures + Parameters")
        if nbin is not None:
            import warnings
            warnings.warn("nbin amount specified, all usual histogram binning function will not be used")
        if bin_edges is not None and len(bin_edges) != sol.N_meas_dim:
            raise MCBBError("bin_edges needs one entry (a range or None) per per-dimension measure")

        def one_measure(mi):
            data = sol.feat[:, :, mi]
            if skip_zero_weight and weights[mi] == 0 and np.isfinite(data).all():
                return None
            import warnings
            with warnings.catch_warnings(record=True) as rec:
                warnings.simplefilter("always")
                edges, bw = _compute_hist_edges(data, sol.N_dim, k_bin, nbin_default, nbin,
                                                None if bin_edges is None else bin_edges[mi])
            return edges, bw, _hist_rows(data, edges, use_ecdf), [str(w.message) for w in rec]
        # the measures are independent and numpy releases the GIL inside its loops: one host thread per measure
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, max(1, sol.N_meas_dim))) as pool:
            per_meas = list(pool.map(one_measure, range(sol.N_meas_dim)))
        cols, glen, gw = [], [], []
        for mi, res in enumerate(per_meas):
            if res is None:
                hist_edges.append(None)
                bin_widths.append(None)
                continue
            edges, bw, H, msgs = res
            for msg in msgs:
                import warnings
                warnings.warn(msg)
            hist_edges.append(edges)
            bin_widths.append(bw)
            cols.append(H.astype(np.float64))
            glen.append(H.shape[1])
            gw.append(weights[mi] * bw)
        k = sol.N_meas_dim
        if sol.matrix is not None:
            cols.append(sol.matrix)
            glen.append(sol.matrix.shape[1])
            gw.append(weights[k])
            k += 1
        if sol.glob is not None:
            for g in range(sol.glob.shape[1]):
                cols.append(sol.glob[:, g:g + 1])
                glen.append(1)
                gw.append(weights[k])
                k += 1
        par = _relative_parameter(prob) if relative_parameter else prob.par
        for p in range(n_par):
            cols.append(np.asarray(par[:, p:p + 1], dtype=np.float64))
            glen.append(1)
            gw.append(weights[k + p])
        X = np.asfortranarray(np.concatenate(cols, axis=1), dtype=np.float64)
        glen, gw = np.asarray(glen, dtype=np.int32), np.asarray(gw, dtype=np.float64)
    else:
        X, glen, gw = feature_groups(sol, prob, weights, relative_parameter)
    return X, glen, gw, hist_edges, bin_widths


def distance_matrix(sol, prob, weights, relative_parameter=False, histograms=False, use_ecdf=True, k_bin=1,
                    nbin_default=50, materialize=True, nbin=None, bin_edges=None):
    """distance_matrix(sol, prob, weights; relative_parameter, histograms, use_ecdf, k_bin, nbin_default) with the
    default L1 distance_func / wasserstein_histogram_distance (eval_clustering.jl:164-241).
    materialize=False returns a B200FeatureMatrix (distances are then recomputed on the fly by dbscan)."""
    X, glen, gw, hist_edges, bin_widths = _assemble_features(sol, prob, weights, relative_parameter, histograms,
                                                             use_ecdf, k_bin, nbin_default, nbin, bin_edges,
                                                             skip_zero_weight=not materialize)
    if not materialize:
        # groups of weight 0 contribute w * d = 0 to every distance as long as their values are finite (0 * NaN would
        # be NaN in the reference, eval_clustering.jl:242-254): leave them out of the on-the-fly tiles
        keep, start = [], 0
        for g, (ln, w) in enumerate(zip(glen, gw)):
            if w != 0 or not np.isfinite(X[:, start:start + ln]).all():
                keep.append((g, start, int(ln)))
            start += int(ln)
        if 0 < len(keep) < len(glen):
            X = np.asfortranarray(np.concatenate([X[:, s0:s0 + ln] for _, s0, ln in keep], axis=1))
            glen, gw = glen[[g for g, _, _ in keep]], gw[[g for g, _, _ in keep]]
        return B200FeatureMatrix(X, glen, gw, weights, relative_parameter)
    n = X.shape[0]
    D = np.zeros((n, n), order="F")
    check(lib().mcbb_distance_matrix(X.ctypes.data_as(abi.c_double_p), n, len(glen),
                                     glen.ctypes.data_as(abi.c_int32_p), gw.ctypes.data_as(abi.c_double_p),
                                     D.ctypes.data_as(abi.c_double_p)))
    if np.isnan(D).any():
        import warnings
        warnings.warn("There are some elements NaN in the distance matrix")
    if np.isinf(D).any():
        import warnings
        warnings.warn("There are some elements Inf in the distance matrix")
    if histograms:
        return DistanceMatrixHist(D, weights, relative_parameter, hist_edges, bin_widths, use_ecdf, k_bin)
    return DistanceMatrix(D, weights, relative_parameter)


class NonzeroSparseMatrix:
    """NonzeroSparseMatrix (src/sparse_clustering.jl:37-110): compressed columns holding d - value for the stored
    entries; every unstored entry reads as `value`.  colptr / rowval are 0-based here (Julia's are 1-based)."""

    def __init__(self, colptr, rowval, nzval, value, n):
        self.colptr, self.rowval, self.nzval, self.value, self.n = colptr, rowval, nzval, value, int(n)

    @property
    def shape(self):
        return (self.n, self.n)

    def __getitem__(self, ij):
        i, j = ij
        lo, hi = self.colptr[j], self.colptr[j + 1]
        k = lo + np.searchsorted(self.rowval[lo:hi], i)
        if k < hi and self.rowval[k] == i:
            return self.nzval[k] + self.value
        return self.value

    def toarray(self):
        D = np.full((self.n, self.n), self.value, dtype=self.nzval.dtype, order="F")
        cols = np.repeat(np.arange(self.n), np.diff(self.colptr))
        D[self.rowval, cols] = self.nzval + self.value
        return D


def distance_matrix_sparse(sol, prob, weights, relative_parameter=False, histograms=False, use_ecdf=True, k_bin=1,
                           nbin_default=50, sparse_threshold=np.inf, el_type=np.float32, nbin=None, bin_edges=None):
    """distance_matrix_sparse (eval_clustering.jl:451-559): only distances below `sparse_threshold` are kept, as a
    NonzeroSparseMatrix.  The columns come straight from the feature matrix (two passes over the pair tiles on the
    GPU), D is never materialised.  Distances are accumulated in Float64 and rounded to `el_type` once (the reference
    accumulates in `el_type`)."""
    if not np.isfinite(sparse_threshold) or not sparse_threshold > 0:
        raise MCBBError("sparse_threshold must be a finite positive number")
    X, glen, gw, hist_edges, bin_widths = _assemble_features(sol, prob, weights, relative_parameter, histograms,
                                                             use_ecdf, k_bin, nbin_default, nbin, bin_edges)
    n = X.shape[0]
    thr = float(el_type(sparse_threshold))
    ip, dp, lp = abi.c_int32_p, abi.c_double_p, abi.c_int64_p
    cnt = np.zeros(n, dtype=np.int32)
    check(lib().mcbb_distance_sparse_count(X.ctypes.data_as(dp), n, len(glen), glen.ctypes.data_as(ip),
                                           gw.ctypes.data_as(dp), thr, cnt.ctypes.data_as(ip)))
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=colptr[1:])
    rowval = np.zeros(int(colptr[-1]), dtype=np.int32)
    nzval = np.zeros(int(colptr[-1]), dtype=np.float64)
    check(lib().mcbb_distance_sparse_fill(X.ctypes.data_as(dp), n, len(glen), glen.ctypes.data_as(ip),
                                          gw.ctypes.data_as(dp), thr, colptr.ctypes.data_as(lp),
                                          rowval.ctypes.data_as(ip), nzval.ctypes.data_as(dp)))
    S = NonzeroSparseMatrix(colptr, rowval, nzval.astype(el_type), el_type(thr), n)
    if histograms:
        return DistanceMatrixHist(S, weights, relative_parameter, hist_edges, bin_widths, use_ecdf, k_bin)
    return DistanceMatrix(S, weights, relative_parameter)


def distance_matrix_mmap(sol, prob, weights, relative_parameter=False, histograms=False, use_ecdf=True, k_bin=1,
                         nbin_default=50, el_type=np.float32, save_name="mmap-distance-matrix.bin", block_rows=None,
                         nbin=None, bin_edges=None):
    """distance_matrix_mmap (eval_clustering.jl:317-406): the matrix goes to `save_name` — two Int64 (N_mc, N_mc)
    followed by the rows in `el_type` — block of rows by block of rows from the GPU, and is returned memory-mapped
    (column-major (m, n) view of the file like Mmap.mmap(f, Matrix{el_type}, (m, n)); D is symmetric)."""
    X, glen, gw, hist_edges, bin_widths = _assemble_features(sol, prob, weights, relative_parameter, histograms,
                                                             use_ecdf, k_bin, nbin_default, nbin, bin_edges)
    n = X.shape[0]
    ip, dp = abi.c_int32_p, abi.c_double_p
    if block_rows is None:
        block_rows = max(1, min(n, (256 << 20) // (8 * n)))
    buf = np.zeros((block_rows, n), dtype=np.float64)
    with open(save_name, "wb") as f:
        np.array([n, n], dtype=np.int64).tofile(f)
        for r0 in range(0, n, block_rows):
            r1 = min(n, r0 + block_rows)
            check(lib().mcbb_distance_matrix_rows(X.ctypes.data_as(dp), n, len(glen), glen.ctypes.data_as(ip),
                                                  gw.ctypes.data_as(dp), r0, r1, buf.ctypes.data_as(dp)))
            blk = buf[: r1 - r0]
            if np.isnan(blk).any():
                import warnings
                warnings.warn("There are some elements NaN in the distance matrix")
            if np.isinf(blk).any():
                import warnings
                warnings.warn("There are some elements Inf in the distance matrix")
            blk.astype(el_type).tofile(f)
    D = reload_mmap_distance_matrix(None, save_name, el_type=el_type).data
    if histograms:
        return DistanceMatrixHist(D, weights, relative_parameter, hist_edges, bin_widths, use_ecdf, k_bin)
    return DistanceMatrix(D, weights, relative_parameter)


def reload_mmap_distance_matrix(old_instance, binary_file, el_type=np.float32):
    """reload_mmap_distance_matrix (eval_clustering.jl:411-433): re-attach a distance-matrix file to an instance."""
    m, n = (int(v) for v in np.fromfile(binary_file, dtype=np.int64, count=2))
    data = np.memmap(binary_file, dtype=el_type, mode="r", offset=16, shape=(m, n), order="F")
    if old_instance is None:
        return DistanceMatrix(data, [], False)
    if isinstance(old_instance, DistanceMatrixHist):
        return DistanceMatrixHist(data, old_instance.weights, old_instance.relative_parameter, old_instance.hist_edges,
                                  old_instance.bin_width, old_instance.ecdf, old_instance.k_bin)
    return DistanceMatrix(data, old_instance.weights, old_instance.relative_parameter)


class DbscanResult:
    """Clustering.DbscanResult: seeds (1-based), assignments (0 = noise), counts."""

    def __init__(self, seeds, assignments, counts):
        self.seeds, self.assignments, self.counts = seeds, assignments, counts


def dbscan(D, eps, minpts):
    """Clustering.dbscan(D, eps, minpts) (eval_clustering.jl:84; algorithm src/sparse_clustering.jl:126-209)."""
    lp = abi.c_int64_p
    S = D.data if isinstance(D, (DistanceMatrix, DistanceMatrixHist)) else D
    if isinstance(S, NonzeroSparseMatrix):  # dbscan(::AbstractNonzeroSparseMatrix, ...), sparse_clustering.jl:126-215
        n = S.n
        a, s, c = (np.zeros(n, dtype=np.int64) for _ in range(3))
        k = np.zeros(1, dtype=np.int64)
        colptr = np.ascontiguousarray(S.colptr, dtype=np.int64)
        rowval = np.ascontiguousarray(S.rowval, dtype=np.int32)
        nzval = np.ascontiguousarray(S.nzval, dtype=np.float64)
        check(lib().mcbb_dbscan_sparse(n, colptr.ctypes.data_as(lp), rowval.ctypes.data_as(abi.c_int32_p),
                                       nzval.ctypes.data_as(abi.c_double_p), float(S.value), float(eps), int(minpts),
                                       a.ctypes.data_as(lp), s.ctypes.data_as(lp), c.ctypes.data_as(lp),
                                       k.ctypes.data_as(lp)))
        return DbscanResult(s[:k[0]].copy(), a, c[:k[0]].copy())
    if isinstance(D, B200FeatureMatrix):
        n = D.X.shape[0]
    else:
        data = D.data if isinstance(D, (DistanceMatrix, DistanceMatrixHist)) else D
        data = np.asfortranarray(data, dtype=np.float64)
        n = data.shape[0]
        if data.shape[1] != n:
            raise MCBBError("D must be a square matrix (%s given)." % (data.shape,))
    a = np.zeros(n, dtype=np.int64)
    s = np.zeros(n, dtype=np.int64)
    c = np.zeros(n, dtype=np.int64)
    k = np.zeros(1, dtype=np.int64)
    if isinstance(D, B200FeatureMatrix):
        check(lib().mcbb_dbscan_features(D.X.ctypes.data_as(abi.c_double_p), n, len(D.group_len),
                                         D.group_len.ctypes.data_as(abi.c_int32_p),
                                         D.group_weight.ctypes.data_as(abi.c_double_p), float(eps), int(minpts),
                                         a.ctypes.data_as(lp), s.ctypes.data_as(lp), c.ctypes.data_as(lp),
                                         k.ctypes.data_as(lp)))
    else:
        check(lib().mcbb_dbscan_dense(data.ctypes.data_as(abi.c_double_p), n, float(eps), int(minpts),
                                      a.ctypes.data_as(lp), s.ctypes.data_as(lp), c.ctypes.data_as(lp),
                                      k.ctypes.data_as(lp)))
    return DbscanResult(s[:k[0]].copy(), a, c[:k[0]].copy())


# ------------------------------------------------------------------------------------------------------------
# cluster_membership (src/eval_clustering.jl:1154-1265, 1434-1457) — host code, as in the reference
# ------------------------------------------------------------------------------------------------------------
def cluster_n_noise(clusters):
    return int(np.count_nonzero(clusters.assignments == 0))


def _jl_rat(x):
    """Base.rat (twiceprecision.jl): continued-fraction ratio with terms bounded by maxintfloat(Float32)."""
    y, a, d, b, c, m = x, 1, 1, 0, 0, 16777216.0
    while abs(y) <= m:
        f = int(y)
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > 16777216:
            return c, d
        if b != 0 and a / b == x:
            break
        if y == 0:
            break
        y = 1.0 / y
    return a, b


def _julia_range(start, step, stop):
    """Julia Base (:)(start, step, stop) for Float64: rational lift first, literal fallback otherwise
    (twiceprecision.jl).  Same restatement as the C library's mcbb_tsave_array uses."""
    between = lambda a, x, b: (a <= x <= b) or (b <= x <= a)
    sn, sd = _jl_rat(step)
    if sd != 0 and sn / sd == step:
        an, ad = _jl_rat(start)
        bn, bd = _jl_rat(stop)
        if ad != 0 and bd != 0 and an / ad == start and bn / bd == stop:
            den = ad // math.gcd(ad, sd) * sd
            if den != 0 and abs(start * den) <= 2 ** 53 and abs(step * den) <= 2 ** 53 and den % ad == 0 \
                    and den % sd == 0:
                start_n, step_n = round(start * den), round(step * den)
                q = den * bn - bd * start_n + step_n * bd
                dv = step_n * bd
                ln = max(0, int(q / dv) if (q < 0) != (dv < 0) else q // dv)  # Julia div truncates toward zero
                if between(start, start + (ln - 1) * step, stop + step / 2) and not between(start, start + ln * step,
                                                                                            stop):
                    return np.array([float(np.longdouble(start_n + i * step_n) / np.longdouble(den))
                                     for i in range(ln)])
    lf = (stop - start) / step
    if lf < 0:
        return np.zeros(0)
    n = int(round(lf)) + 1 if lf != 0 else 1
    if n > 1 and (start < stop < start + (n - 1) * step or start > stop > start + (n - 1) * step):
        n -= 1
    return np.array([float(np.longdouble(start) + np.longdouble(i) * np.longdouble(step)) for i in range(n)])


class ClusterMembershipResult:
    """ClusterMembershipResult (eval_clustering.jl:1268-1340): `par` window grid, `data[..., cluster]` memberships;
    indexable by cluster (1-based like the reference, column 1 = noise), sortable by total membership, summable."""

    def __init__(self, par, data, multidim_flag=False):
        self.par, self.data, self.multidim_flag = par, data, multidim_flag

    @property
    def shape(self):
        return self.data.shape

    def __getitem__(self, i):  # cm[i] / cm[[i, j]]: the selected clusters, window axes kept
        idx = [int(i) - 1] if np.ndim(i) == 0 else [int(q) - 1 for q in i]
        return ClusterMembershipResult(self.par, self.data[..., idx], self.multidim_flag)

    def sort_(self, ignore_first=False):
        """sort!(cm; ignore_first) (:1308-1322): clusters ordered by their summed membership, low to high; with
        ignore_first the noise column stays first."""
        tot = self.data.reshape(-1, self.data.shape[-1]).sum(axis=0)
        if ignore_first:
            order = np.concatenate([[0], np.argsort(tot[1:], kind="stable") + 1])
        else:
            order = np.argsort(tot, kind="stable")
        self.data = self.data[..., order]
        return order

    def sum(self, indices):
        """sum(cm, indices) (:1332-1338): the selected clusters (1-based) added into one column."""
        idx = [int(q) - 1 for q in indices]
        return ClusterMembershipResult(self.par, self.data[..., idx].sum(axis=-1).reshape(-1, 1), self.multidim_flag)


def _cluster_membership_multidim(prob, clusters, window_size, window_offset, normalize, min_members):
    """cluster_membership for several varied parameters (eval_clustering.jl:1206-1265): a window is the open box
    prod_p (ip_p, ip_p + window_size_p); `data` is [N_windows_1, ..., N_windows_P, N_cluster], `par` the mesh of the
    window minima [N_windows_1, ..., N_windows_P, P]."""
    n_par = prob.par.shape[1]
    window_size, window_offset = np.atleast_1d(window_size).astype(float), np.atleast_1d(window_offset).astype(float)
    if len(window_size) != n_par or len(window_offset) != n_par:
        raise MCBBError("Window Size and Window Offset need to have as many elements as they are parameters")
    ca = np.asarray(clusters.assignments)
    n_cluster = len(clusters.seeds) + 1
    mins = []
    for p in range(n_par):
        col = prob.par[:, p]
        mins.append(_julia_range(col.min(), window_offset[p], col.max() - window_size[p]))
        if len(mins[p]) <= 1:
            import warnings
            warnings.warn("Only 1 or less Windows in cluster_membership")
    shape = [len(m) for m in mins]
    data = np.zeros(shape + [n_cluster])
    mesh = np.zeros(shape + [n_par])
    for idx in np.ndindex(*shape):
        ind = np.ones(prob.N_mc, dtype=bool)
        for p in range(n_par):
            lo = mins[p][idx[p]]
            ind &= (prob.par[:, p] > lo) & (prob.par[:, p] < lo + window_size[p])
            mesh[idx + (p,)] = lo
        cnt = np.bincount(ca[ind], minlength=n_cluster).astype(float)
        if normalize:
            with np.errstate(invalid="ignore", divide="ignore"):
                cnt = cnt / np.count_nonzero(ind)
        data[idx] = cnt
    if min_members > 0:
        keep = [0] if cluster_n_noise(clusters) > min_members else []
        keep += [i + 1 for i in range(n_cluster - 1) if clusters.counts[i] > min_members]
        data = data[..., keep]
    return ClusterMembershipResult(mesh, data, True)


def cluster_means(sol, clusters):
    """cluster_means (eval_clustering.jl:770-787): mean of every per-dimension measure per cluster,
    [N_cluster (noise first), N_meas_dim, N_dim]."""
    if len(clusters.counts) == 0:
        import warnings
        warnings.warn("No clusters found")
        return False
    ca = np.asarray(clusters.assignments)
    n_cluster = len(clusters.seeds) + 1
    counts = np.concatenate([[cluster_n_noise(clusters)], np.asarray(clusters.counts)]).astype(np.float64)
    out = np.zeros((n_cluster, sol.N_meas_dim, sol.N_dim))
    with np.errstate(invalid="ignore", divide="ignore"):
        contrib = np.transpose(sol.feat, (0, 2, 1)) / counts[ca][:, None, None]  # summed as x / count, like the reference
    np.add.at(out, ca, contrib)
    return out


def cluster_measure_mean(sol, clusters, i):
    """cluster_measure_mean (eval_clustering.jl:794-802): mean of measure i over all runs AND dimensions of a cluster."""
    meas = np.asarray(get_measure(sol, i), dtype=np.float64).reshape(sol.N_mc, -1)
    ca = np.asarray(clusters.assignments)
    with np.errstate(invalid="ignore"), _quiet():
        return np.array([meas[ca == c].mean() for c in range(len(clusters.seeds) + 1)])


def cluster_measure_std(sol, clusters, i):
    """cluster_measure_std (eval_clustering.jl:809-817): corrected std over all runs and dimensions of a cluster."""
    meas = np.asarray(get_measure(sol, i), dtype=np.float64).reshape(sol.N_mc, -1)
    ca = np.asarray(clusters.assignments)
    with np.errstate(invalid="ignore", divide="ignore"), _quiet():
        return np.array([meas[ca == c].std(ddof=1) for c in range(len(clusters.seeds) + 1)])


class _quiet:
    def __enter__(self):
        import warnings
        self._c = warnings.catch_warnings()
        self._c.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *a):
        return self._c.__exit__(*a)


def get_cluster_measure(sol, i, clusters, i_cluster, prob=None, min_par=-np.inf, max_par=np.inf, i_par=1):
    """get_measure(sol, i, clusters, i_cluster, prob, min_par, max_par, i_par) (eval_clustering.jl:824-858): measure
    i of the runs in cluster i_cluster (1-based, 1 = noise) whose parameter lies in the open interval."""
    ind = np.asarray(clusters.assignments) == (int(i_cluster) - 1)
    if prob is not None:
        par = prob.par[:, i_par - 1] if prob.par.shape[1] > 1 else prob.par[:, 0]
        ind &= (par > min_par) & (par < max_par)
    return np.asarray(get_measure(sol, i))[ind]


class ClusterMeasureResult:
    """ClusterMeasureResult (eval_clustering.jl:928-933)."""

    def __init__(self, par, cluster_measures, cluster_measures_global, multidim_flag):
        self.par, self.cluster_measures = par, cluster_measures
        self.cluster_measures_global, self.multidim_flag = cluster_measures_global, multidim_flag


def cluster_measures(prob, sol, clusters, window_size, window_offset):
    """cluster_measures (eval_clustering.jl:883-906): every per-dimension measure [N_cluster, N_meas_dim, N_dim,
    N_windows...] and every global measure [N_cluster, N_meas_global, N_windows...] on the sliding windows."""
    par, per_dim, glob = None, [], []
    for q in range(1, sol.N_meas_dim + 1):
        par, cm = measure_on_parameter_sliding_window(prob, sol, q, clusters, window_size, window_offset)
        per_dim.append(cm)
    for q in range(sol.N_meas_dim + sol.N_meas_matrix + 1, sol.N_meas + 1):
        par, cm = measure_on_parameter_sliding_window(prob, sol, q, clusters, window_size, window_offset)
        glob.append(cm[:, 0])
    n_cluster = len(clusters.seeds) + 1
    cmeas = np.stack(per_dim, axis=1)
    cglob = np.stack(glob, axis=1) if glob else np.zeros((n_cluster, 0) + cmeas.shape[3:])
    return ClusterMeasureResult(par, cmeas, cglob, prob.par.shape[1] > 1)


def measure_on_parameter_sliding_window(prob, sol, i, *args):
    """measure_on_parameter_sliding_window(prob, sol, i, [clusters,] window_size, window_offset)
    (eval_clustering.jl:1369-1430): the mean of the i-th measure (1-based) over the runs of every cluster inside every
    open parameter window; NaN where a cluster has no member in the window.  Returns (parameter_mesh,
    cluster_meas[N_cluster, N_dim, N_windows...]); without `clusters` all runs form one group."""
    if len(args) == 3:
        clusters, window_size, window_offset = args
        ca, n_cluster = np.asarray(clusters.assignments), len(clusters.seeds) + 1
    elif len(args) == 2:
        window_size, window_offset = args
        ca, n_cluster = np.zeros(prob.N_mc, dtype=np.int64), 1
    else:
        raise MCBBError("measure_on_parameter_sliding_window(prob, sol, i, [clusters,] window_size, window_offset)")
    n_par = prob.par.shape[1]
    window_size, window_offset = np.atleast_1d(window_size).astype(float), np.atleast_1d(window_offset).astype(float)
    if len(window_size) != n_par or len(window_offset) != n_par:
        raise MCBBError("Window Size and Window Offset need to have as many elements as they are parameters")
    mins = [_julia_range(prob.par[:, p].min(), window_offset[p], prob.par[:, p].max() - window_size[p])
            for p in range(n_par)]
    shape = [len(q) for q in mins]
    obs = np.asarray(get_measure(sol, i), dtype=np.float64)
    obs = obs.reshape(len(obs), -1)
    n_dim = obs.shape[1]
    meas = np.zeros([n_cluster, n_dim] + shape)
    mesh = np.zeros([n_par] + shape)
    for idx in np.ndindex(*shape):
        ind = np.ones(prob.N_mc, dtype=bool)
        for p in range(n_par):
            lo = mins[p][idx[p]]
            ind &= (prob.par[:, p] > lo) & (prob.par[:, p] < lo + window_size[p])
            mesh[(p,) + idx] = lo
        cnt = np.bincount(ca[ind], minlength=n_cluster).astype(float)
        acc = np.zeros((n_cluster, n_dim))
        np.add.at(acc, ca[ind], obs[ind])
        with np.errstate(invalid="ignore", divide="ignore"):
            meas[(slice(None), slice(None)) + idx] = np.where(cnt[:, None] > 0, acc / cnt[:, None], np.nan)
    if n_par == 1:
        mesh = mesh[0]
    return mesh, meas


def cluster_membership(prob, clusters, window_size, window_offset, normalize=True, min_members=0):
    """Sliding-window cluster membership (eval_clustering.jl:1206-1265); arrays of window sizes / offsets for several
    varied parameters."""
    if prob.par.shape[1] != 1:
        return _cluster_membership_multidim(prob, clusters, window_size, window_offset, normalize, min_members)
    if np.ndim(window_size) > 0 or np.ndim(window_offset) > 0:
        if np.size(window_size) != 1 or np.size(window_offset) != 1:
            raise MCBBError("Window Size and Window Offset need to have as many elements as they are parameters")
        window_size, window_offset = float(np.ravel(window_size)[0]), float(np.ravel(window_offset)[0])
    ca = np.asarray(clusters.assignments)
    n_cluster = len(clusters.seeds) + 1
    par = prob.par[:, 0]
    mins = _julia_range(par.min(), window_offset, par.max() - window_size)
    if len(mins) <= 1:
        import warnings
        warnings.warn("Only 1 or less Windows in cluster_membership")
    data = np.zeros((len(mins), n_cluster))
    for w, ip in enumerate(mins):
        ind = (par > ip) & (par < ip + window_size)  # strict on both sides, :1222
        data[w] = np.bincount(ca[ind], minlength=n_cluster)
        if normalize:
            with np.errstate(invalid="ignore", divide="ignore"):
                data[w] = data[w] / np.count_nonzero(ind)
    if min_members > 0:
        keep = [0] if cluster_n_noise(clusters) > min_members else []
        keep += [i + 1 for i in range(n_cluster - 1) if clusters.counts[i] > min_members]
        data = data[:, keep]
    return ClusterMembershipResult(mins, data, False)


# ------------------------------------------------------------------------------------------------------------
# eps-selection helpers (src/eval_clustering.jl:1504-1553)
# ------------------------------------------------------------------------------------------------------------
def _knn(D, k, K, block_rows=0):
    if isinstance(D, B200FeatureMatrix):  # D is never materialised: row blocks from the features, sorted per row
        n = D.X.shape[0]
        kth, cum = np.zeros(n), np.zeros(n)
        check(lib().mcbb_knn_dist_features(D.X.ctypes.data_as(abi.c_double_p), n, len(D.group_len),
                                           D.group_len.ctypes.data_as(abi.c_int32_p),
                                           D.group_weight.ctypes.data_as(abi.c_double_p), int(k), int(K),
                                           int(block_rows), kth.ctypes.data_as(abi.c_double_p),
                                           cum.ctypes.data_as(abi.c_double_p)))
        return kth, cum
    data = D.data if isinstance(D, (DistanceMatrix, DistanceMatrixHist)) else D
    data = np.asfortranarray(data, dtype=np.float64)
    n = data.shape[0]
    if data.shape[1] != n:
        raise MCBBError("k_dist: Input Matrix has to be a square matrix")
    kth = np.zeros(n)
    cum = np.zeros(n)
    check(lib().mcbb_knn_dist_dense(data.ctypes.data_as(abi.c_double_p), n, int(k), int(K),
                                    kth.ctypes.data_as(abi.c_double_p), cum.ctypes.data_as(abi.c_double_p)))
    return kth, cum


def k_dist(D, k=4):
    """k_dist(D, k): distance to the k-th nearest neighbour of every point, sorted descending (:1504-1516)."""
    kth, _ = _knn(D, k, 1)
    return np.sort(kth)[::-1]


def KNN_dist(D, K):
    """KNN_dist(D, K): cumulative distance to the K nearest neighbours, N x 1 like sum(k_d, dims=2) (:1536-1553)."""
    _, cum = _knn(D, 1, K)
    return cum[:, None]


def KNN_dist_relative(D, rel_K=0.005):
    """KNN_dist_relative(D, rel_K): K = round(N * rel_K) (:1526-1530)."""
    if isinstance(D, B200FeatureMatrix):
        n = D.X.shape[0]
    else:
        n = (D.data if isinstance(D, (DistanceMatrix, DistanceMatrixHist)) else np.asarray(D)).shape[0]
    return KNN_dist(D, int(round(n * rel_K)))
