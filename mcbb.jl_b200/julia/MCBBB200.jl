# MCBBB200.jl — reference-side binding of libmcbb_b200.so (B200, sm_100a) for the MCBB.jl hot path.
#
# UNTESTED HERE: no `julia` binary exists in the build image or on the GPU box, so this file is the binding a
# maintainer would add; every ccall below is exercised through the identical C-ABI by the Python mirror
# (mcbb.jl_b200/api.py) in tests/.  Written against the reference's pinned environment (Julia 1.3,
# DiffEqBase 6.10.1, Clustering 0.13.3).
#
# Drop-in points (SURVEY §8(b)):
#   1. solve(prob::DEMCBBProblem, ...; custom_solve = MCBBB200.custom_solve_b200)   src/setup_mc_prob.jl:1098-1104
#      NOTE the reference bug on that branch: `solve_i_command = nothing` is passed to the `::Function` field of
#      DEMCBBSol (setup_mc_prob.jl:471, 1104, 1116).  `solve_b200` below repeats lines 1100-1127 with a valid
#      solve_command instead of patching the reference.
#   2. distance_matrix_b200(sol, prob, weights; relative_parameter) -> DistanceMatrix        eval_clustering.jl:164-241
#      feature_matrix_b200(sol, prob, weights)                      -> B200FeatureMatrix <: AbstractDistanceMatrix
#   3. Clustering.dbscan(::B200FeatureMatrix, eps, minpts) / dbscan_b200(D, eps, minpts)     eval_clustering.jl:84
module MCBBB200

using MCBB
using DiffEqBase, Clustering, SparseArrays, Mmap
import Clustering: dbscan

const LIB = get(ENV, "MCBB_B200_LIB", "libmcbb_b200.so")

const MAX_MEAS, MAX_PAR, MAX_GLOBAL, N_SCALARS = 8, 4, 4, 8

# ---- mirrors of the C structs (include/mcbb_b200.h); isbits, passed by reference -----------------------------
struct SystemDesc
    system::Int32; n_osc::Int32; n_dim::Int32; n_edges::Int32
    mat_a::Ptr{Float64}; mat_b::Ptr{Float64}; vec_a::Ptr{Float64}; vec_b::Ptr{Float64}; vec_c::Ptr{Float64}
    scalars::NTuple{N_SCALARS,Float64}
    n_par::Int32
    par_slot::NTuple{MAX_PAR,Int32}
end

struct SolverOpts
    alg::Int32; reserved::Int32
    t0::Float64; t1::Float64; abstol::Float64; reltol::Float64; dt::Float64; dtmin::Float64; dtmax::Float64
    maxiters::Int64
    qmin::Float64; qmax::Float64; gamma::Float64; beta1::Float64; beta2::Float64; qoldinit::Float64
end

struct FeatOpts
    n_meas::Int32
    meas::NTuple{MAX_MEAS,Int32}
    meas_comp::NTuple{MAX_MEAS,Int32}
    n_filter::Int32
    state_filter::Ptr{Int32}
    cyclic_setback::Int32; matrix_meas::Int32; corr_nbins::Int32
    n_global::Int32
    global_meas::NTuple{MAX_GLOBAL,Int32}
    kl_bins::Int32
    kl_n_stds::Float64; kl_sig_tol::Float64
end

const SYS_LOGISTIC, SYS_KURAMOTO_NETWORK, SYS_SECOND_ORDER_KURAMOTO, SYS_NONLOCAL_RING, SYS_STUART_LANDAU,
      SYS_ROESSLER, SYS_KURAMOTO = Int32.(0:6)
const ALG_TSIT5, ALG_RK4, ALG_FUNCTIONMAP = Int32.(0:2)
const MEAS_MEAN, MEAS_STD, MEAS_KL_HIST = Int32.(0:2)

function check(status::Cint)
    status == 0 && return
    buf = Vector{UInt8}(undef, 2048)
    ccall((:mcbb_last_error, LIB), Csize_t, (Ptr{UInt8}, Csize_t), buf, length(buf))
    error(unsafe_string(pointer(buf)))
end

pad(t, n, z) = ntuple(i -> i <= length(t) ? t[i] : z, n)

# ---- system descriptors: the flat form of the @with_kw parameter structs of src/systems.jl --------------------
# Returns (SystemDesc, keep_alive): the arrays the descriptor points into must outlive the ccall.
function system_desc(p::MCBB.kuramoto_network_parameters, names)
    A = Matrix{Float64}(p.A); w = Vector{Float64}(p.w)
    slots = Dict(:K => 0)
    SystemDesc(SYS_KURAMOTO_NETWORK, p.N, p.N, 0, pointer(A), C_NULL, pointer(w), C_NULL, C_NULL,
               pad((Float64(p.K),), N_SCALARS, 0.0), length(names), pad(Int32[slots[n] for n in names], MAX_PAR, Int32(0))), (A, w)
end
function system_desc(p::MCBB.logistic_parameters, names)
    SystemDesc(SYS_LOGISTIC, 1, 1, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, pad((p.r,), N_SCALARS, 0.0),
               length(names), pad(Int32[0 for _ in names], MAX_PAR, Int32(0))), ()
end
function system_desc(p::MCBB.second_order_kuramoto_parameters, names)
    E = Matrix{Float64}(p.incidence); d = Vector{Float64}(p.drive)
    slots = Dict(:damping => 0, :coupling => 1)
    SystemDesc(SYS_SECOND_ORDER_KURAMOTO, p.systemsize, 2p.systemsize, size(E, 2), pointer(E), C_NULL, pointer(d), C_NULL,
               C_NULL, pad((p.damping, p.coupling), N_SCALARS, 0.0), length(names),
               pad(Int32[slots[n] for n in names], MAX_PAR, Int32(0))), (E, d)
end
function system_desc(p::MCBB.non_local_kuramoto_ring_parameters, names)
    # the coupling_function closure cannot cross the ABI: tabulate G(fak*(k-j)), k-j = -(N-1):(N-1)
    G = Float64[p.coupling_function(p.fak * d) for d in -(p.N - 1):(p.N - 1)]
    slots = Dict(:omega_0 => 0, :phase_delay => 1)
    SystemDesc(SYS_NONLOCAL_RING, p.N, p.N, 0, pointer(G), C_NULL, C_NULL, C_NULL, C_NULL,
               pad((Float64(p.omega_0), Float64(p.phase_delay), p.fak), N_SCALARS, 0.0), length(names),
               pad(Int32[slots[n] for n in names], MAX_PAR, Int32(0))), (G,)
end
function system_desc(p::MCBB.roessler_parameters, names)
    L = Matrix{Float64}(p.L)
    SystemDesc(SYS_ROESSLER, p.N, 3p.N, 0, pointer(L), C_NULL, pointer(p.a), pointer(p.b), pointer(p.c),
               pad((p.K,), N_SCALARS, 0.0), length(names), pad(Int32[0 for _ in names], MAX_PAR, Int32(0))), (L,)
end

function system_desc(p::MCBB.kuramoto_parameters, names)   # all-to-all, src/systems.jl:85-94
    w = Vector{Float64}(p.w)
    SystemDesc(SYS_KURAMOTO, p.N, p.N, 0, C_NULL, C_NULL, pointer(w), C_NULL, C_NULL, pad((Float64(p.K),), N_SCALARS, 0.0),
               length(names), pad(Int32[0 for _ in names], MAX_PAR, Int32(0))), (w,)
end

"""
    stuart_landau_desc(p, names)

The Stuart-Landau ring is not part of src/systems.jl: its parameter struct `stuart_landau_sathiyadevi_pars` (fields
`lambda, omega, eps, N, P_1, P_2, A_1, A_2`) lives in the paper notebooks (paper/stuartlandau/stuart_landau_sathiyadevi-
standard-config.ipynb cell 1).  Any struct with those fields works; register it with
`MCBBB200.system_desc(p::MyPars, names) = MCBBB200.stuart_landau_desc(p, names)`.
"""
function stuart_landau_desc(p, names)
    A1 = Matrix{Float64}(p.A_1); A2 = Matrix{Float64}(p.A_2)       # Bool adjacency -> 0.0 / 1.0 (non-zero = true)
    slots = Dict(:eps => 0, :lambda => 1, :omega => 2, :P_1 => 3, :P_2 => 4)
    SystemDesc(SYS_STUART_LANDAU, p.N, 2p.N, 0, pointer(A1), pointer(A2), C_NULL, C_NULL, C_NULL,
               pad((Float64(p.eps), Float64(p.lambda), Float64(p.omega), Float64(p.P_1), Float64(p.P_2)), N_SCALARS, 0.0),
               length(names), pad(Int32[slots[n] for n in names], MAX_PAR, Int32(0))), (A1, A2)
end

"Matrix{Complex{Float64}} (N_mc x N) -> N_mc x 2N Float64 with (re, im) interleaved per oscillator: the `ic` layout of complex systems."
function interleave_complex(ic::AbstractMatrix{<:Complex})
    out = Matrix{Float64}(undef, size(ic, 1), 2size(ic, 2))
    out[:, 1:2:end] .= real.(ic)
    out[:, 2:2:end] .= imag.(ic)
    out
end

# ---- eval_ode_run as data: the closure is opaque, the user states what it computes ---------------------------
"""
    B200Eval(; measures=[:mean, :std, :kl_hist], state_filter=nothing, cyclic_setback=false)

Describes the `eval_ode_run` variant the problem was built with (src/eval_mc_prob.jl:75-198).  The default is the
reference's default `eval_ode_run(sol, i)` (eval_mc_prob.jl:191-198).
"""
Base.@kwdef struct B200Eval
    measures::Vector{Symbol} = [:mean, :std, :kl_hist]   # complex states: :mean_real, :std_imag, :kl_real, ... as well
    state_filter::Union{Nothing,Vector{Int}} = nothing
    cyclic_setback::Bool = false
    correlation_ecdf::Bool = false     # matrix_eval_funcs = [correlation_ecdf] (30 bins)
    correlation_hist::Bool = false     # matrix_eval_funcs = [correlation_hist] (30 bins, raw weights)
    global_sum::Bool = false           # global_eval_funcs = [(x)->sum(x)]
end
const SOLVER_FIELD = Dict(:abstol => Int32(1), :reltol => Int32(2), :dt => Int32(3), :dtmax => Int32(4), :dtmin => Int32(5))
const MEAS_ID = Dict(:mean => MEAS_MEAN, :std => MEAS_STD, :kl_hist => MEAS_KL_HIST,
                     :mean_real => MEAS_MEAN, :std_real => MEAS_STD, :kl_real => MEAS_KL_HIST,
                     :mean_imag => MEAS_MEAN, :std_imag => MEAS_STD, :kl_imag => MEAS_KL_HIST)
# component a measure works on (the notebook's mean_real / std_imag / ... closures): 0 = real part, 1 = imaginary part
meas_comp(m::Symbol) = endswith(String(m), "_imag") ? Int32(1) : Int32(0)

"""
    custom_solve_b200(prob::DEMCBBProblem, t_save; eval=B200Eval(), alg=ALG_TSIT5, kwargs...)

The `custom_solve(prob, t_save)` plug-in of `MCBB.solve` (src/setup_mc_prob.jl:1102-1104): integrates and
featurizes all `prob.N_mc` trajectories on the GPU and returns an `EnsembleSolution` whose `u[i]` is the tuple of
measure vectors `eval_ode_run` would have produced, in the original run order.
"""
function custom_solve_b200(prob::MCBB.DEMCBBProblem, t_save::AbstractVector; eval::B200Eval=B200Eval(),
                           alg::Int32=(prob.p.prob isa DiscreteProblem ? ALG_FUNCTIONMAP : ALG_TSIT5),
                           abstol=0.0, reltol=0.0, dt=0.0, maxiters=0)
    # SolverParameterVar (src/solver_parametervar.jl): the parameter column is a per-run SOLVER option, the system
    # parameters stay fixed -> no swept system parameter, and a one-shot request for the next solve of this thread
    solver_var = prob.par_var isa MCBB.AbstractSolverParameterVar
    names = solver_var ? Symbol[] : [prob.par_var[i].name for i in 1:length(prob.par_var)]
    solver_vals = solver_var ? Vector{Float64}(prob.par[:, 1]) : Float64[]   # the library keeps the POINTER until the solve
    desc, keep = system_desc(prob.p.prob.p, names)
    tspan = prob.p.prob.tspan
    opts = SolverOpts(alg, 0, tspan[1], tspan[2], abstol, reltol, dt, 0.0, 0.0, maxiters, 0, 0, 0, 0, 0, 0)
    filt = eval.state_filter === nothing ? Int32[] : Int32.(eval.state_filter)
    nf = eval.state_filter === nothing ? size(prob.ic, 2) : length(filt)   # complex ic: one column per oscillator
    nm = length(eval.measures)
    eval.correlation_ecdf && eval.correlation_hist && error("one matrix measure per problem")
    has_mat = eval.correlation_ecdf || eval.correlation_hist
    fo = FeatOpts(nm, pad(Int32[MEAS_ID[m] for m in eval.measures], MAX_MEAS, Int32(0)),
                  pad(Int32[meas_comp(m) for m in eval.measures], MAX_MEAS, Int32(0)),
                  length(filt), isempty(filt) ? C_NULL : pointer(filt), eval.cyclic_setback,
                  eval.correlation_ecdf ? 1 : (eval.correlation_hist ? 2 : 0), 30, eval.global_sum ? 1 : 0,
                  pad(Int32[], MAX_GLOBAL, Int32(0)), 31, 3.0, 1e-4)
    mat = has_mat ? Matrix{Float64}(undef, prob.N_mc, 30) : Matrix{Float64}(undef, 0, 0)
    glob = eval.global_sum ? Matrix{Float64}(undef, prob.N_mc, 1) : Matrix{Float64}(undef, 0, 0)
    # N_mc x N_dim, column-major: already struct-of-arrays over runs; Complex{Float64} states as interleaved (re, im)
    ic = eltype(prob.ic) <: Complex ? interleave_complex(prob.ic) : Matrix{Float64}(prob.ic)
    par = Matrix{Float64}(prob.par)
    ts = Vector{Float64}(t_save)
    feat = Array{Float64}(undef, prob.N_mc, nf, nm)
    retcode = Vector{Int32}(undef, prob.N_mc)
    nsteps = Matrix{Int64}(undef, prob.N_mc, 2)
    GC.@preserve keep filt ic par ts feat mat glob retcode nsteps solver_vals begin
        solver_var && check(ccall((:mcbb_set_solver_parameter_var, LIB), Cint, (Int32, Ptr{Float64}, Int64),
                                  SOLVER_FIELD[prob.par_var.name], solver_vals, length(solver_vals)))
        check(ccall((:mcbb_solve_features, LIB), Cint,
                    (Ref{SystemDesc}, Ref{SolverOpts}, Ref{FeatOpts}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                     Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
                    desc, opts, fo, prob.N_mc, ic, par, ts, length(ts), feat,
                    has_mat ? pointer(mat) : C_NULL, eval.global_sum ? pointer(glob) : C_NULL,
                    retcode, nsteps))
    end
    # tuple order of eval_ode_run: per-dimension measures, matrix measures, global measures (eval_mc_prob.jl:187)
    u = [(ntuple(m -> feat[i, :, m], nm)..., (has_mat ? (mat[i, :],) : ())...,
          (eval.global_sum ? (glob[i, 1],) : ())...) for i in 1:prob.N_mc]
    DiffEqBase.EnsembleSolution(u, 0.0, true)
end

"`solve` of src/setup_mc_prob.jl:1098-1128 with the GPU back end and a valid `solve_command` (see header)."
function solve_b200(prob::MCBB.DEMCBBProblem, N_t::Int=400; eval::B200Eval=B200Eval(), sort_results=true, kwargs...)
    t_save = collect(MCBB.tsave_array(prob.p.prob, N_t, prob.rel_transient_time))
    sol = custom_solve_b200(prob, t_save; eval=eval, kwargs...)
    solve_i = (prob_i) -> DiffEqBase.solve(prob_i, dense=false, save_everystep=false, saveat=t_save, savestart=false)
    mysol = MCBB.DEMCBBSol(sol, prob.N_mc, N_t, length(sol[1][1]), MCBB.get_measure_dimensions(sol, length(sol[1][1]))..., solve_i)
    inf_nan = MCBB.check_inf_nan(mysol)
    (length(inf_nan["Inf"]) > 0) | (length(inf_nan["NaN"]) > 0) &&
        @warn "Some elements of the solution are Inf or NaN, check the solution with check_inf_nan again!"
    sort_results && MCBB.sort!(mysol, prob)
    mysol
end

# ---- distance matrix / DBSCAN ----------------------------------------------------------------------------------
"Feature matrix [N_mc x F] + (group_len, group_weight) in distance_matrix's measure order (eval_clustering.jl:205-224)."
function feature_groups(sol::MCBB.myMCSol, prob::MCBB.myMCProblem, weights::AbstractVector; relative_parameter::Bool=false)
    cols = Matrix{Float64}[]; glen = Int32[]
    for k in 1:sol.N_meas
        M = MCBB.get_measure(sol, k)
        M = ndims(M) == 1 ? reshape(Vector{Float64}(M), :, 1) : Matrix{Float64}(M)
        push!(cols, M); push!(glen, size(M, 2))
    end
    par = relative_parameter ? MCBB._relative_parameter(prob) : MCBB.parameter(prob)
    par = ndims(par) == 1 ? reshape(Vector{Float64}(par), :, 1) : Matrix{Float64}(par)
    for p in 1:size(par, 2)
        push!(cols, par[:, p:p]); push!(glen, 1)
    end
    length(weights) == length(glen) || error("Amount of weights not the same as Measures + Parameters")
    hcat(cols...), glen, Vector{Float64}(weights)
end

"""
    hist_rows_b200(k, sol, edges; ecdf=true)

`_compute_hist_weights(k, sol, edges, ecdf)` (src/eval_clustering.jl:587-598, 759-762) on the device (`mcbb_hist_rows`):
N_mc x (length(edges) - 1) matrix of the per-run Float32 histograms / ECDFs of measure `k` over the equidistant `edges`.
"""
function hist_rows_b200(k::Int, sol, edges; ecdf::Bool=true)
    vals = Matrix{Float64}(MCBB.get_measure(sol, k))   # N_mc x N_dim, column-major = the [n_dim][n_mc] slab of the C-ABI
    e = collect(Float64, edges)
    rows = Matrix{Float64}(undef, size(vals, 1), length(e) - 1)
    check(ccall((:mcbb_hist_rows, LIB), Cint,
                (Ptr{Float64}, Int64, Int32, Ptr{Int64}, Int32, Float64, Float64, Ptr{Float64}, Int32, Int32, Ptr{Float64}),
                vals, size(vals, 1), size(vals, 2), C_NULL, 0, 0.0, 1.0, e, length(e), ecdf ? 1 : 0, rows))
    rows
end

function distance_matrix_b200(sol, prob, weights; relative_parameter::Bool=false, histograms::Bool=false,
                              k_bin::Number=1, nbin_default::Int=50)
    X, glen, gw = feature_groups(sol, prob, weights; relative_parameter=relative_parameter)
    if histograms  # per-dimension measures -> Float32 ECDF rows over common edges; weight w * bin_width
        cols = Matrix{Float64}[]; glen = Int32[]; gw = Float64[]
        for k in 1:sol.N_meas_dim
            edges, bw = MCBB._compute_hist_edges(k, sol, k_bin; nbin_default=nbin_default)
            push!(cols, hist_rows_b200(k, sol, edges))
            push!(glen, length(edges) - 1); push!(gw, weights[k] * bw)
        end
        par = relative_parameter ? MCBB._relative_parameter(prob) : MCBB.parameter(prob)
        par = ndims(par) == 1 ? reshape(Vector{Float64}(par), :, 1) : Matrix{Float64}(par)
        for p in 1:size(par, 2)
            push!(cols, par[:, p:p]); push!(glen, 1); push!(gw, weights[sol.N_meas + p])
        end
        X = hcat(cols...)
    end
    n = size(X, 1)
    D = Matrix{Float64}(undef, n, n)
    check(ccall((:mcbb_distance_matrix, LIB), Cint, (Ptr{Float64}, Int64, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                X, n, length(glen), glen, gw, D))
    MCBB.DistanceMatrix(D, weights, (x, y) -> sum(abs.(x .- y)), nothing, relative_parameter)
end

"""
    distance_matrix_sparse_b200(sol, prob, weights; sparse_threshold, el_type=Float32, relative_parameter=false)

`distance_matrix_sparse` (src/eval_clustering.jl:451-559) without the dense rows: the library counts the entries below
the threshold per column, Julia forms `colptr`, the library fills `rowval` / `nzval = d - threshold`; the result is the
reference's own `NonzeroSparseMatrix`, so `dbscan(::NonzeroSparseMatrix, ...)` and the cluster analysis run unchanged.
"""
function distance_matrix_sparse_b200(sol, prob, weights; sparse_threshold::Number, el_type=Float32, relative_parameter::Bool=false)
    X, glen, gw = feature_groups(sol, prob, weights; relative_parameter=relative_parameter)
    n = size(X, 1); thr = Float64(el_type(sparse_threshold))
    cnt = Vector{Int32}(undef, n)
    check(ccall((:mcbb_distance_sparse_count, LIB), Cint, (Ptr{Float64}, Int64, Int32, Ptr{Int32}, Ptr{Float64}, Float64, Ptr{Int32}),
                X, n, length(glen), glen, gw, thr, cnt))
    colptr0 = vcat(Int64(0), cumsum(Int64.(cnt)))                      # 0-based for the library
    rowval0 = Vector{Int32}(undef, colptr0[end]); nzval = Vector{Float64}(undef, colptr0[end])
    check(ccall((:mcbb_distance_sparse_fill, LIB), Cint,
                (Ptr{Float64}, Int64, Int32, Ptr{Int32}, Ptr{Float64}, Float64, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}),
                X, n, length(glen), glen, gw, thr, colptr0, rowval0, nzval))
    sp = SparseArrays.SparseMatrixCSC(n, n, colptr0 .+ 1, Int64.(rowval0) .+ 1, el_type.(nzval))
    MCBB.DistanceMatrix(MCBB.NonzeroSparseMatrix(sp, el_type(thr)), weights, (x, y) -> sum(abs.(x .- y)), nothing, relative_parameter)
end

"""
    distance_matrix_mmap_b200(sol, prob, weights; el_type=Float32, save_name="mmap-distance-matrix.bin", block_rows=4096)

`distance_matrix_mmap` (src/eval_clustering.jl:317-406): same file (two `Int64`, then the rows in `el_type`), the rows
come from the device in blocks.
"""
function distance_matrix_mmap_b200(sol, prob, weights; el_type=Float32, save_name="mmap-distance-matrix.bin",
                                   block_rows::Int=4096, relative_parameter::Bool=false)
    X, glen, gw = feature_groups(sol, prob, weights; relative_parameter=relative_parameter)
    n = size(X, 1)
    buf = Matrix{Float64}(undef, n, min(block_rows, n))                # column k = row (r0 + k) of D
    open(save_name, "w+") do f
        write(f, Int64(n)); write(f, Int64(n))
        for r0 in 0:size(buf, 2):(n - 1)
            r1 = min(n, r0 + size(buf, 2))
            check(ccall((:mcbb_distance_matrix_rows, LIB), Cint,
                        (Ptr{Float64}, Int64, Int32, Ptr{Int32}, Ptr{Float64}, Int64, Int64, Ptr{Float64}),
                        X, n, length(glen), glen, gw, r0, r1, buf))
            write(f, el_type.(view(buf, :, 1:(r1 - r0))))
        end
    end
    f = open(save_name); m = read(f, Int); n2 = read(f, Int)
    MCBB.DistanceMatrix(Mmap.mmap(f, Matrix{el_type}, (m, n2)), weights, (x, y) -> sum(abs.(x .- y)), nothing, relative_parameter)
end

"Holds the featurized runs instead of the N x N matrix; `dbscan` recomputes distances tile by tile on the device."
struct B200FeatureMatrix <: MCBB.AbstractDistanceMatrix{Float64}
    X::Matrix{Float64}; group_len::Vector{Int32}; group_weight::Vector{Float64}
    weights::AbstractVector; relative_parameter::Bool
end
Base.size(fm::B200FeatureMatrix) = (size(fm.X, 1), size(fm.X, 1))
feature_matrix_b200(sol, prob, weights; relative_parameter::Bool=false) =
    B200FeatureMatrix(feature_groups(sol, prob, weights; relative_parameter=relative_parameter)..., weights, relative_parameter)

function _dbscan_result(n, f)
    a = Vector{Int64}(undef, n); s = Vector{Int64}(undef, n); c = Vector{Int64}(undef, n); k = Ref{Int64}(0)
    check(f(a, s, c, k))
    Clustering.DbscanResult(s[1:k[]], a, c[1:k[]])
end

function dbscan(fm::B200FeatureMatrix, eps::Real, minpts::Int)
    n = size(fm.X, 1)
    _dbscan_result(n, (a, s, c, k) -> ccall((:mcbb_dbscan_features, LIB), Cint,
        (Ptr{Float64}, Int64, Int32, Ptr{Int32}, Ptr{Float64}, Float64, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}),
        fm.X, n, length(fm.group_len), fm.group_len, fm.group_weight, Float64(eps), Int32(minpts), a, s, c, k))
end

"GPU replacement of `Clustering.dbscan(D.data, eps, minpts)` on a materialised matrix (eval_clustering.jl:84)."
function dbscan_b200(D::AbstractMatrix{Float64}, eps::Real, minpts::Int)
    n = size(D, 1)
    size(D, 2) == n || throw(ArgumentError("D must be a square matrix ($(size(D)) given)."))
    Dd = Matrix{Float64}(D)
    _dbscan_result(n, (a, s, c, k) -> ccall((:mcbb_dbscan_dense, LIB), Cint,
        (Ptr{Float64}, Int64, Float64, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}),
        Dd, n, Float64(eps), Int32(minpts), a, s, c, k))
end
dbscan_b200(D::MCBB.AbstractDistanceMatrix, eps::Real, minpts::Int) = dbscan_b200(D.data, eps, minpts)

# ---- eps selection: k_dist / KNN_dist (src/eval_clustering.jl:1504-1553) ----------------------------------------
function _knn(D::AbstractMatrix{Float64}, k::Int, K::Int)
    n = size(D, 1)
    kth = Vector{Float64}(undef, n); cum = Vector{Float64}(undef, n)
    check(ccall((:mcbb_knn_dist_dense, LIB), Cint, (Ptr{Float64}, Int64, Int32, Int32, Ptr{Float64}, Ptr{Float64}),
                Matrix{Float64}(D), n, k, K, kth, cum))
    kth, cum
end
k_dist_b200(D::AbstractMatrix{Float64}, k::Int=4) = sort(_knn(D, k, 1)[1], rev=true)
KNN_dist_b200(D::AbstractMatrix{Float64}, K::Int) = reshape(_knn(D, 1, K)[2], :, 1)
KNN_dist_relative_b200(D::AbstractMatrix{Float64}, rel_K::Float64=0.005) = KNN_dist_b200(D, Int(round(size(D, 1) * rel_K)))

# ---- several GPUs from one Julia process (the reference scales out with addprocs + EnsembleDistributed, docs/src/hpc.md) ---
"Binds `n` devices and creates the NCCL communicators inside the library; afterwards `solve_b200(...; multi=true)`-style calls shard themselves."
multi_init(n::Integer) = check(ccall((:mcbb_multi_init, LIB), Cint, (Int32, Ptr{Int32}), n, C_NULL))
multi_finalize() = check(ccall((:mcbb_multi_finalize, LIB), Cint, ()))

"`dbscan` of a B200FeatureMatrix over all devices of `multi_init` (labels identical to the single-device call)."
function dbscan_multi(fm::B200FeatureMatrix, eps::Real, minpts::Int)
    n = size(fm.X, 1)
    _dbscan_result(n, (a, s, c, k) -> ccall((:mcbb_dbscan_features_multi, LIB), Cint,
        (Ptr{Float64}, Int64, Int32, Ptr{Int32}, Ptr{Float64}, Float64, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}),
        fm.X, n, length(fm.group_len), fm.group_len, fm.group_weight, Float64(eps), Int32(minpts), a, s, c, k))
end
# custom_solve_b200 over all devices: replace :mcbb_solve_features by :mcbb_solve_features_multi in the ccall above (same
# argument list); mcbb_solve_cluster_multi runs solve -> sort -> dbscan without the features leaving the devices and
# returns the sortperm MCBB.sort! would apply (`order_out`) next to the DbscanResult of the sorted runs.

# ---- response analysis (eval_ode_run(sol, i, state_filter, eval_funcs, par_var, eps, par_bounds, distance_matrix_func),
# src/eval_mc_prob.jl:202-257): the three short runs at p - eps, p, p + eps restart from the END STATES of the main solve.
"End states of all runs (`sol[end]` per trajectory) in the `prob.ic` layout: input of the three continuation ensembles."
function end_states_b200(prob::MCBB.DEMCBBProblem, t_save::AbstractVector; eval::B200Eval=B200Eval(), alg::Int32=ALG_TSIT5)
    names = [prob.par_var[i].name for i in 1:length(prob.par_var)]
    desc, keep = system_desc(prob.p.prob.p, names)
    tspan = prob.p.prob.tspan
    opts = SolverOpts(alg, 0, tspan[1], tspan[2], 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    nm = length(eval.measures)
    fo = FeatOpts(nm, pad(Int32[MEAS_ID[m] for m in eval.measures], MAX_MEAS, Int32(0)),
                  pad(Int32[meas_comp(m) for m in eval.measures], MAX_MEAS, Int32(0)), 0, C_NULL, eval.cyclic_setback, 0, 30, 0,
                  pad(Int32[], MAX_GLOBAL, Int32(0)), 31, 3.0, 1e-4)
    ic = eltype(prob.ic) <: Complex ? interleave_complex(prob.ic) : Matrix{Float64}(prob.ic)
    par = Matrix{Float64}(prob.par); ts = Vector{Float64}(t_save)
    feat = Array{Float64}(undef, prob.N_mc, size(prob.ic, 2), nm)
    retcode = Vector{Int32}(undef, prob.N_mc); nsteps = Matrix{Int64}(undef, prob.N_mc, 2)
    last = Matrix{Float64}(undef, prob.N_mc, size(ic, 2))           # MCBB_SAVE_LAST = 2: the ic layout
    GC.@preserve keep ic par ts feat retcode nsteps last begin
        check(ccall((:mcbb_solve_features_saved, LIB), Cint,
                    (Ref{SystemDesc}, Ref{SolverOpts}, Ref{FeatOpts}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Int32, Ptr{Float64}),
                    desc, opts, fo, prob.N_mc, ic, par, ts, length(ts), feat, C_NULL, C_NULL, retcode, nsteps, Int32(2), last))
    end
    last
end
# The response D(p - eps, p) + D(p, p + eps) of every run is then: three `custom_solve_b200` calls on a problem whose `ic`
# is `end_states_b200(...)` and whose parameter column is shifted by -eps, 0, +eps (clamped to par_bounds), followed by
# `feature_groups` + an L1 distance between corresponding runs — what mcbb.jl_b200/api.py::_continuation_response does.

end # module
