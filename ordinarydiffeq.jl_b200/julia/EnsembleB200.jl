# EnsembleB200.jl — the reference-side binding of include/b200ode.h.
#
# Drop-in ensemble algorithm for
#     solve(EnsembleProblem(prob; prob_func), alg, EnsembleB200(); trajectories, saveat, reltol, abstol, …)
# It adds ONE method to SciMLBase.__solve (the dispatch point of
# `solve(::AbstractEnsembleProblem, alg, ensemblealg)`), harvests (u0_i, p_i) from prob_func,
# emits the RHS / Jacobian as C with Symbolics `build_function(...; target = CTarget())`, and
# `ccall`s libb200ode.so.  Nothing below constructs an ODEIntegrator.
#
# NOTE: this file could not be executed in the build environment (no Julia toolchain there);
# the same ABI is exercised end-to-end by the Python host mirror (ordinarydiffeq.jl_b200/ensemble.py),
# whose structs mirror the ones below field by field.
module EnsembleB200Mod

using SciMLBase, Symbolics, StaticArrays
import SciMLBase: __solve, AbstractEnsembleProblem, EnsembleAlgorithm, EnsembleSolution, build_solution
using OrdinaryDiffEqTsit5: Tsit5
using OrdinaryDiffEqVerner: Vern7
using OrdinaryDiffEqRosenbrock: Rosenbrock23, Rodas5P

export EnsembleB200

const LIB = get(ENV, "B200ODE_LIB", "libb200ode.so")

struct EnsembleB200 <: EnsembleAlgorithm
    device::Int
end
EnsembleB200() = EnsembleB200(0)

# ---- C structs (include/b200ode.h) ----------------------------------------------------------
struct B200Problem
    trajectories::Int64
    u0::Ptr{Cvoid}; u0_shared::Int32
    p::Ptr{Cvoid}; p_shared::Int32
    t0::Float64; tf::Float64
end
struct B200Opts
    reltol::Float64; abstol::Float64; dt::Float64; dtmin::Float64; dtmax::Float64
    maxiters::Int64
    saveat::Ptr{Float64}; nsaveat::Int32
    save_start::Int32; save_end::Int32; flags::Int32; reserved::Int32
    tstops::Ptr{Float64}; ntstops::Int32; reserved2::Int32
end
mutable struct B200Result
    u_final::Ptr{Cvoid}; t_final::Ptr{Float64}; us::Ptr{Cvoid}; ts::Ptr{Float64}
    nsaved::Ptr{Int32}; naccept::Ptr{Int32}; nreject::Ptr{Int32}; nf::Ptr{Int32}
    njacs::Ptr{Int32}; nw::Ptr{Int32}; nsolve::Ptr{Int32}; retcode::Ptr{Int32}
    kernel_ms::Float64; total_ms::Float64
end

struct B200Ragged            # save_everystep rows (b200ode_solve_everystep); buffers are released with b200ode_free
    total_rows::Int64
    row_offsets::Ptr{Int64}; ts::Ptr{Float64}; us::Ptr{Cvoid}
end

# B200ODE_ALG_* of include/b200ode.h
alg_id(::Tsit5) = 1; alg_id(::Vern7) = 2; alg_id(::Rosenbrock23) = 3; alg_id(::Rodas5P) = 4
alg_id(::DP5) = 5; alg_id(::BS3) = 6
alg_id(::Rodas5) = 7; alg_id(::Rodas4) = 8; alg_id(::Rodas42) = 9; alg_id(::Rodas4P) = 10; alg_id(::Rodas4P2) = 11
alg_id(::Vern6) = 12; alg_id(::Vern8) = 13; alg_id(::Vern9) = 14; alg_id(::Rosenbrock32) = 15; alg_id(::Rodas5Pe) = 16
const StiffAlgs = Union{Rosenbrock23, Rosenbrock32, Rodas5P, Rodas5Pe, Rodas5, Rodas4, Rodas42, Rodas4P, Rodas4P2}
const B200Algs = Union{Tsit5, Vern6, Vern7, Vern8, Vern9, DP5, BS3, StiffAlgs}
isstiff(alg) = alg isa StiffAlgs
const RETCODES = (ReturnCode.Default, ReturnCode.Success, ReturnCode.MaxIters, ReturnCode.DtLessThanMin,
                  ReturnCode.Unstable, ReturnCode.DtNaN)

function check(rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:b200ode_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))
    rc == -1 ? throw(ArgumentError(msg)) : error("b200ode ($rc): $msg")
end

# ---- code generation: Symbolics traces f(u,p,t) and emits C ------------------------------------
function c_sources(prob, alg, ::Type{T}) where {T}
    n, np = length(prob.u0), prob.p === nothing ? 0 : length(prob.p)
    @variables u[1:n] p[1:max(np, 1)] t
    us, ps = collect(u), collect(p)
    du = SciMLBase.isinplace(prob) ? (d = similar(us, Num); prob.f.f(d, us, ps, t); d) : collect(prob.f.f(us, ps, t))
    fix(s) = T === Float32 ? replace(s, "double" => "float") : s
    rhs = fix(build_function(du, us, ps, t; target = Symbolics.CTarget(), fname = :diffeqf))
    jac = tgr = nothing
    if isstiff(alg)
        J = Symbolics.jacobian(du, us)
        jac = fix(build_function(vec(J), us, ps, t; target = Symbolics.CTarget(), fname = :diffeqjac))  # column major
        tgr = fix(build_function(Symbolics.derivative.(du, t), us, ps, t; target = Symbolics.CTarget(), fname = :diffeqtgrad))
    end
    return n, np, rhs, jac, tgr
end

const HANDLES = Dict{Int, Ptr{Cvoid}}()
function handle(dev)
    get!(HANDLES, dev) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:b200ode_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), h, dev))
        h[]
    end
end

const ALLOWED = (:trajectories, :batch_size, :saveat, :save_start, :save_end, :save_everystep, :save_idxs, :tstops, :reltol, :abstol,
                 :dt, :dtmin, :dtmax, :maxiters, :adaptive, :dense)

function __solve(eprob::AbstractEnsembleProblem, alg::B200Algs, ens::EnsembleB200;
                 trajectories, batch_size = trajectories, kwargs...)
    prob = eprob.prob
    kw = merge(NamedTuple(prob.kwargs), NamedTuple(kwargs))          # merge_problem_kwargs
    for k in keys(kw)
        k in ALLOWED || throw(ArgumentError("EnsembleB200 does not support the keyword $k"))
    end
    T = eltype(prob.u0)
    T <: Union{Float32, Float64} || throw(ArgumentError("EnsembleB200 supports Float32/Float64 states"))
    t0, tf = Float64.(prob.tspan)
    tf > t0 || throw(ArgumentError("EnsembleB200 integrates forward in time only"))
    saveat = get(kw, :saveat, ())
    grid = saveat isa Number ? collect(Float64, (t0 + abs(saveat)):abs(saveat):tf) :
           sort!(Float64[s for s in saveat if t0 < s <= tf])
    everystep = get(kw, :save_everystep, isempty(grid))            # solve.jl:138
    n, np, rhs, jac, tgr = c_sources(prob, alg, T)
    idxs = get(kw, :save_idxs, nothing)
    idxs isa Integer && (idxs = [idxs])
    w = idxs === nothing ? n : length(idxs)                        # components per saved row
    extra = String[]
    everystep && push!(extra, "-DB200_EVERYSTEP=1")
    idxs === nothing || push!(extra, "-DB200_SAVE_IDXS=" * join(idxs .- 1, ","))   # the C side is 0-based
    tstops = collect(Float64, get(kw, :tstops, ()))
    isempty(tstops) || push!(extra, "-DB200_TSTOPS=1")
    adaptive = get(kw, :adaptive, true)
    adaptive || push!(extra, "-DB200_ADAPTIVE=0")
    adaptive || get(kw, :dt, nothing) !== nothing || !isempty(tstops) ||
        throw(ArgumentError("Fixed timestep methods require a choice of dt or choosing the tstops"))
    extra_opt = isempty(extra) ? C_NULL : join(extra, " ")
    h = handle(ens.device)
    prog = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b200ode_compile, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cstring, Cstring, Cstring, Cstring, Cstring, Cstring, Cstring),
                h, prog, alg_id(alg), T === Float32 ? 1 : 0, n, np, rhs, "diffeqf",
                jac === nothing ? C_NULL : jac, "diffeqjac", tgr === nothing ? C_NULL : tgr, "diffeqtgrad", extra_opt))
    # defaults of solve.jl:141-143,596-599 (the C ABI only sees the expanded grid)
    dflt(tend) = everystep || isempty(saveat) || saveat isa Number || tend in saveat
    ss = something(get(kw, :save_start, nothing), dflt(prob.tspan[1]))
    se = get(kw, :save_end, nothing)
    se === nothing && !dflt(prob.tspan[2]) && (se = false)
    opts = B200Opts(get(kw, :reltol, 0.0), get(kw, :abstol, 0.0), something(get(kw, :dt, nothing), 0.0),
                    get(kw, :dtmin, 0.0), get(kw, :dtmax, 0.0), get(kw, :maxiters, 0),
                    isempty(grid) ? Ptr{Float64}(C_NULL) : pointer(grid), length(grid),
                    ss === nothing ? -1 : Int32(ss), se === nothing ? -1 : Int32(se), 0, 0,
                    isempty(tstops) ? Ptr{Float64}(C_NULL) : pointer(tstops), length(tstops), 0)
    tstart = time()
    u = eprob.u_init === nothing ? [] : eprob.u_init
    converged = false
    for b0 in 1:batch_size:trajectories
        I = b0:min(b0 + batch_size - 1, trajectories)
        N = length(I)
        # harvest prob_func on the host: flat AoS tables (Vector{SVector{n,T}} is already this layout)
        U0 = Matrix{T}(undef, n, N); P = Matrix{T}(undef, max(np, 1), N)
        for (k, i) in enumerate(I)
            q = eprob.prob_func(prob, SciMLBase.EnsembleContext(i, 1, nothing))
            U0[:, k] .= q.u0
            np > 0 && (P[:, k] .= q.p)
        end
        cprob = B200Problem(N, pointer(U0), 0, pointer(P), 0, t0, tf)
        nslots = ccall((:b200ode_nslots, LIB), Cint, (Ref{B200Problem}, Ref{B200Opts}), cprob, opts)
        uf = Matrix{T}(undef, n, N); tfin = Vector{Float64}(undef, N)
        us = Array{T, 3}(undef, w, max(nslots, 1), N); ts = Vector{Float64}(undef, max(nslots, 1))
        cnt = [Vector{Int32}(undef, N) for _ in 1:8]
        res = B200Result(pointer(uf), pointer(tfin), nslots > 0 ? pointer(us) : C_NULL, pointer(ts),
                         pointer.(cnt)..., 0.0, 0.0)
        rag = Ref(B200Ragged(0, C_NULL, C_NULL, C_NULL))
        GC.@preserve U0 P grid tstops uf tfin us ts cnt begin
            if everystep     # ragged rows: trajectory k owns rows offs[k]+1 : offs[k+1]
                check(ccall((:b200ode_solve_everystep, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Cvoid}, Ref{B200Problem}, Ref{B200Opts}, Ref{B200Result}, Ref{B200Ragged}),
                            h, prog[], cprob, opts, res, rag))
            else
                check(ccall((:b200ode_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{B200Problem}, Ref{B200Opts}, Ref{B200Result}),
                            h, prog[], cprob, opts, res))
            end
        end
        nsaved, naccept, nreject, nf, njacs, nw, nsolve, rc = cnt
        offs = everystep ? unsafe_wrap(Array, rag[].row_offsets, N + 1) : Int64[]
        rts = everystep ? unsafe_wrap(Array, rag[].ts, rag[].total_rows) : Float64[]
        rus = everystep ? unsafe_wrap(Array, Ptr{T}(rag[].us), (w, Int(rag[].total_rows))) : Matrix{T}(undef, 0, 0)
        sel(v) = idxs === nothing ? v : v[idxs]
        batch = map(1:N) do k
            tk = everystep ? rts[(offs[k] + 1):offs[k + 1]] : nslots > 0 ? ts[1:nsaved[k]] : [t0, tfin[k]]
            uk = everystep ? [SVector{w, T}(rus[:, r]) for r in (offs[k] + 1):offs[k + 1]] :
                 nslots > 0 ? [SVector{w, T}(us[:, s, k]) for s in 1:nsaved[k]] :
                 [SVector{w, T}(sel(U0[:, k])), SVector{w, T}(sel(uf[:, k]))]
            stats = SciMLBase.DEStats(Int(nf[k]), 0, 0, Int(nw[k]), Int(nsolve[k]), Int(njacs[k]), 0, 0, 0, 0,
                                      Int(naccept[k]), Int(nreject[k]), 0.0)
            sol = build_solution(prob, alg, tk, uk; dense = false, stats, retcode = RETCODES[rc[k] + 1])
            out, rerun = eprob.output_func(sol, SciMLBase.EnsembleContext(I[k], 1, nothing))
            rerun && error("rerun is served by re-submitting the trajectory; see ensemble.py for the loop")
            out
        end
        if everystep
            for ptr in (rag[].row_offsets, rag[].ts, rag[].us)
                ccall((:b200ode_free, LIB), Cvoid, (Ptr{Cvoid},), ptr)
            end
        end
        u, converged = eprob.reduction(u, batch, I)
        converged && break
    end
    return EnsembleSolution(u, time() - tstart, converged)
end

end # module
