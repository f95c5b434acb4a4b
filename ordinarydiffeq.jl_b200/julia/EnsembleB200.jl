# EnsembleB200.jl — the reference-side binding of include/b200ode.h.
#
# Drop-in ensemble algorithm for
#     solve(EnsembleProblem(prob; prob_func), alg, EnsembleB200(); trajectories, saveat, reltol, abstol, …)
# It adds ONE method to SciMLBase.__solve (the dispatch point of
# `solve(::AbstractEnsembleProblem, alg, ensemblealg)`), harvests (u0_i, p_i) from prob_func,
# emits the RHS / Jacobian as C with Symbolics `build_function(...; target = CTarget())`, and
# `ccall`s libb200ode.so.  Nothing below constructs an ODEIntegrator.
#
# STATUS: EXPERIMENTAL — this file has never been executed: no Julia toolchain exists in the build
# container or on the GPU boxes (`which julia` is empty on both).  The same ABI is exercised end to end
# by the Python host mirror (ordinarydiffeq.jl_b200/ensemble.py), whose structs mirror the ones below
# field by field, and by tests/test_abi.py.  Every name used here is imported explicitly below.
module EnsembleB200Mod

using SciMLBase, Symbolics, StaticArrays
import SciMLBase: __solve, AbstractEnsembleProblem, EnsembleAlgorithm, EnsembleSolution, build_solution, ReturnCode
using OrdinaryDiffEqTsit5: Tsit5
using OrdinaryDiffEqVerner: Vern6, Vern7, Vern8, Vern9
using OrdinaryDiffEqLowOrderRK: DP5, BS3
using OrdinaryDiffEqRosenbrock: Rosenbrock23, Rosenbrock32, Rodas5P, Rodas5Pe, Rodas5, Rodas4, Rodas42, Rodas4P, Rodas4P2,
                                Rodas3P, Rodas23W
using OrdinaryDiffEqCore: CompositeAlgorithm, AutoSwitch

export EnsembleB200, B200ContinuousCallback, B200DiscreteCallback

const LIB = get(ENV, "B200ODE_LIB", "libb200ode.so")

"""
    EnsembleB200()                 # device 0
    EnsembleB200(2)                # device 2
    EnsembleB200([0, 1, 2, 3])     # several GPUs from this one process (b200ode_multi_*; chunks dealt round-robin)
    EnsembleB200(:all)             # every visible GPU
"""
struct EnsembleB200 <: EnsembleAlgorithm
    devices::Vector{Int}           # empty = all visible devices
end
EnsembleB200() = EnsembleB200([0])
EnsembleB200(dev::Integer) = EnsembleB200([Int(dev)])
EnsembleB200(s::Symbol) = s === :all ? EnsembleB200(Int[]) : throw(ArgumentError("EnsembleB200(:all) or a device list"))
ismulti(e::EnsembleB200) = length(e.devices) != 1

# ---- C structs (include/b200ode.h) ----------------------------------------------------------
struct B200Problem
    trajectories::Int64
    u0::Ptr{Cvoid}; u0_shared::Int32
    p::Ptr{Cvoid}; p_shared::Int32
    t0::Float64; tf::Float64
    tspans::Ptr{Float64}           # C_NULL, or (t0_i, tf_i) pairs: a prob_func that remakes tspan (B200ODE_OPT_TSPANS programs)
end
struct B200Opts
    reltol::Float64; abstol::Float64; dt::Float64; dtmin::Float64; dtmax::Float64
    maxiters::Int64
    saveat::Ptr{Float64}; nsaveat::Int32
    save_start::Int32; save_end::Int32; flags::Int32; reserved::Int32
    tstops::Ptr{Float64}; ntstops::Int32; reserved2::Int32
    abstol_vec::Ptr{Float64}; reltol_vec::Ptr{Float64}      # per-component tolerances or C_NULL
    d_discontinuities::Ptr{Float64}; nd_discontinuities::Int32; reserved3::Int32
end
struct B200CallbackSrc       # include/b200ode.h
    kind::Int32; rootfind::Int32
    condition_src::Cstring; condition_name::Cstring
    affect_src::Cstring; affect_name::Cstring
    affect_neg_src::Cstring; affect_neg_name::Cstring
    interp_points::Int32; save_before::Int32; save_after::Int32; reserved::Int32
    abstol::Float64; repeat_nudge::Float64
end

"""
Callbacks cross the boundary as C source (the closures of a SciMLBase callback cannot be traced into `affect!` code in
general): `condition` is `real NAME(const real* u, const real* p, const real t)`, `affect` / `affect_neg` are
`void NAME(real* u, real* p, const real t, int* terminate)`, each given as a `(source, name)` pair or `nothing`.
`solve(...; callback = B200ContinuousCallback(...))` or a tuple of them.  Tsit5 only.
"""
struct B200ContinuousCallback
    condition::Tuple{String, String}
    affect::Union{Nothing, Tuple{String, String}}
    affect_neg::Union{Nothing, Tuple{String, String}}
    rootfind::Int                 # 0 NoRootFind, 1 LeftRootFind, 2 RightRootFind
    interp_points::Int
    save_positions::Tuple{Bool, Bool}
    abstol::Float64
    repeat_nudge::Float64
end
B200ContinuousCallback(condition, affect, affect_neg = affect; rootfind = 1, interp_points = 10,
                       save_positions = (true, true), abstol = 10eps(), repeat_nudge = 1 / 100) =
    B200ContinuousCallback(condition, affect, affect_neg, rootfind, interp_points, save_positions, abstol, repeat_nudge)
struct B200DiscreteCallback
    condition::Tuple{String, String}
    affect::Union{Nothing, Tuple{String, String}}
    save_positions::Tuple{Bool, Bool}
end
B200DiscreteCallback(condition, affect; save_positions = (true, true)) = B200DiscreteCallback(condition, affect, save_positions)
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
alg_id(::Rodas3P) = 18; alg_id(::Rodas23W) = 19
# AutoTsit5(Rosenbrock23()) = CompositeAlgorithm((Tsit5(), Rosenbrock23()), AutoSwitch(...)) with the default switch parameters
const AutoTsit5Ros23 = CompositeAlgorithm{<:Any, <:Tuple{Tsit5, Rosenbrock23}, <:AutoSwitch}
function alg_id(alg::AutoTsit5Ros23)
    c = alg.choice_function
    (c.maxstiffstep == 10 && c.maxnonstiffstep == 3 && c.nonstifftol == 9 // 10 && c.stifftol == 9 // 10 && c.dtfac == 2 &&
     !c.stiffalgfirst && c.switch_max == 5) || throw(ArgumentError("EnsembleB200 serves AutoTsit5(Rosenbrock23()) with the default AutoSwitch parameters"))
    return 17
end
const StiffAlgs = Union{Rosenbrock23, Rosenbrock32, Rodas5P, Rodas5Pe, Rodas5, Rodas4, Rodas42, Rodas4P, Rodas4P2, Rodas3P, Rodas23W,
                        AutoTsit5Ros23}
const B200Algs = Union{Tsit5, Vern6, Vern7, Vern8, Vern9, DP5, BS3, StiffAlgs}
isstiff(alg) = alg isa StiffAlgs
const RETCODES = (ReturnCode.Default, ReturnCode.Success, ReturnCode.MaxIters, ReturnCode.DtLessThanMin,
                  ReturnCode.Unstable, ReturnCode.DtNaN, ReturnCode.Terminated)

# the structs above against the library they are handed to (b200ode_struct_size), once per session
const ABI_CHECKED = Ref(false)
function check_abi()
    ABI_CHECKED[] && return
    for (which, S) in enumerate((B200Problem, B200Opts, B200Result, nothing, nothing, nothing, B200CallbackSrc, B200Ragged))
        S === nothing && continue
        n = ccall((:b200ode_struct_size, LIB), Cint, (Cint,), which - 1)
        n == sizeof(S) || error("EnsembleB200: $(S) is $(sizeof(S)) bytes here and $n in $LIB — binding and library are out of step")
    end
    ABI_CHECKED[] = true
end

function check(rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:b200ode_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))
    rc == -1 ? throw(ArgumentError(msg)) : error("b200ode ($rc): $msg")
end

# ---- code generation: Symbolics traces f(u,p,t) and emits C ------------------------------------
function c_sources(prob, alg, ::Type{T}) where {T}
    n, np = length(prob.u0), (prob.p === nothing || prob.p isa SciMLBase.NullParameters) ? 0 : length(prob.p)
    @variables u[1:n] p[1:max(np, 1)] t
    us, ps = collect(u), collect(p)
    du = SciMLBase.isinplace(prob) ? (d = similar(us, Num); prob.f.f(d, us, ps, t); d) : collect(prob.f.f(us, ps, t))
    fix(s) = T === Float32 ? replace(s, "double" => "float") : s
    rhs = fix(build_function(du, us, ps, t; target = Symbolics.CTarget(), fname = :diffeqf))
    jac = tgr = nothing
    if isstiff(alg)
        # ODEFunction(f; jac, tgrad) given by the user: trace those; otherwise differentiate symbolically
        J = prob.f.jac === nothing ? Symbolics.jacobian(du, us) :
            (SciMLBase.isinplace(prob) ? (M = Matrix{Num}(undef, n, n); prob.f.jac(M, us, ps, t); M) : collect(prob.f.jac(us, ps, t)))
        dT = prob.f.tgrad === nothing ? Symbolics.derivative.(du, t) :
             (SciMLBase.isinplace(prob) ? (v = similar(us, Num); prob.f.tgrad(v, us, ps, t); v) : collect(prob.f.tgrad(us, ps, t)))
        jac = fix(build_function(vec(J), us, ps, t; target = Symbolics.CTarget(), fname = :diffeqjac))  # column major
        tgr = fix(build_function(dT, us, ps, t; target = Symbolics.CTarget(), fname = :diffeqtgrad))
    end
    return n, np, rhs, jac, tgr
end

# ---- handles and the program cache ----------------------------------------------------------------
# One handle per device list for the life of the process; one compiled program per
# (devices, alg, T, n, np, hash of the generated sources, variant options).  Programs are destroyed
# (b200ode_[multi_]program_destroy) when the cache is emptied or at exit — never leaked per solve.
const HANDLES = Dict{Vector{Int}, Ptr{Cvoid}}()
const PROGRAMS = Dict{Any, Tuple{Ptr{Cvoid}, Bool}}()
const CACHE_LOCK = ReentrantLock()

function handle(ens::EnsembleB200)
    lock(CACHE_LOCK) do
        get!(HANDLES, ens.devices) do
            h = Ref{Ptr{Cvoid}}(C_NULL)
            if ismulti(ens)
                ids = Cint.(ens.devices)
                check(ccall((:b200ode_multi_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Cint}, Cint),
                            h, isempty(ids) ? C_NULL : pointer(ids), length(ids)))
            else
                check(ccall((:b200ode_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), h, ens.devices[1]))
            end
            h[]
        end
    end
end

cb_list(::Nothing) = ()
cb_list(cb::Union{B200ContinuousCallback, B200DiscreteCallback}) = (cb,)
cb_list(cbs::Tuple) = cbs
cb_list(cb) = throw(ArgumentError("EnsembleB200: pass callbacks as B200ContinuousCallback / B200DiscreteCallback (C source), got $(typeof(cb))"))

# b200ode_compile_callbacks (single device).  The Cstring fields point into `keep`, which outlives the ccall.
function compile_with_callbacks(h, prog, alg, ::Type{T}, n, np, rhs, jac, tgr, extra, cbs) where {T}
    keep = String[]
    cs(x) = x === nothing ? Cstring(C_NULL) : (push!(keep, x); Base.unsafe_convert(Cstring, keep[end]))
    pair(x) = x === nothing ? (nothing, nothing) : x
    arr = map(cbs) do cb
        cont = cb isa B200ContinuousCallback
        a, an = pair(cb.affect); ng, ngn = cont ? pair(cb.affect_neg) : (nothing, nothing)
        B200CallbackSrc(cont ? 1 : 0, cont ? cb.rootfind : 1, cs(cb.condition[1]), cs(cb.condition[2]), cs(a), cs(an),
                        cs(ng), cs(ngn), cont ? cb.interp_points : 0, cb.save_positions[1], cb.save_positions[2], 0,
                        cont ? cb.abstol : -1.0, cont ? cb.repeat_nudge : -1.0)
    end |> collect
    GC.@preserve keep arr begin
        check(ccall((:b200ode_compile_callbacks, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cstring, Cstring, Cstring, Cstring, Cstring, Cstring,
                     Ptr{B200CallbackSrc}, Cint, Cstring),
                    h, prog, alg_id(alg), T === Float32 ? 1 : 0, n, np, rhs, "diffeqf", jac === nothing ? C_NULL : jac, "diffeqjac",
                    tgr === nothing ? C_NULL : tgr, "diffeqtgrad", pointer(arr), length(arr), isempty(extra) ? C_NULL : extra))
    end
end

function program(ens::EnsembleB200, h, alg, ::Type{T}, n, np, rhs, jac, tgr, extra::String, cbs = ()) where {T}
    key = (ens.devices, alg_id(alg), T, n, np, hash(rhs), hash(jac), hash(tgr), extra, hash(cbs))
    lock(CACHE_LOCK) do
        get!(PROGRAMS, key) do
            prog = Ref{Ptr{Cvoid}}(C_NULL)
            args = (h, prog, alg_id(alg), T === Float32 ? 1 : 0, n, np, rhs, "diffeqf",
                    jac === nothing ? C_NULL : jac, "diffeqjac", tgr === nothing ? C_NULL : tgr, "diffeqtgrad",
                    isempty(extra) ? C_NULL : extra)
            if !isempty(cbs)
                ismulti(ens) && throw(ArgumentError("EnsembleB200: callbacks are single-device"))
                compile_with_callbacks(h, prog, alg, T, n, np, rhs, jac, tgr, extra, cbs)
            elseif ismulti(ens)
                check(ccall((:b200ode_multi_compile, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cstring, Cstring, Cstring, Cstring, Cstring, Cstring, Cstring),
                            args...))
            else
                check(ccall((:b200ode_compile, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cstring, Cstring, Cstring, Cstring, Cstring, Cstring, Cstring),
                            args...))
            end
            (prog[], ismulti(ens))
        end
    end
end

"Destroy every cached program and handle (also registered with atexit)."
function release!()
    lock(CACHE_LOCK) do
        for (prog, multi) in values(PROGRAMS)
            multi ? ccall((:b200ode_multi_program_destroy, LIB), Cint, (Ptr{Cvoid},), prog) :
                    ccall((:b200ode_program_destroy, LIB), Cint, (Ptr{Cvoid},), prog)
        end
        empty!(PROGRAMS)
        for (devs, h) in HANDLES
            length(devs) != 1 ? ccall((:b200ode_multi_destroy, LIB), Cint, (Ptr{Cvoid},), h) :
                                ccall((:b200ode_destroy, LIB), Cint, (Ptr{Cvoid},), h)
        end
        empty!(HANDLES)
    end
end
__init__() = atexit(release!)

# page-lock an output array for the duration of a call (D2H then runs at PCIe speed, profiles/r1_time_pageable_vs_pinned.json)
function with_pinned(f, arrays...)
    pinned = Any[]
    try
        for a in arrays
            (a === nothing || sizeof(a) < (1 << 20)) && continue
            ccall((:b200ode_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), pointer(a), sizeof(a)) == 0 && push!(pinned, a)
        end
        return f()
    finally
        for a in pinned
            ccall((:b200ode_host_unregister, LIB), Cint, (Ptr{Cvoid},), pointer(a))
        end
    end
end

const ALLOWED = (:trajectories, :batch_size, :saveat, :save_start, :save_end, :save_everystep, :save_idxs, :tstops, :d_discontinuities, :reltol, :abstol,
                 :dt, :dtmin, :dtmax, :maxiters, :adaptive, :dense, :verbose, :progress, :callback)

# One batch of trajectories I (global sim ids) with repeat counters `rep`: harvest prob_func, solve, wrap.
function solve_ids(eprob, prob, alg, ens, h, prog, multi, opts, grid, tstops, everystep, idxs, w, n, np, T, t0, tf, I, rep, tspans_variant = false)
    N = length(I)
    U0 = Matrix{T}(undef, n, N); P = Matrix{T}(undef, max(np, 1), N)
    spans = Matrix{Float64}(undef, 2, N)
    for (k, i) in enumerate(I)
        q = eprob.prob_func(prob, SciMLBase.EnsembleContext(i, rep[k], nothing))
        spans[1, k], spans[2, k] = q.tspan
        U0[:, k] .= q.u0
        np > 0 && (P[:, k] .= q.p)
    end
    # a prob_func that changes tspan: the program must be the per-trajectory-span variant (chosen by __solve from
    # trajectory 1; a prob_func that changes tspan for some trajectories only is not served)
    varying = any(spans[1, k] != t0 || spans[2, k] != tf for k in 1:N)
    varying && !tspans_variant && throw(ArgumentError("EnsembleB200: prob_func changes tspan for some trajectories but not for trajectory 1"))
    cprob = B200Problem(N, pointer(U0), 0, pointer(P), 0, t0, tf, tspans_variant ? pointer(spans) : Ptr{Float64}(C_NULL))
    nslots = ccall((:b200ode_nslots, LIB), Cint, (Ref{B200Problem}, Ref{B200Opts}), cprob, opts)
    tspans_variant && (nslots = 0)      # no rectangular rows with per-trajectory spans: [t0_i, tf_i] / [u0_i, u(tf_i)] are built below
    uf = Matrix{T}(undef, n, N); tfin = Vector{Float64}(undef, N)
    us = Array{T, 3}(undef, w, max(nslots, 1), N); ts = Vector{Float64}(undef, max(nslots, 1))
    cnt = [zeros(Int32, N) for _ in 1:8]
    res = B200Result(pointer(uf), pointer(tfin), nslots > 0 ? pointer(us) : C_NULL, pointer(ts),
                     pointer.(cnt)..., 0.0, 0.0)
    rag = Ref(B200Ragged(0, C_NULL, C_NULL, C_NULL))
    GC.@preserve U0 P spans grid tstops uf tfin us ts cnt begin
        with_pinned(us, uf) do
            if everystep     # ragged rows: trajectory k owns rows offs[k]+1 : offs[k+1]  (single device)
                check(ccall((:b200ode_solve_everystep, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Cvoid}, Ref{B200Problem}, Ref{B200Opts}, Ref{B200Result}, Ref{B200Ragged}),
                            h, prog, cprob, opts, res, rag))
            elseif multi
                check(ccall((:b200ode_multi_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{B200Problem}, Ref{B200Opts}, Ref{B200Result}),
                            h, prog, cprob, opts, res))
            else
                check(ccall((:b200ode_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{B200Problem}, Ref{B200Opts}, Ref{B200Result}),
                            h, prog, cprob, opts, res))
            end
        end
    end
    nsaved, naccept, nreject, nf, njacs, nw, nsolve, rc = cnt
    offs = everystep ? copy(unsafe_wrap(Array, rag[].row_offsets, N + 1)) : Int64[]
    rts = everystep ? copy(unsafe_wrap(Array, rag[].ts, rag[].total_rows)) : Float64[]
    rus = everystep ? copy(unsafe_wrap(Array, Ptr{T}(rag[].us), (w, Int(rag[].total_rows)))) : Matrix{T}(undef, 0, 0)
    if everystep
        for ptr in (rag[].row_offsets, rag[].ts, rag[].us)
            ccall((:b200ode_free, LIB), Cvoid, (Ptr{Cvoid},), ptr)
        end
    end
    sel(v) = idxs === nothing ? v : v[idxs]
    ss = opts.save_start != 0
    se = opts.save_end != 0
    sols = map(1:N) do k
        if everystep
            r = (offs[k] + 1):offs[k + 1]
            tk = rts[r]; uk = [SVector{w, T}(@view rus[:, j]) for j in r]
        elseif nslots > 0
            tk = ts[1:nsaved[k]]; uk = [SVector{w, T}(@view us[:, s, k]) for s in 1:nsaved[k]]
        else    # no saveat grid, save_everystep = false: the start and/or end point, as save_start / save_end say
            tk = Float64[]; uk = SVector{w, T}[]
            ss && (push!(tk, tspans_variant ? spans[1, k] : t0); push!(uk, SVector{w, T}(sel(@view U0[:, k]))))
            se && (push!(tk, tfin[k]); push!(uk, SVector{w, T}(sel(@view uf[:, k]))))
        end
        stats = SciMLBase.DEStats(Int(nf[k]), 0, 0, Int(nw[k]), Int(nsolve[k]), Int(njacs[k]), 0, 0, 0, 0,
                                  Int(naccept[k]), Int(nreject[k]), 0.0)
        build_solution(prob, alg, tk, uk; dense = false, stats, retcode = RETCODES[rc[k] + 1])
    end
    return sols
end

function __solve(eprob::AbstractEnsembleProblem, alg::B200Algs, ens::EnsembleB200;
                 trajectories, batch_size = trajectories, kwargs...)
    prob = eprob.prob
    kw = merge(NamedTuple(prob.kwargs), NamedTuple(kwargs))          # merge_problem_kwargs
    for k in keys(kw)
        k in ALLOWED || throw(ArgumentError("EnsembleB200 does not support the keyword $k"))
    end
    check_abi()
    T = eltype(prob.u0)
    T <: Union{Float32, Float64} || throw(ArgumentError("EnsembleB200 supports Float32/Float64 states"))
    t0, tf = Float64.(prob.tspan)
    tf != t0 || throw(ArgumentError("EnsembleB200: tspan must have tf != t0"))
    # tspan[2] < tspan[1] (tdir = -1, solve.jl:273): a program compiled with B200ODE_OPT_REVERSE_TIME; the saveat list is
    # handed over in the order the integrator meets it (initialize_saveat, solve.jl:1103-1124)
    tdir = sign(tf - t0)
    saveat = get(kw, :saveat, ())
    grid = saveat isa Number ? collect(Float64, (t0 + tdir * abs(saveat)):(tdir * abs(saveat)):tf) :
           sort!(Float64[s for s in saveat if tdir * t0 < tdir * s <= tdir * tf]; rev = tdir < 0)
    everystep = get(kw, :save_everystep, isempty(grid))            # solve.jl:138
    everystep && ismulti(ens) && throw(ArgumentError("EnsembleB200: save_everystep output is ragged and single-device; pass one device"))
    n, np, rhs, jac, tgr = c_sources(prob, alg, T)
    idxs = get(kw, :save_idxs, nothing)
    idxs isa Integer && (idxs = [idxs])
    w = idxs === nothing ? n : length(idxs)                        # components per saved row
    extra = String[]
    tdir < 0 && push!(extra, "-DB200_REVERSE=1")
    everystep && push!(extra, "-DB200_EVERYSTEP=1")
    idxs === nothing || push!(extra, "-DB200_SAVE_IDXS=" * join(idxs .- 1, ","))   # the C side is 0-based
    tstops = collect(Float64, get(kw, :tstops, ()))
    discs = collect(Float64, get(kw, :d_discontinuities, ()))     # stops with a one-ulp shift and a fresh first stage
    (isempty(tstops) && isempty(discs)) || push!(extra, "-DB200_TSTOPS=1")
    adaptive = get(kw, :adaptive, true)
    adaptive || push!(extra, "-DB200_ADAPTIVE=0")
    adaptive || get(kw, :dt, nothing) !== nothing || !isempty(tstops) ||
        throw(ArgumentError("Fixed timestep methods require a choice of dt or choosing the tstops"))
    # callbacks: rows forced by save_positions make the output ragged even with save_everystep = false
    cbs = cb_list(get(kw, :callback, nothing))
    flags = Int32(0)
    if !isempty(cbs)
        alg isa Tsit5 || throw(ArgumentError("EnsembleB200: callbacks are available with Tsit5()"))
        if any(any(cb.save_positions) for cb in cbs) && !everystep
            push!(extra, "-DB200_EVERYSTEP=1"); flags = Int32(2)         # B200ODE_FLAG_NO_STEP_ROWS
            everystep = true
        end
    end
    # abstol / reltol as vectors: one tolerance per component (solve.jl:377-399)
    rtol, atol = get(kw, :reltol, 0.0), get(kw, :abstol, 0.0)
    rtv = rtol isa AbstractVector ? collect(Float64, rtol) : Float64[]
    atv = atol isa AbstractVector ? collect(Float64, atol) : Float64[]
    (isempty(rtv) && isempty(atv)) || push!(extra, "-DB200_VECTOR_TOL=1")
    # prob_func remakes tspan (probed on trajectory 1): per-trajectory spans — final states or the ragged output only
    tspans_variant = eprob.prob_func(prob, SciMLBase.EnsembleContext(1, 1, nothing)).tspan != prob.tspan
    tdir < 0 && tspans_variant &&
        throw(ArgumentError("EnsembleB200: reverse-time integration is not combined with per-trajectory tspan"))
    if tspans_variant
        (isempty(tstops) && isempty(discs) && isempty(cbs)) ||
            throw(ArgumentError("EnsembleB200: per-trajectory tspan is not combined with tstops, d_discontinuities or callbacks"))
        (everystep || isempty(saveat)) ||
            throw(ArgumentError("EnsembleB200: per-trajectory tspan with saveat needs save_everystep = true (ragged output)"))
        push!(extra, "-DB200_TSPANS=1")
    end
    # Vern7 on a wide state with nothing but start / end rows: the kernel that keeps k1..k10 in shared memory
    # (B200ODE_OPT_SMEM_STAGES; bit-identical results, it serves no interior saveat rows and no callbacks)
    if alg isa Vern7 && n >= 24 && !everystep && isempty(cbs) && tdir > 0       # measured crossover: scripts/time_wide_threshold.py
        t0w, tfw = prob.tspan
        grid_pts = saveat isa Number ? (saveat > 0 && saveat < abs(tfw - t0w) ? (1,) : ()) : filter(t -> t0w < t < tfw, collect(saveat))
        isempty(grid_pts) && push!(extra, "-DB200_WIDE=1")
    end
    h = handle(ens)
    prog, multi = program(ens, h, alg, T, n, np, rhs, jac, tgr, join(extra, " "), cbs)
    # defaults of solve.jl:141-143,596-599 (the C ABI only sees the expanded grid)
    dflt(tend) = everystep || isempty(saveat) || saveat isa Number || tend in saveat
    ss = something(get(kw, :save_start, nothing), dflt(prob.tspan[1]))
    se = get(kw, :save_end, nothing)
    se === nothing && !dflt(prob.tspan[2]) && (se = false)
    opts = B200Opts(isempty(rtv) ? Float64(rtol) : 0.0, isempty(atv) ? Float64(atol) : 0.0, something(get(kw, :dt, nothing), 0.0),
                    get(kw, :dtmin, 0.0), get(kw, :dtmax, 0.0), get(kw, :maxiters, 0),
                    isempty(grid) ? Ptr{Float64}(C_NULL) : pointer(grid), length(grid),
                    Int32(ss), se === nothing ? Int32(-1) : Int32(se), flags, 0,
                    isempty(tstops) ? Ptr{Float64}(C_NULL) : pointer(tstops), length(tstops), 0,
                    isempty(atv) ? Ptr{Float64}(C_NULL) : pointer(atv), isempty(rtv) ? Ptr{Float64}(C_NULL) : pointer(rtv),
                    isempty(discs) ? Ptr{Float64}(C_NULL) : pointer(discs), length(discs), 0)
    tstart = time()
    tol_keep = (atv, rtv, discs)   # opts points into these: keep them reachable until the last batch returns
    u = eprob.u_init === nothing ? [] : eprob.u_init
    converged = false
    for b0 in 1:batch_size:trajectories
        I = collect(b0:min(b0 + batch_size - 1, trajectories))
        batch = Vector{Any}(undef, length(I))
        pending = collect(1:length(I))           # positions of the batch still to be (re)solved
        rep = ones(Int, length(I))               # ctx.repeat of each position
        # output_func's rerun flag (SciMLBase solve_batch: `while rerun; prob_func(prob, ctx(repeat += 1)); solve; end`),
        # served batch-wise: everything flagged is re-submitted to the GPU together until nothing is flagged
        while !isempty(pending)
            sols = solve_ids(eprob, prob, alg, ens, h, prog, multi, opts, grid, tstops, everystep, idxs, w, n, np, T, t0, tf,
                             I[pending], rep[pending], tspans_variant)
            again = Int[]
            for (j, pos) in enumerate(pending)
                out, rerun = eprob.output_func(sols[j], SciMLBase.EnsembleContext(I[pos], rep[pos], nothing))
                if rerun
                    rep[pos] += 1
                    push!(again, pos)
                else
                    batch[pos] = out
                end
            end
            pending = again
        end
        u, converged = eprob.reduction(u, batch, I)
        converged && break
    end
    GC.@preserve tol_keep nothing
    return EnsembleSolution(u, time() - tstart, converged)
end

end # module
