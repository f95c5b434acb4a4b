# julia_golden.jl — generate golden vectors from the REAL reference for the ensemble hot path.
#
#   julia --project=<env with OrdinaryDiffEq, StaticArrays> -t auto scripts/julia_golden.jl [outdir]
#
# No Julia runtime exists in the build container or on the GPU boxes (probed: `which julia` is empty on both,
# DESIGN.md §2), so this script has never been run there; it is committed so that anyone with a Julia install can
# produce tests/golden/julia/*.json, after which tests/test_julia_golden.py compares the CPU oracle (and, with -m gpu,
# the CUDA path) against the reference's own numbers and the "parity unpinned" xfail turns into a real check.
#
# What it runs: for BASELINE configs 1-4 on the first 256 trajectories of the SplitMix64 tables (SURVEY §8(d);
# same tables as ordinarydiffeq.jl_b200/problems_library.py) the reference's own
#     solve(EnsembleProblem(prob; prob_func), alg, EnsembleThreads(); trajectories, saveat, reltol, abstol)
# (ensemble construction as in lib/DiffEqBase/test/downstream/ensemble.jl:51-112) with out-of-place SVector problems and
# ODEFunction(f; jac, tgrad) for the Rosenbrock methods, and dumps per trajectory: naccept, nreject, nf (+ njacs, nw,
# nsolve), retcode, u(tf) and the saveat rows.  Floats are printed with Julia's shortest round-trip representation,
# so the JSON holds the exact bits.
using OrdinaryDiffEq, StaticArrays

const OUT = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia")
mkpath(OUT)
const NTRAJ = 256

function U(i::Integer, j::Integer)
    x = 0x9E3779B97F4A7C15 ⊻ (UInt64(i) * UInt64(4) + UInt64(j))
    z = x + 0x9E3779B97F4A7C15
    z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
    z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
    z = z ⊻ (z >> 31)
    return Float64(z >> 11) / 9007199254740992.0
end

# ---- problems (operation order identical to problems_library.py's C sources) ----
lorenz(u, p, t) = SVector(p[1] * (u[2] - u[1]), u[1] * (p[2] - u[3]) - u[2], u[1] * u[2] - p[3] * u[3])

rober(u, p, t) = SVector(-p[1] * u[1] + p[3] * u[2] * u[3],
                         p[1] * u[1] - p[2] * (u[2] * u[2]) - p[3] * u[2] * u[3],
                         p[2] * (u[2] * u[2]))
function rober_jac(u, p, t)
    T = eltype(u)
    # column-major fill order = (row, col) entries of d f_i / d u_j
    return SMatrix{3, 3, T}(-p[1], p[1], zero(T),
                            p[3] * u[3], -T(2) * p[2] * u[2] - p[3] * u[3], T(2) * p[2] * u[2],
                            p[3] * u[2], -p[3] * u[2], zero(T))
end
rober_tgrad(u, p, t) = zero(u)

function pleiades(u, p, t)
    T = eltype(u)
    du = MVector{28, T}(undef)
    for i in 1:7
        du[i] = u[14 + i]; du[7 + i] = u[21 + i]
    end
    for i in 1:7
        ax = zero(T); ay = zero(T)
        for j in 1:7
            j == i && continue
            dx = u[j] - u[i]; dy = u[7 + j] - u[7 + i]
            r = sqrt(dx * dx + dy * dy); r3 = r * r * r
            m = T(j)
            ax = ax + m * dx / r3; ay = ay + m * dy / r3
        end
        du[14 + i] = ax; du[21 + i] = ay
    end
    return SVector(du)
end
const PLEIADES_U0 = [3.0, 3.0, -1.0, -3.0, 2.0, -2.0, 2.0, 3.0, -3.0, 2.0, 0, 0, -4.0, 4.0,
                     0, 0, 0, 0, 0, 1.75, -1.5, 0, 0, 0, -1.25, 1, 0, 0]

# ---- JSON without a package ----
jval(x::AbstractFloat) = isfinite(x) ? string(x) : (isnan(x) ? "\"NaN\"" : (x > 0 ? "\"Inf\"" : "\"-Inf\""))
jval(x::Integer) = string(x)
jval(x::AbstractString) = "\"" * x * "\""
jval(v::Union{AbstractArray, Tuple}) = "[" * join((jval(x) for x in v), ",") * "]"

function dump(name, sim, meta)
    fields = Pair{String, Any}[meta...]
    stat(f) = [Int(f(s.stats)) for s in sim.u]
    push!(fields, "naccept" => stat(st -> st.naccept), "nreject" => stat(st -> st.nreject), "nf" => stat(st -> st.nf),
          "njacs" => stat(st -> st.njacs), "nw" => stat(st -> st.nw), "nsolve" => stat(st -> st.nsolve),
          "retcode" => [string(s.retcode) for s in sim.u],
          "t" => [collect(s.t) for s in sim.u],
          "u" => [[collect(u) for u in s.u] for s in sim.u])
    open(joinpath(OUT, name * ".json"), "w") do io
        println(io, "{")
        for (i, (k, v)) in enumerate(fields)
            println(io, "  ", jval(k), ": ", jval(v), i == length(fields) ? "" : ",")
        end
        println(io, "}")
    end
end

# prob_func in the reference's current form (prob, ctx) with ctx.sim_id
# (lib/DiffEqBase/test/downstream/ensemble.jl:85-88); tables instead of ctx.rng so the inputs are reproducible anywhere
function run_case(name, prob, P, U0, alg; T = Float64, kw...)
    prob_func = function (prob, ctx)
        i = ctx.sim_id
        return remake(prob; u0 = U0 === nothing ? prob.u0 : U0[i], p = P === nothing ? prob.p : P[i])
    end
    ens = EnsembleProblem(prob; prob_func, safetycopy = false)
    sim = solve(ens, alg, EnsembleThreads(); trajectories = NTRAJ, kw...)
    dump(name, sim, ["case" => name, "alg" => string(nameof(typeof(alg))), "dtype" => string(T),
                     "trajectories" => NTRAJ, "julia" => string(VERSION)])
    println(name, ": ", sum(s.stats.naccept for s in sim.u), " accepted steps")
end

# config 1: Lorenz/Tsit5, reltol 1e-8, final state
let
    P = [SVector(10.0, 28.0 * (0.5 + U(i - 1, 0)), 8 / 3) for i in 1:NTRAJ]
    prob = ODEProblem{false}(lorenz, SVector(1.0, 0.0, 0.0), (0.0, 10.0), P[1])
    run_case("cfg1_lorenz_tsit5_reltol1e-8", prob, P, nothing, Tsit5(); reltol = 1e-8, save_everystep = false)
    # config 2: saveat = 0.1, default tolerances, FP64 and FP32
    run_case("cfg2_lorenz_tsit5_saveat_f64", prob, P, nothing, Tsit5(); saveat = 0.1)
    P32 = [SVector{3, Float32}(p) for p in P]
    prob32 = ODEProblem{false}(lorenz, SVector(1.0f0, 0.0f0, 0.0f0), (0.0f0, 10.0f0), P32[1])
    run_case("cfg2_lorenz_tsit5_saveat_f32", prob32, P32, nothing, Tsit5(); T = Float32, saveat = 0.1f0)
    # other explicit steppers of §8(f) row 3 on the same inputs
    for (nm, alg) in (("dp5", DP5()), ("bs3", BS3()), ("vern6", Vern6()), ("vern7", Vern7()), ("vern8", Vern8()), ("vern9", Vern9()))
        run_case("lorenz_$(nm)_saveat_f64", prob, P, nothing, alg; saveat = 0.5)
    end
end

# keywords added in the last third of round 2: tstops + d_discontinuities (the stop, the one-ulp shift and the fresh first
# stage change the step sequence even for an autonomous RHS), and a prob_func that remakes tspan
let
    P = [SVector(10.0, 28.0 * (0.5 + U(i - 1, 0)), 8 / 3) for i in 1:NTRAJ]
    prob = ODEProblem{false}(lorenz, SVector(1.0, 0.0, 0.0), (0.0, 10.0), P[1])
    run_case("lorenz_tsit5_d_discontinuities", prob, P, nothing, Tsit5(); d_discontinuities = [2.5, 5.0], tstops = [7.5],
             saveat = 0.5)
    run_case("lorenz_vern7_d_discontinuities", prob, P, nothing, Vern7(); d_discontinuities = [0.0, 2.5], save_everystep = false)
    spans = [(0.0, 5.0 + 5.0 * U(i - 1, 1)) for i in 1:NTRAJ]
    prob_func = (prob, ctx) -> remake(prob; p = P[ctx.sim_id], tspan = spans[ctx.sim_id])
    sim = solve(EnsembleProblem(prob; prob_func, safetycopy = false), Tsit5(), EnsembleThreads(); trajectories = NTRAJ,
                save_everystep = false)
    dump("lorenz_tsit5_tspans", sim, ["case" => "lorenz_tsit5_tspans", "alg" => "Tsit5", "dtype" => "Float64",
                                      "trajectories" => NTRAJ, "julia" => string(VERSION)])
end

# reverse time (tspan[2] < tspan[1], tdir = -1): Lorenz backwards with a saveat step and a stop, Robertson backwards over
# a short span with Rodas5P (the CUDA path integrates the mirrored problem, B200ODE_OPT_REVERSE_TIME)
let
    P = [SVector(10.0, 28.0 * (0.5 + U(i - 1, 0)), 8 / 3) for i in 1:NTRAJ]
    prob = ODEProblem{false}(lorenz, SVector(1.0, 0.0, 0.0), (1.0, 0.0), P[1])
    run_case("lorenz_tsit5_reverse", prob, P, nothing, Tsit5(); saveat = 0.1, tstops = [0.5])
    run_case("lorenz_vern7_reverse", prob, P, nothing, Vern7(); reltol = 1e-8, abstol = 1e-10, save_everystep = false)
    base = (0.04, 3.0e7, 1.0e4)
    PR = [SVector(ntuple(j -> base[j] * (0.5 + U(i - 1, j - 1)), 3)) for i in 1:NTRAJ]
    f = ODEFunction{false}(rober; jac = rober_jac, tgrad = rober_tgrad)
    probr = ODEProblem(f, SVector(1.0, 0.0, 0.0), (1.0e-3, 0.0), PR[1])
    run_case("robertson_rodas5p_reverse", probr, PR, nothing, Rodas5P(); reltol = 1e-6, abstol = 1e-8, save_everystep = false)
end

# config 3: Robertson, Rodas5P and Rosenbrock23 (+ the RodasTableau family), jac + tgrad supplied
let
    base = (0.04, 3.0e7, 1.0e4)
    P = [SVector(ntuple(j -> base[j] * (0.5 + U(i - 1, j - 1)), 3)) for i in 1:NTRAJ]
    f = ODEFunction{false}(rober; jac = rober_jac, tgrad = rober_tgrad)
    prob = ODEProblem(f, SVector(1.0, 0.0, 0.0), (0.0, 1.0e5), P[1])
    for (nm, alg) in (("rodas5p", Rodas5P()), ("rosenbrock23", Rosenbrock23()), ("rodas5", Rodas5()), ("rodas4", Rodas4()),
                      ("rodas42", Rodas42()), ("rodas4p", Rodas4P()), ("rodas4p2", Rodas4P2()), ("rodas5pe", Rodas5Pe()),
                      ("rodas3p", Rodas3P()), ("rodas23w", Rodas23W()), ("autotsit5_rosenbrock23", AutoTsit5(Rosenbrock23())))
        run_case("cfg3_robertson_$(nm)", prob, P, nothing, alg; reltol = 1e-6, abstol = 1e-8, save_everystep = false)
    end
    run_case("cfg3_robertson_rodas5p_saveat", prob, P, nothing, Rodas5P(); reltol = 1e-6, abstol = 1e-8,
             saveat = [100.0, 1000.0, 5.0e4])
end

# config 4: Pleiades/Vern7
let
    U0 = map(1:NTRAJ) do i
        u = copy(PLEIADES_U0)
        for j in 0:13
            u[j + 1] += 0.01 * (2.0 * U((i - 1) * 4 + j ÷ 4, j % 4) - 1.0)
        end
        SVector{28}(u)
    end
    prob = ODEProblem{false}(pleiades, U0[1], (0.0, 3.0))
    run_case("cfg4_pleiades_vern7", prob, nothing, U0, Vern7(); reltol = 1e-6, abstol = 1e-8, save_everystep = false)
end

# callbacks (Tsit5): the bouncing balls of tests/test_gpu_parity.py::test_callbacks_bouncing_ball_parity
# (p = (g, e): y'' = -g, v -> -e v at y = 0, downcrossings only; default save_positions, saveat = 0.5)
let
    ball(u, p, t) = SVector(u[2], -p[1])
    P = [SVector(9.81 * (0.5 + U(i - 1, 0)), 0.8 + 0.2 * U(i - 1, 1)) for i in 1:NTRAJ]
    prob = ODEProblem{false}(ball, SVector(50.0, 0.0), (0.0, 15.0), P[1])
    condition(u, t, integrator) = u[1]
    bounce!(integrator) = (integrator.u = SVector(integrator.u[1], -integrator.p[2] * integrator.u[2]))
    cb = ContinuousCallback(condition, nothing, bounce!)
    run_case("callbacks_bouncing_ball_tsit5", prob, P, nothing, Tsit5(); callback = cb, saveat = 0.5)
    term = ContinuousCallback(condition, terminate!)
    run_case("callbacks_bouncing_ball_terminate_tsit5", prob, P, nothing, Tsit5(); callback = term)
end
