"""C sources of the benchmark right-hand sides named by BASELINE.json's configs, written
operation-for-operation like the reference's Julia definitions, plus the deterministic
synthetic parameter tables.

    Lorenz     /root/reference/lib/OrdinaryDiffEqCore/src/precompilation_setup.jl:1-10, with
               (sigma, rho, beta) as parameters so prob_func can randomise rho
    Robertson  /root/reference/benchmark/benchmarks.jl:93-107   (+ analytic jac, tgrad = 0)
    Pleiades   /root/reference/benchmark/benchmarks.jl:29-67    (standard RHS: accelerations
               accumulate from zero — the file's `fill!(du[15:21], 0.0)` zero-fills copies)
"""
import numpy as np


def _ty(f32):
    return "float" if f32 else "double"


def lorenz_source(f32=False, name="lorenz_rhs"):
    T = _ty(f32)
    return ("void %s(%s* du, const %s* u, const %s* p, const %s t) {\n"
            "  du[0] = p[0] * (u[1] - u[0]);\n"
            "  du[1] = u[0] * (p[1] - u[2]) - u[1];\n"
            "  du[2] = u[0] * u[1] - p[2] * u[2];\n"
            "}\n" % (name, T, T, T, T)), name


def robertson_sources(f32=False):
    T = _ty(f32)
    sig = "(%s* %%s, const %s* u, const %s* p, const %s t)" % (T, T, T, T)
    rhs = ("void rober_rhs" + sig % "du" + " {\n"
           "  du[0] = -p[0] * u[0] + p[2] * u[1] * u[2];\n"
           "  du[1] = p[0] * u[0] - p[1] * (u[1] * u[1]) - p[2] * u[1] * u[2];\n"
           "  du[2] = p[1] * (u[1] * u[1]);\n"
           "}\n")
    zero = "0.0f" if f32 else "0.0"
    two = "2.0f" if f32 else "2.0"
    # column-major 3x3: J[i + 3*j] = d f_i / d u_j
    jac = ("void rober_jac" + sig % "J" + " {\n"
           "  J[0] = -p[0];\n"
           "  J[1] = p[0];\n"
           "  J[2] = %s;\n"
           "  J[3] = p[2] * u[2];\n"
           "  J[4] = -%s * p[1] * u[1] - p[2] * u[2];\n"
           "  J[5] = %s * p[1] * u[1];\n"
           "  J[6] = p[2] * u[1];\n"
           "  J[7] = -p[2] * u[1];\n"
           "  J[8] = %s;\n"
           "}\n" % (zero, two, two, zero))
    tgrad = ("void rober_tgrad" + sig % "dT" + " {\n"
             "  dT[0] = %s; dT[1] = %s; dT[2] = %s;\n"
             "}\n" % (zero, zero, zero))
    return (rhs, "rober_rhs"), (jac, "rober_jac"), (tgrad, "rober_tgrad")


def pleiades_source(f32=False, name="pleiades_rhs", loops=False):
    """Pleiades right-hand side.  loops=True mirrors the reference's Julia loops
    (`for i in 1:7, j in 1:7; i != j`) — a small loop body that stays in the instruction cache;
    loops=False emits the same arithmetic as straight-line code (42 unrolled pairs).  Both forms
    perform identical operations in identical order."""
    if loops:
        T = _ty(f32)
        sq = "sqrtf" if f32 else "sqrt"
        suf = "f" if f32 else ""
        return ("void %s(%s* du, const %s* u, const %s* p, const %s t) {\n"
                "  for (int i = 0; i < 7; ++i) { du[i] = u[14 + i]; du[7 + i] = u[21 + i]; }\n"
                "  #pragma unroll 1\n"
                "  for (int i = 0; i < 7; ++i) {\n"
                "    %s ax = 0.0%s, ay = 0.0%s;\n"
                "    #pragma unroll 1\n"
                "    for (int j = 0; j < 7; ++j) {\n"
                "      if (j == i) continue;\n"
                "      %s dx = u[j] - u[i]; %s dy = u[7 + j] - u[7 + i];\n"
                "      %s r = %s(dx * dx + dy * dy); %s r3 = r * r * r;\n"
                "      %s m = (%s)(j + 1);\n"
                "      ax = ax + m * dx / r3; ay = ay + m * dy / r3;\n"
                "    }\n"
                "    du[14 + i] = ax; du[21 + i] = ay;\n"
                "  }\n"
                "}\n" % (name, T, T, T, T, T, suf, suf, T, T, T, sq, T, T, T)), name
    T = _ty(f32)
    sq = "sqrtf" if f32 else "sqrt"
    suf = "f" if f32 else ""
    L = ["void %s(%s* du, const %s* u, const %s* p, const %s t) {" % (name, T, T, T, T)]
    for i in range(7):
        L.append("  du[%d] = u[%d];" % (i, 14 + i))
    for i in range(7):
        L.append("  du[%d] = u[%d];" % (7 + i, 21 + i))
    L.append("  %s dx, dy, r, r3, ax, ay;" % T)
    for i in range(7):
        L.append("  ax = 0.0%s; ay = 0.0%s;" % (suf, suf))
        for j in range(7):
            if i == j:
                continue
            L.append("  dx = u[%d] - u[%d]; dy = u[%d] - u[%d];" % (j, i, 7 + j, 7 + i))
            L.append("  r = %s(dx * dx + dy * dy); r3 = r * r * r;" % sq)
            L.append("  ax = ax + %d.0%s * dx / r3; ay = ay + %d.0%s * dy / r3;" % (j + 1, suf, j + 1, suf))
        L.append("  du[%d] = ax; du[%d] = ay;" % (14 + i, 21 + i))
    L.append("}")
    return "\n".join(L) + "\n", name


def pleiades_pairs_source(f32=False, name="pleiades_rhs_pairs"):
    """Pleiades in full-vector form with every UNORDERED pair evaluated once (21 instead of 42 distance
    evaluations), for kernels that inline the RHS once as straight-line code (B200ODE_OPT_SMEM_STAGES).

    Bit-identical to pleiades_source / the reference loop (benchmark/benchmarks.jl:45-57): for a < b,
    dx_ba = x[a] - x[b] = -(x[b] - x[a]) exactly, so dx^2 + dy^2, r and r3 are the same floating-point numbers for (a, b)
    and (b, a), and (m dx_ba) / r3 = -((m dx_ab) / r3) exactly — body b subtracts what the reference adds negated.
    Pairs run in lexicographic order, which visits the partners of every body in ascending j (all (a, i) with a < i
    precede all (i, b)), so each acceleration accumulates from 0.0 in the reference's order.  The four quotients of a pair
    share the divisor r3; written with B200_DIV they also share the reciprocal refinement on the device."""
    T = _ty(f32)
    sq = "sqrtf" if f32 else "sqrt"
    suf = "f" if f32 else ""
    L = ["#ifndef B200_DIV", "#define B200_DIV(a, b) ((a) / (b))", "#endif",
         "void %s(%s* du, const %s* u, const %s* p, const %s t) {" % (name, T, T, T, T)]
    for i in range(14):
        L.append("  du[%d] = u[%d];" % (i, 14 + i))
    L.append("  %s %s;" % (T, ", ".join("ax%d = 0.0%s, ay%d = 0.0%s" % (i, suf, i, suf) for i in range(7))))
    for a in range(7):
        for b in range(a + 1, 7):
            L.append("  {")
            L.append("    const %s dx = u[%d] - u[%d], dy = u[%d] - u[%d];" % (T, b, a, 7 + b, 7 + a))
            L.append("    const %s r = %s(dx * dx + dy * dy); const %s r3 = r * r * r;" % (T, sq, T))
            L.append("    ax%d = ax%d + B200_DIV(%d.0%s * dx, r3); ay%d = ay%d + B200_DIV(%d.0%s * dy, r3);"
                     % (a, a, b + 1, suf, a, a, b + 1, suf))
            L.append("    ax%d = ax%d - B200_DIV(%d.0%s * dx, r3); ay%d = ay%d - B200_DIV(%d.0%s * dy, r3);"
                     % (b, b, a + 1, suf, b, b, a + 1, suf))
            L.append("  }")
    for i in range(7):
        L.append("  du[%d] = ax%d; du[%d] = ay%d;" % (14 + i, i, 21 + i, i))
    L.append("}")
    return "\n".join(L) + "\n", name


def pleiades_component_source(f32=False, name="pleiades_rhs_i"):
    """Pleiades in COMPONENT FORM for the lane-group kernel (B200ODE_OPT_COMPONENT_RHS): du_i as a function of the
    run-time index i.  Components 0..13 copy the velocities; components 14..27 are the accelerations of body
    b = (i - 14) % 7 along x (i < 21) or y, accumulated over the six other bodies j in ascending order with the
    reference's operations (dx, dy, r = sqrt(dx dx + dy dy), r3 = r r r, m dq / r3), so every du_i has the same bits
    as pleiades_source's.  The loop runs over k = 0..5 with j = k + (k >= b) — the reference's `j != i` skip without a
    branch — so the six terms sit in one basic block and overlap on the FP64 pipe; the mass m = j + 1 is formed by an
    exact floating-point increment instead of an int->double conversion per term."""
    T = _ty(f32)
    sq = "sqrtf" if f32 else "sqrt"
    suf = "f" if f32 else ""
    return ("#ifndef B200_DIV\n#define B200_DIV(a, b) ((a) / (b))\n#endif\n"
            "%(T)s %(name)s(int i, const %(T)s* u, const %(T)s* p, const %(T)s t) {\n"
            "  if (i < 14) return u[14 + i];\n"
            "  const int yaxis = (i >= 21);\n"
            "  const int b = i - (yaxis ? 21 : 14);\n"
            "  const %(T)s xb = u[b], yb = u[7 + b];\n"
            "  %(T)s acc = 0.0%(s)s;\n"
            "  for (int k = 0; k < 6; ++k) {\n"
            "    const int skip = (k >= b);\n"
            "    const int j = k + skip;\n"
            "    const %(T)s dx = u[j] - xb, dy = u[7 + j] - yb;\n"
            "    const %(T)s r = %(sq)s(dx * dx + dy * dy); const %(T)s r3 = r * r * r;\n"
            "    const %(T)s m = (%(T)s)(k + 1) + (skip ? 1.0%(s)s : 0.0%(s)s);\n"
            "    acc = acc + B200_DIV(m * (yaxis ? dy : dx), r3);\n"
            "  }\n"
            "  return acc;\n"
            "}\n" % dict(T=T, name=name, s=suf, sq=sq)), name


PLEIADES_U0 = np.array([3.0, 3.0, -1.0, -3.0, 2.0, -2.0, 2.0, 3.0, -3.0, 2.0, 0, 0, -4.0, 4.0,
                        0, 0, 0, 0, 0, 1.75, -1.5, 0, 0, 0, -1.25, 1, 0, 0], dtype=np.float64)


# ---- deterministic synthetic inputs ---------------------------------
def splitmix64_uniform(i, j):
    """U(i,j) = top 53 bits of SplitMix64(seed = 0x9E3779B97F4A7C15 xor (4 i + j)) / 2^53.

    i may be a numpy array of trajectory indices."""
    with np.errstate(over="ignore"):
        x = np.uint64(0x9E3779B97F4A7C15) ^ (np.asarray(i, dtype=np.uint64) * np.uint64(4) + np.uint64(j))
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def lorenz_params(N, offset=0, f32=False, sweep_total=None):
    """p[i] = (10, rho_i, 8/3); rho_i = 28 (0.5 + U(i,0)) in [14, 42), or the pure sweep
    rho_i = 14 + 28 i / sweep_total of config 5."""
    idx = np.arange(offset, offset + N, dtype=np.uint64)
    if sweep_total is None:
        rho = 28.0 * (0.5 + splitmix64_uniform(idx, 0))
    else:
        rho = 14.0 + 28.0 * idx.astype(np.float64) / float(sweep_total)
    p = np.empty((N, 3), dtype=np.float64)
    p[:, 0] = 10.0
    p[:, 1] = rho
    p[:, 2] = 8.0 / 3.0
    return p.astype(np.float32) if f32 else p


def robertson_params(N, offset=0, f32=False):
    idx = np.arange(offset, offset + N, dtype=np.uint64)
    base = np.array([0.04, 3.0e7, 1.0e4])
    p = np.empty((N, 3), dtype=np.float64)
    for j in range(3):
        p[:, j] = base[j] * (0.5 + splitmix64_uniform(idx, j))
    return p.astype(np.float32) if f32 else p


def pleiades_u0(N, offset=0, f32=False):
    """positions (components 0..13) += 0.01 (2 U(i,j) - 1)."""
    idx = np.arange(offset, offset + N, dtype=np.uint64)
    u0 = np.tile(PLEIADES_U0, (N, 1))
    for j in range(14):
        # U(i,j) is defined for j < 4 by the 4 i + j packing; use a disjoint stream per j
        u0[:, j] += 0.01 * (2.0 * splitmix64_uniform(idx * np.uint64(4) + np.uint64(j // 4), j % 4) - 1.0)
    return u0.astype(np.float32) if f32 else u0


# ---- stiff systems with n != 3: the LU branch of the Rosenbrock linear solve -----------------------
# Van der Pol   /root/reference/benchmark/benchmarks.jl:110-123   (n = 2, p = [mu])
# HIRES n = 5   /root/reference/test/InterfaceI/static_array_tests.jl:125-143  (SVector form, StaticWOperator inv path)
# HIRES n = 8   /root/reference/test/InterfaceI/static_array_tests.jl:145-167  (n > 7: StaticWOperator keeps lu(W))
# The reference solves these with AD Jacobians; here jac/tgrad are derived symbolically (codegen.py), the
# ODEFunction(f; jac, tgrad) branch of calc_J / calc_tderivative.  Rate constants are parameters so that
# prob_func can randomise them.
def _vdp(u, p, t):
    x, y = u
    return [y, p[0] * ((1 - x * x) * y - x)]


def _hires5(u, p, t):
    y1, y2, y3, y4, y5 = u
    return [-p[0] * y1 + 0.43 * y2 + 8.32 * y3 + 0.0007,
            p[0] * y1 - 8.75 * y2,
            -10.03 * y3 + 0.43 * y4 + 0.035 * y5,
            8.32 * y2 + p[0] * y3 - 1.12 * y4,
            -1.745 * y5 + 0.43 * y2 + 0.43 * y4]


def _hires8(u, p, t):
    y1, y2, y3, y4, y5, y6, y7, y8 = u
    return [-p[0] * y1 + 0.43 * y2 + 8.32 * y3 + 0.0007,
            p[0] * y1 - 8.75 * y2,
            -10.03 * y3 + 0.43 * y4 + 0.035 * y5,
            8.32 * y2 + p[0] * y3 - 1.12 * y4,
            -1.745 * y5 + 0.43 * y6 + 0.43 * y7,
            -p[1] * y6 * y8 + 0.69 * y4 + p[0] * y5 - 0.43 * y6 + 0.69 * y7,
            p[1] * y6 * y8 - 1.81 * y7,
            -p[1] * y6 * y8 + 1.81 * y7]


def _chain16(u, p, t):
    """Stiff reaction-diffusion chain, n = 16: u_i' = D (u_{i-1} - 2 u_i + u_{i+1}) - k u_i^3 + s_i with fixed ends
    (D = p[0], k = p[1]); a tridiagonal Jacobian with eigenvalues down to about -4 D."""
    n = 16
    out = []
    for i in range(n):
        left = u[i - 1] if i > 0 else 0
        right = u[i + 1] if i < n - 1 else 0
        src = 1.0 if i == 0 else 0
        out.append(p[0] * (left - 2 * u[i] + right) - p[1] * u[i] * u[i] * u[i] + src)
    return out


STIFF_PROBLEMS = {
    # name: (f, n, np, u0, tspan, nominal p)
    "vdp": (_vdp, 2, 1, [1.0, 1.0], (0.0, 6.3), [1.0e3]),
    "hires5": (_hires5, 5, 1, [1.0, 0.0, 0.0, 0.0, 0.0057], (0.0, 321.8122), [1.71]),
    "chain16": (_chain16, 16, 2, [0.0] * 16, (0.0, 10.0), [400.0, 1.0]),
    "hires8": (_hires8, 8, 2, [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0057], (0.0, 321.8122), [1.71, 280.0]),
}


def stiff_sources(name, f32=False):
    """(rhs, jac, tgrad) C sources (each (text, function name)), n, np, u0, tspan for a STIFF_PROBLEMS entry."""
    from . import codegen
    f, n, np_, u0, tspan, _ = STIFF_PROBLEMS[name]
    rhs = codegen.build_function_c(f, n, np_, fname=name + "_rhs", f32=f32)
    jac = codegen.build_jacobian_c(f, n, np_, fname=name + "_jac", f32=f32)
    tg = codegen.build_tgrad_c(f, n, np_, fname=name + "_tgrad", f32=f32)
    return rhs, jac, tg, n, np_, np.asarray(u0, dtype=np.float32 if f32 else np.float64), tspan


def stiff_params(name, N, offset=0, f32=False):
    """p[i][j] = nominal_j (0.5 + U(i, j))."""
    nominal = STIFF_PROBLEMS[name][5]
    idx = np.arange(offset, offset + N, dtype=np.uint64)
    p = np.empty((N, len(nominal)), dtype=np.float64)
    for j, v in enumerate(nominal):
        p[:, j] = v * (0.5 + splitmix64_uniform(idx, j))
    return p.astype(np.float32) if f32 else p


def stiff_component_sources(name, f32=False):
    """Component-form sources of a STIFF_PROBLEMS entry for the lane-group Rosenbrock23 (B200ODE_OPT_COMPONENT_RHS):
    (rhs_i, jac_ij, tgrad_i or None), n, np, u0, tspan — the same expressions as stiff_sources, behind an index."""
    from . import codegen
    f, n, np_, u0, tspan, _ = STIFF_PROBLEMS[name]
    rhs = codegen.build_function_component_c(f, n, np_, fname=name + "_rhs_i", f32=f32)
    jac = codegen.build_jacobian_entry_c(f, n, np_, fname=name + "_jac_ij", f32=f32)
    tg = codegen.build_tgrad_component_c(f, n, np_, fname=name + "_tgrad_i", f32=f32)
    return rhs, jac, tg, n, np_, np.asarray(u0, dtype=np.float32 if f32 else np.float64), tspan
