"""GPU-vs-oracle parity over all four algorithms (small N), plus timings at larger N."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
from oracle import oracle
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)

def compare(tag, g, o, f32=False):
    ok = {k: bool(np.array_equal(g[k], o[k])) for k in ("naccept", "nreject", "nf", "retcode", "nsaved", "njacs", "nw", "nsolve")}
    view = np.uint32 if f32 else np.uint64
    ok["u_final_bits"] = bool(np.array_equal(g["u_final"].view(view), o["u_final"].view(view)))
    ok["t_final"] = bool(np.array_equal(g["t_final"], o["t_final"]))
    if o.get("us") is not None:
        ok["us_bits"] = bool(np.array_equal(g["us"].view(view), o["us"].view(view)))
    bad = [k for k, v in ok.items() if not v]
    print(tag, "OK" if not bad else "MISMATCH %s" % bad, "kernel_ms %.3f" % g["kernel_ms"], "naccept mean %.1f nreject mean %.2f rc!=1: %d"
          % (g["naccept"].mean(), g["nreject"].mean(), int((g["retcode"] != 1).sum())), flush=True)
    if bad:
        d = np.abs(g["u_final"].astype(np.float64) - o["u_final"].astype(np.float64))
        print("   max|du_final| %.3e; naccept differs at %d traj" % (np.nanmax(d), int((g["naccept"] != o["naccept"]).sum())))
    return not bad

allok = True
N = 4096
for f32 in (False, True):
    dt_ = pkg.F32 if f32 else pkg.F64
    # Lorenz Tsit5
    rhs = pl.lorenz_source(f32); p = pl.lorenz_params(N, f32=f32); u0 = np.array([1.0, 0, 0])
    prog = h.compile(pkg.ALG_TSIT5, dt_, 3, 3, rhs[0], rhs[1])
    for kw in (dict(), dict(saveat=np.arange(1, 101) / 10.0), dict(reltol=1e-8 if not f32 else 1e-5)):
        g = ll.solve_host(prog, u0, p, (0.0, 10.0), **kw)
        o = oracle.solve(oracle.ALG_TSIT5, rhs, u0, p, (0.0, 10.0), 3, 3, f32=f32, **kw)
        allok &= compare("lorenz tsit5 f32=%d %s" % (f32, list(kw)), g, o, f32)
    # Lorenz Vern7
    prog = h.compile(pkg.ALG_VERN7, dt_, 3, 3, rhs[0], rhs[1])
    for kw in (dict(), dict(saveat=np.arange(1, 21) / 2.0)):
        g = ll.solve_host(prog, u0, p, (0.0, 10.0), **kw)
        o = oracle.solve(oracle.ALG_VERN7, rhs, u0, p, (0.0, 10.0), 3, 3, f32=f32, **kw)
        allok &= compare("lorenz vern7 f32=%d %s" % (f32, list(kw)), g, o, f32)
    # Robertson
    (r, j, tg) = pl.robertson_sources(f32); p = pl.robertson_params(N, f32=f32)
    for alg, oalg, name in ((pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23, "ros23"), (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P, "rodas5p")):
        prog = h.compile(alg, dt_, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1])
        tol = dict(reltol=1e-6, abstol=1e-8) if not f32 else dict(reltol=1e-3, abstol=1e-5)
        tf = 1e5 if not f32 else 1e3
        for kw in (dict(), dict(saveat=np.array([1.0, 10.0, 100.0, 1000.0]) * (tf / 1e3))):
            kw = dict(kw, **tol)
            g = ll.solve_host(prog, u0, p, (0.0, tf), **kw)
            o = oracle.solve(oalg, r, u0, p, (0.0, tf), 3, 3, f32=f32, jac=j, tgrad=tg, **kw)
            allok &= compare("rober %s f32=%d %s" % (name, f32, list(kw)), g, o, f32)
# Pleiades Vern7 f64
Np = 1024
rhs = pl.pleiades_source(False); u0 = pl.pleiades_u0(Np)
prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, rhs[0], rhs[1])
print("pleiades program", prog.info)
for kw in (dict(reltol=1e-6, abstol=1e-8), dict(reltol=1e-6, abstol=1e-8, saveat=np.array([0.5, 1.0, 1.5, 2.0, 2.5, 3.0]))):
    g = ll.solve_host(prog, u0, None, (0.0, 3.0), **kw)
    o = oracle.solve(oracle.ALG_VERN7, rhs, u0, None, (0.0, 3.0), 28, 0, **kw)
    allok &= compare("pleiades vern7 %s" % list(kw), g, o)
print("ALL OK" if allok else "SOME MISMATCH")

# timings
N = 1 << 20
u0 = np.array([1.0, 0, 0])
for f32 in (False, True):
    rhs = pl.lorenz_source(f32); p = pl.lorenz_params(N, f32=f32)
    prog = h.compile(pkg.ALG_TSIT5, pkg.F32 if f32 else pkg.F64, 3, 3, rhs[0], rhs[1])
    for kw in (dict(), dict(saveat=np.arange(1, 101) / 10.0)):
        best = min(ll.solve_host(prog, u0, p, (0.0, 10.0), **kw)["kernel_ms"] for _ in range(3))
        print("TIMING lorenz tsit5 f32=%d N=%d %s kernel_ms %.3f -> %.1f M traj/s" % (f32, N, list(kw), best, N / best / 1e3), flush=True)
(r, j, tg) = pl.robertson_sources(False); p = pl.robertson_params(N)
for alg, name in ((pkg.ALG_ROSENBROCK23, "ros23"), (pkg.ALG_RODAS5P, "rodas5p")):
    prog = h.compile(alg, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1])
    best = min(ll.solve_host(prog, u0, p, (0.0, 1e5), reltol=1e-6, abstol=1e-8)["kernel_ms"] for _ in range(2))
    print("TIMING rober %s N=%d kernel_ms %.3f -> %.1f M traj/s" % (name, N, best, N / best / 1e3), flush=True)
Np = 1 << 18
rhs = pl.pleiades_source(False); u0 = pl.pleiades_u0(Np)
prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, rhs[0], rhs[1])
best = min(ll.solve_host(prog, u0, None, (0.0, 3.0), reltol=1e-6, abstol=1e-8)["kernel_ms"] for _ in range(2))
print("TIMING pleiades vern7 N=%d kernel_ms %.3f -> %.3f M traj/s" % (Np, best, Np / best / 1e3), flush=True)
