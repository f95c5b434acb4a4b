"""Small runs of every algorithm for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
u0 = np.array([1.0, 0, 0])
N = 777
rhs = pl.lorenz_source(False); p = pl.lorenz_params(N)
for alg in (pkg.ALG_TSIT5, pkg.ALG_VERN7):
    prog = h.compile(alg, pkg.F64, 3, 3, rhs[0], rhs[1])
    for flags in (0, 1):
        g = ll.solve_host(prog, u0, p, (0.0, 2.0), saveat=[0.5, 1.0, 1.7], flags=flags, maxiters=40)
        print(alg, flags, g["retcode"][:5], g["nsaved"][:5])
(r, j, tg) = pl.robertson_sources(False); k = pl.robertson_params(N)
for alg in (pkg.ALG_ROSENBROCK23, pkg.ALG_RODAS5P):
    prog = h.compile(alg, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1])
    g = ll.solve_host(prog, u0, k, (0.0, 100.0), saveat=[10.0, 50.0], reltol=1e-6, abstol=1e-8)
    print(alg, g["retcode"][:5], g["nsaved"][:5])
rhs = pl.pleiades_source(False); up = pl.pleiades_u0(65)
prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, rhs[0], rhs[1])
g = ll.solve_host(prog, up, None, (0.0, 0.5), saveat=[0.25], reltol=1e-6, abstol=1e-8)
print("pleiades", g["retcode"][:5])
print("done")
