"""compute-sanitizer runs for the widened paths: ragged save_everystep (count + fill passes), dense evaluation,
save_idxs rows, and the added steppers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
u0 = np.array([1.0, 0, 0])
N = 333
s, n = pl.lorenz_source(False); p = pl.lorenz_params(N)
E = pkg._lib.OPT_EVERYSTEP
tq = np.linspace(0.0, 1.0, 7)
for alg in (pkg.ALG_TSIT5, pkg.ALG_DP5, pkg.ALG_BS3, pkg.ALG_VERN6, pkg.ALG_VERN7, pkg.ALG_VERN8, pkg.ALG_VERN9):
    prog = h.compile(alg, pkg.F64, 3, 3, s, n, extra_options=E)
    g = ll.solve_host_everystep(prog, u0, p, (0.0, 1.0), saveat=[0.3, 0.9], maxiters=30)
    d = ll.solve_host_dense(prog, u0, p, (0.0, 1.0), tq)
    prog2 = h.compile(alg, pkg.F64, 3, 3, s, n, extra_options=pkg._lib.opt_save_idxs([2, 0]))
    r = ll.solve_host(prog2, u0, p, (0.0, 1.0), saveat=[0.5, 1.0])
    print(alg, int(g["row_offsets"][-1]), d["dense"].shape, r["us"].shape)
prog3 = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, n, extra_options=E + " " + pkg._lib.opt_save_idxs([1]))
g = ll.solve_host_everystep(prog3, u0, p, (0.0, 1.0))
print("ragged save_idxs", g["us"].shape)
(r_, j, tg) = pl.robertson_sources(False); k = pl.robertson_params(N)
for alg in (pkg.ALG_ROSENBROCK23, pkg.ALG_RODAS5P, pkg.ALG_RODAS5, pkg.ALG_RODAS4):
    prog = h.compile(alg, pkg.F64, 3, 3, r_[0], r_[1], j[0], j[1], tg[0], tg[1], extra_options=E)
    g = ll.solve_host_everystep(prog, u0, k, (0.0, 10.0), reltol=1e-6, abstol=1e-8)
    d = ll.solve_host_dense(prog, u0, k, (0.0, 10.0), tq * 10, reltol=1e-6, abstol=1e-8)
    print(alg, int(g["row_offsets"][-1]), d["dense"].shape)
print("done")
