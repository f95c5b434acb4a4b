"""Small runs of the paths added in the second half of round 2, for compute-sanitizer (memcheck / racecheck / synccheck /
initcheck): callbacks (ragged + rectangular), the composite algorithm, Rodas3P/23W, vector tolerances, isoutofdomain, the
lane-group Rosenbrock23 (shared memory + warp shuffles), the lane-group Vern7."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import b200_import
from helpers import ball_sources
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
U0 = np.array([1.0, 0, 0])
N = 333
# callbacks
rhs, cond, bounce, stop = ball_sources()
idx = np.arange(N, dtype=np.uint64)
p = np.stack([9.81 * (0.5 + pl.splitmix64_uniform(idx, 0)), 0.8 + 0.2 * pl.splitmix64_uniform(idx, 1)], axis=1)
cb = dict(kind="continuous", condition=cond, affect=None, affect_neg=bounce)
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 2, 2, rhs[0], rhs[1], extra_options=pkg._lib.OPT_EVERYSTEP, callbacks=[cb])
g = ll.solve_host_everystep(prog, np.array([50.0, 0.0]), p, (0.0, 8.0), saveat=[1.0, 4.0])
print("callbacks ragged", g["retcode"][:4], g["nsaved"][:4])
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 2, 2, rhs[0], rhs[1], callbacks=[dict(cb, save_positions=(False, False)),
                                                                         dict(kind="continuous", condition=cond, affect=stop, save_positions=(False, False))])
g = ll.solve_host(prog, np.array([50.0, 0.0]), p, (0.0, 8.0), saveat=[1.0, 4.0])
print("callbacks rectangular", g["retcode"][:4], g["nsaved"][:4])
# composite, Rodas3P / Rodas23W
r, j, tg, n, np_, u0v, _ = pl.stiff_sources("vdp")
mu = (0.5 * (1000.0 ** pl.splitmix64_uniform(idx, 0))).reshape(N, 1)
prog = h.compile(pkg.ALG_AUTOTSIT5_ROSENBROCK23, pkg.F64, n, np_, r[0], r[1], j[0], j[1], tg[0], tg[1])
g = ll.solve_host(prog, u0v, mu, (0.0, 10.0), saveat=[2.0, 5.0])
print("autotsit5", g["retcode"][:4], int((g["njacs"] > 0).sum()))
rr, jj, tt = pl.robertson_sources(); k = pl.robertson_params(N)
for alg in (pkg.ALG_RODAS3P, pkg.ALG_RODAS23W):
    prog = h.compile(alg, pkg.F64, 3, 3, rr[0], rr[1], jj[0], jj[1], tt[0], tt[1])
    g = ll.solve_host(prog, U0, k, (0.0, 100.0), saveat=[10.0, 50.0], reltol=1e-6, abstol=1e-8)
    print(alg, g["retcode"][:4])
# vector tolerances, isoutofdomain
s, nm = pl.lorenz_source(False); lp = pl.lorenz_params(N)
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, nm, extra_options=pkg._lib.OPT_VECTOR_TOL)
g = ll.solve_host(prog, U0, lp, (0.0, 2.0), reltol=[1e-5, 1e-3, 1e-4], abstol=[1e-7, 1e-4, 1e-6])
print("vector tol", g["retcode"][:4])
out = ("double lz_out(const double* u, const double* p, const double t) { return u[2] > 45.0 ? 1.0 : 0.0; }\n", "lz_out")
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, nm, callbacks=[dict(kind="isoutofdomain", condition=out)])
g = ll.solve_host(prog, U0, lp, (0.0, 2.0), maxiters=500)
print("isoutofdomain", np.bincount(g["retcode"]))
# lane-group Rosenbrock23 (n = 5: idle lanes in the group; n = 8) and lane-group Vern7
for name in ("hires5", "hires8"):
    rc, jc, tc, n, np_, u0v, tspan = pl.stiff_component_sources(name)
    prog = h.compile(pkg.ALG_ROSENBROCK23, pkg.F64, n, np_, rc[0], rc[1], jc[0], jc[1], tc[0] if tc else None, tc[1] if tc else None,
                     extra_options=pkg._lib.OPT_COMPONENT_RHS)
    g = ll.solve_host(prog, u0v, pl.stiff_params(name, 61), (0.0, 5.0), saveat=[1.0], reltol=1e-5, abstol=1e-7)
    print("lane-group ros23", name, g["retcode"][:4])
cs = pl.pleiades_component_source(False)
prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, cs[0], cs[1], extra_options=pkg._lib.OPT_COMPONENT_RHS)
g = ll.solve_host(prog, pl.pleiades_u0(37), None, (0.0, 0.3), saveat=[0.1], reltol=1e-6, abstol=1e-8)
print("lane-group vern7", g["retcode"][:4])
print("done")
