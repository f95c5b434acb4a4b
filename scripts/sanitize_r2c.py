"""Small runs of the paths added in the last third of round 2, for compute-sanitizer (memcheck / racecheck / synccheck /
initcheck): the shared-memory stage kernel (Vern7, dynamic shared memory, disabled lanes), the staged saveat queue
(per-warp shared-memory rings, warp-wide drains), the regrouped Rosenbrock divisions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
U0 = np.array([1.0, 0, 0])
# shared-memory stage kernel: Pleiades (112 of 128 threads own a trajectory), Lorenz (512 per CTA), FP32
ps = pl.pleiades_pairs_source(False)
prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, ps[0], ps[1], extra_options=pkg._lib.OPT_SMEM_STAGES)
g = ll.solve_host(prog, pl.pleiades_u0(300), None, (0.0, 0.5), reltol=1e-6, abstol=1e-8, saveat=[0.5])
print("smem stages pleiades", g["retcode"][:4], int(g["naccept"].sum()))
ps32 = pl.pleiades_pairs_source(True)
prog = h.compile(pkg.ALG_VERN7, pkg.F32, 28, 0, ps32[0], ps32[1], extra_options=pkg._lib.OPT_SMEM_STAGES)
g = ll.solve_host(prog, pl.pleiades_u0(300, f32=True), None, (0.0, 0.5), reltol=1e-4, abstol=1e-5)
print("smem stages pleiades f32", g["retcode"][:4])
s, nm = pl.lorenz_source(False); lp = pl.lorenz_params(700)
prog = h.compile(pkg.ALG_VERN7, pkg.F64, 3, 3, s, nm, extra_options=pkg._lib.OPT_SMEM_STAGES)
g = ll.solve_host(prog, U0, lp, (0.0, 2.0))
print("smem stages lorenz", g["retcode"][:4])
# staged saveat queue: headline grid, many rows per step, failures, static schedule, FP32
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, nm, extra_options=pkg._lib.OPT_STAGED_SAVEAT)
for kw in (dict(saveat=[0.1 * k for k in range(1, 31)]), dict(saveat=[0.005 * k for k in range(1, 201)]),
           dict(saveat=[0.5, 1.0, 1.5], maxiters=12), dict(saveat=[0.25 * k for k in range(1, 9)], flags=pkg._lib.FLAG_STATIC_SCHEDULE)):
    g = ll.solve_host(prog, U0, lp, (0.0, 3.0 if len(kw["saveat"]) != 200 else 1.0), **kw)
    print("staged saveat", len(kw["saveat"]), np.bincount(g["retcode"]), int(g["nsaved"].sum()))
s32, nm32 = pl.lorenz_source(True)
prog = h.compile(pkg.ALG_TSIT5, pkg.F32, 3, 3, s32, nm32, extra_options=pkg._lib.OPT_STAGED_SAVEAT)
g = ll.solve_host(prog, U0.astype(np.float32), pl.lorenz_params(700, f32=True), (0.0, 3.0), saveat=[0.1 * k for k in range(1, 31)])
print("staged saveat f32", np.bincount(g["retcode"]))
# Rosenbrock steppers with the grouped divisions (n = 3 inverse, Rodas dtC), incl. a singular W (dt -> huge)
rr, jj, tt = pl.robertson_sources(); k = pl.robertson_params(333)
for alg in (pkg.ALG_ROSENBROCK23, pkg.ALG_RODAS5P, pkg.ALG_RODAS4, pkg.ALG_ROSENBROCK32):
    prog = h.compile(alg, pkg.F64, 3, 3, rr[0], rr[1], jj[0], jj[1], tt[0], tt[1])
    g = ll.solve_host(prog, U0, k, (0.0, 100.0), saveat=[10.0, 50.0], reltol=1e-6, abstol=1e-8, maxiters=2000)
    print("rosenbrock", alg, np.bincount(g["retcode"]))
# d_discontinuities, per-trajectory spans (final states + ragged rows), the lazily built no-saveat variant, shrunken CTAs
fs = ("void forced(double* du, const double* u, const double* p, const double t) { du[0] = -p[0] * u[0] + (t > 1.0 ? p[1] : 0.0); du[1] = u[0] - u[1]; }\n", "forced")
fp = np.stack([0.5 + pl.splitmix64_uniform(np.arange(333, dtype=np.uint64), 0), 1.0 + pl.splitmix64_uniform(np.arange(333, dtype=np.uint64), 1)], axis=1)
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 2, 2, fs[0], fs[1], extra_options=pkg._lib.OPT_TSTOPS)
g = ll.solve_host(prog, np.array([1.0, 0.0]), fp, (0.0, 3.0), d_discontinuities=[0.0, 1.0], saveat=[0.5, 1.0, 2.0], tstops=[2.5])
print("d_discontinuities", np.bincount(g["retcode"]), int(g["nsaved"].sum()))
spans = np.stack([np.zeros(700), 0.5 + 2.0 * pl.splitmix64_uniform(np.arange(700, dtype=np.uint64), 3)], axis=1)
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, nm, extra_options=pkg._lib.OPT_TSPANS)
g = ll.solve_host(prog, U0, lp, spans)
print("tspans final", np.bincount(g["retcode"]))
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, nm, extra_options=pkg._lib.OPT_TSPANS + " " + pkg._lib.OPT_EVERYSTEP)
g = ll.solve_host_everystep(prog, U0, lp, spans, saveat=[0.25, 1.0, 2.0])
print("tspans ragged", int(g["row_offsets"][-1]))
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, nm)
for N in (1, 33, 700):      # no saveat: the second build; few trajectories: CTAs of 32..64 threads
    g = ll.solve_host(prog, U0, lp[:N], (0.0, 2.0), save_start=False)
    g = ll.solve_host(prog, U0, lp[:N], (0.0, 2.0), saveat=[1.0, 2.0])
    print("nosave variant / small CTAs", N, np.bincount(g["retcode"]))
print("done")
