"""ncu driver for the non-Lorenz workloads: python scripts/prof_generic.py {pleiades|rodas5p|ros23} N"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
which = sys.argv[1]; N = int(sys.argv[2]); extra = sys.argv[3] if len(sys.argv) > 3 else None
h = pkg.Handle(0)
if which == "pleiades":
    rhs = pl.pleiades_source(False); u0 = pl.pleiades_u0(N)
    prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, rhs[0], rhs[1], extra_options=extra)
    run = lambda: ll.solve_host(prog, u0, None, (0.0, 3.0), reltol=1e-6, abstol=1e-8)
else:
    (r, j, tg) = pl.robertson_sources(False); p = pl.robertson_params(N)
    alg = pkg.ALG_RODAS5P if which == "rodas5p" else pkg.ALG_ROSENBROCK23
    prog = h.compile(alg, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1], extra_options=extra)
    run = lambda: ll.solve_host(prog, np.array([1.0, 0, 0]), p, (0.0, 1e5), reltol=1e-6, abstol=1e-8)
print(prog.info)
for _ in range(2):
    g = run()
    print("kernel_ms", g["kernel_ms"], "steps/traj", float((g["naccept"] + g["nreject"]).mean()))
