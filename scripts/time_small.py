"""BASELINE configs[0] (Lorenz / Tsit5, 10 k trajectories, reltol 1e-8) and other small ensembles: device time per solve."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
rhs = pl.lorenz_source(False)
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
u0 = np.array([1.0, 0, 0])
for N in (1000, 10000, 50000, 200000):
    p = pl.lorenz_params(N)
    best = min(ll.solve_host(prog, u0, p, (0.0, 10.0), reltol=1e-8)["kernel_ms"] for _ in range(5))
    print("N=%d reltol=1e-8 kernel_ms %.3f -> %.2f M traj/s" % (N, best, N / best / 1e3), flush=True)
