"""Small driver for ncu: Lorenz/Tsit5 ensemble, N trajectories, optional saveat."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
N = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
saveat = np.arange(1, 101) / 10.0 if (len(sys.argv) > 2 and sys.argv[2] == "saveat") else None
f32 = len(sys.argv) > 3 and sys.argv[3] == "f32"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
h = pkg.Handle(0)
rhs = pl.lorenz_source(f32)
prog = h.compile(pkg.ALG_TSIT5, pkg.F32 if f32 else pkg.F64, 3, 3, rhs[0], rhs[1])
p = pl.lorenz_params(N, f32=f32)
u0 = np.array([1.0, 0.0, 0.0])
for r in range(reps):
    g = ll.solve_host(prog, u0, p, (0.0, 10.0), saveat=saveat)
    print("kernel_ms", g["kernel_ms"])
