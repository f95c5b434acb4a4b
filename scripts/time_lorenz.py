"""Kernel timing for the Lorenz workloads (1 Mi trajectories)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
N = 1 << 20
u0 = np.array([1.0, 0, 0])
extra = sys.argv[1] if len(sys.argv) > 1 else None
for f32 in (False, True):
    rhs = pl.lorenz_source(f32); p = pl.lorenz_params(N, f32=f32)
    prog = h.compile(pkg.ALG_TSIT5, pkg.F32 if f32 else pkg.F64, 3, 3, rhs[0], rhs[1], extra_options=extra)
    for kw in (dict(), dict(saveat=np.arange(1, 101) / 10.0)):
        best = min(ll.solve_host(prog, u0, p, (0.0, 10.0), **kw)["kernel_ms"] for _ in range(4))
        print("TIMING lorenz tsit5 f32=%d N=%d %s regs %d kernel_ms %.3f -> %.1f M traj/s" % (f32, N, list(kw), prog.info["regs_integrate"], best, N / best / 1e3), flush=True)
