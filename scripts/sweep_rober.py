import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
N = 1 << 20
(r, j, tg) = pl.robertson_sources(False); p = pl.robertson_params(N); u0 = np.array([1.0, 0, 0])
for alg, name in ((pkg.ALG_RODAS5P, "rodas5p"), (pkg.ALG_ROSENBROCK23, "ros23")):
    for v in ["-DB200_MINBLOCKS=2", "-DB200_MINBLOCKS=3", "-DB200_MINBLOCKS=4", "-DB200_MINBLOCKS=5", "-DB200_BLOCK=256 -DB200_MINBLOCKS=1", "-DB200_BLOCK=512 -DB200_MINBLOCKS=1"]:
        prog = h.compile(alg, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1], extra_options=v)
        best = min(ll.solve_host(prog, u0, p, (0.0, 1e5), reltol=1e-6, abstol=1e-8)["kernel_ms"] for _ in range(3))
        print(name, v, "regs", prog.info["regs_integrate"], "local", prog.info["local_bytes_integrate"], "blocks/SM", prog.info["blocks_per_sm"],
              "kernel_ms %.3f -> %.1f M traj/s" % (best, N / best / 1e3), flush=True)
        prog.close()
