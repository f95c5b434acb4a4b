"""Sweep launch bounds (registers/occupancy) for Lorenz/Tsit5 and report kernel time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
N = 1 << 20
h = pkg.Handle(0)
f32 = len(sys.argv) > 1 and sys.argv[1] == "f32"
rhs = pl.lorenz_source(f32)
p = pl.lorenz_params(N, f32=f32)
u0 = np.array([1.0, 0.0, 0.0])
variants = [("-DB200_BLOCK=128 -DB200_MINBLOCKS=4"), ("-DB200_BLOCK=128 -DB200_MINBLOCKS=5"),
            ("-DB200_BLOCK=128 -DB200_MINBLOCKS=6"), ("-DB200_BLOCK=64 -DB200_MINBLOCKS=10"),
            ("-DB200_BLOCK=64 -DB200_MINBLOCKS=12"), ("-DB200_BLOCK=256 -DB200_MINBLOCKS=2"),
            ("-DB200_BLOCK=128 -DB200_MINBLOCKS=3"), ("-DB200_BLOCK=128 -DB200_MINBLOCKS=8")]
for v in variants:
    prog = h.compile(pkg.ALG_TSIT5, pkg.F32 if f32 else pkg.F64, 3, 3, rhs[0], rhs[1], extra_options=v)
    res = []
    for saveat in (None, np.arange(1, 101) / 10.0):
        best = 1e9
        for rep in range(3):
            g = ll.solve_host(prog, u0, p, (0.0, 10.0), saveat=saveat)
            best = min(best, g["kernel_ms"])
        res.append(round(best, 3))
    print(v, "regs", prog.info["regs_integrate"], "local", prog.info["local_bytes_integrate"], "blocks/SM",
          prog.info["blocks_per_sm"], "kernel_ms (final-only, saveat)", res, flush=True)
    prog.close()
