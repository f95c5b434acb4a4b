import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
N = 1 << 20
u0 = np.array([1.0, 0, 0])
rhs = pl.lorenz_source(True); p = pl.lorenz_params(N, f32=True)
for v in ["-DB200_MINBLOCKS=4", "-DB200_MINBLOCKS=5", "-DB200_MINBLOCKS=6", "-DB200_MINBLOCKS=8", "-DB200_BLOCK=256 -DB200_MINBLOCKS=3", "-DB200_BLOCK=512 -DB200_MINBLOCKS=1"]:
    prog = h.compile(pkg.ALG_TSIT5, pkg.F32, 3, 3, rhs[0], rhs[1], extra_options=v)
    res = []
    for kw in (dict(), dict(saveat=np.arange(1, 101) / 10.0)):
        res.append(round(min(ll.solve_host(prog, u0, p, (0.0, 10.0), **kw)["kernel_ms"] for _ in range(3)), 3))
    print("f32", v, "regs", prog.info["regs_integrate"], "local", prog.info["local_bytes_integrate"], "blocks/SM", prog.info["blocks_per_sm"], res, flush=True)
    prog.close()
