"""ncu driver: Pleiades/Vern7, 2^16 trajectories, lane-group kernel (default) or any extra options given as argv[1]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
opt = sys.argv[1] if len(sys.argv) > 1 else "-DB200_COOP=1"
coop = "B200_COOP=1" in opt
N = 1 << 16
h = pkg.Handle(0)
wide = "B200_WIDE=1" in opt
src = pl.pleiades_component_source() if coop else (pl.pleiades_pairs_source() if wide else pl.pleiades_source(False, loops=True))
prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, src[0], src[1], extra_options=opt)
u0 = pl.pleiades_u0(N)
for _ in range(2):
    g = ll.solve_host(prog, u0, None, (0.0, 3.0), reltol=1e-6, abstol=1e-8)
print("ok", g["kernel_ms"], int(g["naccept"].sum()))
