"""racecheck target: the shared-memory (sliced) Vern7 kernel on a small Pleiades ensemble."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
src, name = pl.pleiades_source(False)
for opt in ("-DB200_SLICED=1", "-DB200_SLICED=1 -DB200_G=7 -DB200_K=2"):
    prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, src, name, extra_options=opt)
    g = ll.solve_host(prog, pl.pleiades_u0(70), None, (0.0, 0.2), saveat=[0.05, 0.1], reltol=1e-6, abstol=1e-8)
    print(opt, g["retcode"][:4], g["nsaved"][:4])
print("done")
