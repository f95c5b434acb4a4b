"""Timing of the save_everystep (ragged) and dense-output paths on 2^20 Lorenz trajectories, tspan (0, 10).
kernel_ms = device time of count pass + fill pass (+ dense evaluation); total_ms includes H2D/D2H."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
N = 1 << 20
h = pkg.Handle(0)
out = {}
for f32 in (False, True):
    s, n = pl.lorenz_source(f32)
    p = pl.lorenz_params(N, f32=f32)
    prog = h.compile(pkg.ALG_TSIT5, pkg.F32 if f32 else pkg.F64, 3, 3, s, n, extra_options=pkg._lib.OPT_EVERYSTEP)
    tag = "f32" if f32 else "f64"
    for rep in range(3):
        g = ll.solve_host_everystep(prog, np.array([1.0, 0, 0]), p, (0.0, 10.0))
    rows = int(g["row_offsets"][-1])
    out["everystep_" + tag] = dict(kernel_ms=g["kernel_ms"], total_ms=g["total_ms"], rows=rows,
                                   rows_per_traj=rows / N, bytes_rows=rows * (3 + 2) * (4 if f32 else 8))
    tq = np.linspace(0.0, 10.0, 101)
    for rep in range(3):
        d = ll.solve_host_dense(prog, np.array([1.0, 0, 0]), p, (0.0, 10.0), tq)
    out["dense101_" + tag] = dict(kernel_ms=d["kernel_ms"], total_ms=d["total_ms"])
print(json.dumps(out, indent=1))
