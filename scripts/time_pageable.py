"""Host-call timing with pageable (plain numpy) vs pinned result buffers: 2^20 Lorenz trajectories, saveat=0.1."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
N = 1 << 20
h = pkg.Handle(0)
s, n = pl.lorenz_source()
p = pl.lorenz_params(N)
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, n)
grid = pkg.ranges.saveat_grid(0.1, (0.0, 10.0))
out = {}
buf = {}
for rep in range(3):
    t0 = time.perf_counter()
    g = ll.solve_host(prog, np.array([1.0, 0, 0]), p, (0.0, 10.0), saveat=grid, out=buf)
    wall = (time.perf_counter() - t0) * 1e3
out["pageable"] = dict(kernel_ms=g["kernel_ms"], total_ms=g["total_ms"], wall_ms=wall)
us = torch.empty((N, 101, 3), dtype=torch.float64).pin_memory()
buf2 = {"us": us.numpy()}
for rep in range(3):
    t0 = time.perf_counter()
    g2 = ll.solve_host(prog, np.array([1.0, 0, 0]), p, (0.0, 10.0), saveat=grid, out=buf2)
    wall = (time.perf_counter() - t0) * 1e3
out["pinned"] = dict(kernel_ms=g2["kernel_ms"], total_ms=g2["total_ms"], wall_ms=wall)
out["same"] = bool(np.array_equal(g["us"], g2["us"]))
print(json.dumps(out, indent=1))
