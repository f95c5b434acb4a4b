"""One-thread vs lane-group Rosenbrock23 on the stiff systems with n = 5, 8 (device-resident timing through solve_host's
kernel_ms): python scripts/time_stiff_lanegroup.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
N = 1 << 18
for name in ("hires5", "hires8", "chain16"):
    r, j, tg, n, np_, u0, tspan = pl.stiff_sources(name)
    rc, jc, tc = pl.stiff_component_sources(name)[:3]
    p = pl.stiff_params(name, N)
    if name == "chain16":
        N = 1 << 16
        p = pl.stiff_params(name, N)
    progs = {"one-thread": h.compile(pkg.ALG_ROSENBROCK23, pkg.F64, n, np_, r[0], r[1], j[0], j[1], tg[0], tg[1]),
             "lane-group": h.compile(pkg.ALG_ROSENBROCK23, pkg.F64, n, np_, rc[0], rc[1], jc[0], jc[1], tc[0] if tc else None,
                                     tc[1] if tc else None, extra_options=pkg._lib.OPT_COMPONENT_RHS),
             "lane-group, 2 CTAs/SM": h.compile(pkg.ALG_ROSENBROCK23, pkg.F64, n, np_, rc[0], rc[1], jc[0], jc[1], tc[0] if tc else None,
                                                tc[1] if tc else None, extra_options=pkg._lib.OPT_COMPONENT_RHS + " -DB200_MINBLOCKS=2")}
    ref = None
    for tag, prog in progs.items():
        best = min(ll.solve_host(prog, u0, p, tspan, reltol=1e-6, abstol=1e-8)["kernel_ms"] for _ in range(3))
        g = ll.solve_host(prog, u0, p, tspan, reltol=1e-6, abstol=1e-8)
        if ref is None:
            ref = g
        same = np.array_equal(g["u_final"].view(np.uint64), ref["u_final"].view(np.uint64)) and np.array_equal(g["naccept"], ref["naccept"])
        print("TIMING %s rosenbrock23 %s N=%d regs %d local %d B kernel_ms %.3f -> %.2f M traj/s  steps/traj %.1f  same_bits=%s"
              % (name, tag, N, prog.info["regs_integrate"], prog.info["local_bytes_integrate"], best, N / best / 1e3,
                 float((g["naccept"] + g["nreject"]).mean()), same), flush=True)
