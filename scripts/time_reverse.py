"""Per-step cost of a reverse-time program (B200ODE_OPT_REVERSE_TIME: mirrored problem, wrapped RHS) against the ordinary
forward program: Lorenz / Tsit5, 1 Mi trajectories over one time unit in either direction, final states and saveat = 0.1;
Robertson / Rodas5P over (1e-3, 0) and (0, 1e-3)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll, L = pkg.problems_library, pkg.lowlevel, pkg._lib
h = pkg.Handle(0)
N = 1 << 20
out = []


def run(tag, prog, u0, p, span, **kw):
    best, r = None, None
    for _ in range(4):
        r = ll.solve_host(prog, u0, p, span, **kw)
        best = r["kernel_ms"] if best is None else min(best, r["kernel_ms"])
    steps = int(r["naccept"].astype(np.int64).sum() + r["nreject"].astype(np.int64).sum())
    rec = dict(case=tag, span=list(span), kernel_ms=round(best, 4), attempted_steps=steps, ns_per_step=round(best * 1e6 / steps, 4),
               regs=prog.info["regs_integrate"], success=float((r["retcode"] == 1).mean()))
    print("TIMING", json.dumps(rec), flush=True)
    out.append(rec)


u0 = np.array([1.0, 0, 0])
rhs = pl.lorenz_source(False); p = pl.lorenz_params(N)
fwd = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
rev = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1], extra_options=L.OPT_REVERSE_TIME)
run("lorenz_tsit5_forward_final", fwd, u0, p, (0.0, 1.0))
run("lorenz_tsit5_reverse_final", rev, u0, p, (1.0, 0.0))
run("lorenz_tsit5_forward_saveat", fwd, u0, p, (0.0, 1.0), saveat=pkg.ranges.saveat_grid(0.01, (0.0, 1.0)))
run("lorenz_tsit5_reverse_saveat", rev, u0, p, (1.0, 0.0), saveat=pkg.ranges.saveat_grid(0.01, (1.0, 0.0)))
(r, rn), (j, jn), (tg, tgn) = pl.robertson_sources(False)
pr = pl.robertson_params(N)
fwd = h.compile(pkg.ALG_RODAS5P, pkg.F64, 3, 3, r, rn, j, jn, tg, tgn)
rev = h.compile(pkg.ALG_RODAS5P, pkg.F64, 3, 3, r, rn, j, jn, tg, tgn, extra_options=L.OPT_REVERSE_TIME)
run("robertson_rodas5p_forward_final", fwd, u0, pr, (0.0, 1e-3), reltol=1e-6, abstol=1e-8)
run("robertson_rodas5p_reverse_final", rev, u0, pr, (1e-3, 0.0), reltol=1e-6, abstol=1e-8)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2d_time_reverse.json", "w"), indent=1)
