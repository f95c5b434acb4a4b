"""First GPU contact: parity vs oracle on Lorenz/Tsit5 + a timing at 1M trajectories."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
from oracle import oracle
pl, ll = pkg.problems_library, pkg.lowlevel

h = pkg.Handle(0)
print("fma peak f64", h.measure_fma_peak(pkg.F64), "f32", h.measure_fma_peak(pkg.F32))
rhs = pl.lorenz_source()
prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
print(prog.info)
u0 = np.array([1.0, 0.0, 0.0])
for (N, kw) in [(10000, dict(reltol=1e-8)), (10000, dict(saveat=np.arange(1, 101) / 10.0)), (10000, dict())]:
    p = pl.lorenz_params(N)
    for flags in (0, 1):
        g = ll.solve_host(prog, u0, p, (0.0, 10.0), flags=flags, **kw)
        o = oracle.solve(oracle.ALG_TSIT5, rhs, u0, p, (0.0, 10.0), 3, 3, **kw)
        ok = {k: bool(np.array_equal(g[k], o[k])) for k in ("naccept", "nreject", "nf", "retcode", "nsaved")}
        ok["u_final_bits"] = bool(np.array_equal(g["u_final"].view(np.uint64), o["u_final"].view(np.uint64)))
        ok["t_final"] = bool(np.array_equal(g["t_final"], o["t_final"]))
        if o["us"] is not None:
            ok["us_bits"] = bool(np.array_equal(g["us"].view(np.uint64), o["us"].view(np.uint64)))
            ok["ts"] = bool(np.array_equal(g["ts"], o["ts"]))
        print(N, list(kw), "flags", flags, ok, "kernel_ms", g["kernel_ms"], "total_ms", g["total_ms"])
        if not all(ok.values()):
            bad = np.nonzero(g["naccept"] != o["naccept"])[0]
            print("  mismatching naccept idx", bad[:10], g["naccept"][bad[:5]], o["naccept"][bad[:5]])
            d = np.abs(g["u_final"] - o["u_final"]).max()
            print("  max |du_final|", d)

for N in (1 << 20,):
    p = pl.lorenz_params(N)
    for kw in (dict(), dict(saveat=np.arange(1, 101) / 10.0)):
        for flags in (0, 1):
            for rep in range(3):
                g = ll.solve_host(prog, u0, p, (0.0, 10.0), flags=flags, **kw)
            steps = int(g["naccept"].sum() + g["nreject"].sum())
            print("N", N, list(kw), "flags", flags, "kernel_ms", round(g["kernel_ms"], 3), "total_ms",
                  round(g["total_ms"], 3), "traj/s(kernel)", N / g["kernel_ms"] * 1e3, "steps", steps,
                  "GFLOP/s", 247.0 * steps / g["kernel_ms"] / 1e6)
