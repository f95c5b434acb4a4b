"""Where the shared-memory stage kernel starts to pay: Vern7 on systems of n = 8, 12, 16, 20, 28 (a nonlinear chain
u_i' = -c u_i + d (u_{i-1} - 2 u_i + u_{i+1}) + sin-free cubic term; Pleiades for 28), plain kernel against
B200ODE_OPT_SMEM_STAGES, 2^17 trajectories; both must agree bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
N = 1 << 17


def chain_source(n):
    L = ["void chain_rhs(double* du, const double* u, const double* p, const double t) {"]
    for i in range(n):
        lo = "u[%d]" % (i - 1) if i > 0 else "0.0"
        hi = "u[%d]" % (i + 1) if i < n - 1 else "0.0"
        L.append("  du[%d] = p[0] * (%s - 2.0 * u[%d] + %s) - p[1] * u[%d] * u[%d] * u[%d] + %s;" % (i, lo, i, hi, i, i, i, "1.0" if i == 0 else "0.0"))
    L.append("}")
    return "\n".join(L) + "\n", "chain_rhs"


idx = np.arange(N, dtype=np.uint64)
p = np.stack([1.0 + 4.0 * pl.splitmix64_uniform(idx, 0), 0.5 + pl.splitmix64_uniform(idx, 1)], axis=1)
for n in (8, 12, 16, 20, 24):
    src, name = chain_source(n)
    u0 = np.zeros(n)
    res = {}
    for tag, opt in (("plain", None), ("smem", pkg._lib.OPT_SMEM_STAGES)):
        prog = h.compile(pkg.ALG_VERN7, pkg.F64, n, 2, src, name, extra_options=opt)
        best = 1e9
        for _ in range(3):
            g = ll.solve_host(prog, u0, p, (0.0, 10.0), reltol=1e-6, abstol=1e-8)
            best = min(best, g["kernel_ms"])
        res[tag] = (best, g, prog.info["regs_integrate"], prog.info["local_bytes_integrate"], prog.info["block"])
        prog.close()
    same = np.array_equal(res["plain"][1]["u_final"].view(np.uint64), res["smem"][1]["u_final"].view(np.uint64)) and \
        np.array_equal(res["plain"][1]["naccept"], res["smem"][1]["naccept"])
    print("n=%d plain %.2f ms (regs %d local %d) | smem stages %.2f ms (regs %d, block %d) | identical %s | steps/traj %.1f" %
          (n, res["plain"][0], res["plain"][2], res["plain"][3], res["smem"][0], res["smem"][2], res["smem"][4], same,
           float((g["naccept"] + g["nreject"]).mean())), flush=True)
