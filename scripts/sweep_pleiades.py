"""Pleiades/Vern7 (BASELINE config 4), 2^18 trajectories: the lane-group kernel at several launch shapes against the
one-thread-per-trajectory kernel; parity of the first 512 trajectories against the oracle.
python scripts/sweep_pleiades.py [variants...]   (a variant starting with 'T:' uses the one-thread kernel, 'W:' the
shared-memory stage kernel with the pair-shared source, 'WP:' the same kernel with the reference's plain double loop)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200_import
pkg = b200_import.load()
from oracle import oracle
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
N = 1 << 18
full = pl.pleiades_source(False); comp = pl.pleiades_component_source(False); u0 = pl.pleiades_u0(N)
kw = dict(reltol=1e-6, abstol=1e-8)
o = oracle.solve(oracle.ALG_VERN7, full, u0[:512], None, (0.0, 3.0), 28, 0, **kw)
variants = sys.argv[1:] or ["-DB200_COOP=1", "-DB200_COOP=1 -DB200_MINBLOCKS=3", "-DB200_COOP=1 -DB200_MINBLOCKS=2",
                            "-DB200_COOP=1 -DB200_BLOCK=256 -DB200_MINBLOCKS=2", "-DB200_COOP=1 -DB200_BLOCK=64 -DB200_MINBLOCKS=8",
                            "-DB200_COOP=1 -DB200_L=32 -DB200_MINBLOCKS=4", "-DB200_COOP=1 -DB200_L=8 -DB200_MINBLOCKS=2"]
for v in variants:
    thread = v.startswith("T:")
    opt = v[2:] if thread else v
    src = pl.pleiades_source(False, loops=True) if thread else comp
    if v.startswith("W:"):
        opt = ("-DB200_WIDE=1 " + v[2:]).strip(); src = pl.pleiades_pairs_source(False)
    if v.startswith("WP:"):
        opt = ("-DB200_WIDE=1 " + v[3:]).strip(); src = pl.pleiades_source(False)
    prog = h.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, src[0], src[1], extra_options=opt or None)
    best = 1e9
    for _ in range(3):
        g = ll.solve_host(prog, u0, None, (0.0, 3.0), **kw)
        best = min(best, g["kernel_ms"])
    ok = np.array_equal(g["naccept"][:512], o["naccept"]) and np.array_equal(g["u_final"][:512].view(np.uint64), o["u_final"].view(np.uint64))
    print(repr(v), "regs", prog.info["regs_integrate"], "local", prog.info["local_bytes_integrate"], "blocks/SM", prog.info["blocks_per_sm"],
          "block", prog.info["block"], "kernel_ms %.2f" % best, "-> %.2f M traj/s" % (N / best / 1e3), "parity", ok, "compile_s %.1f" % (prog.info["compile_ms"] / 1e3), flush=True)
    prog.close()
