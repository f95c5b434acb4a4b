"""Device-resident timing sweep (torch events around b200ode_solve_device): python scripts/sweep_dev.py {f32|f64} opts..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
h = pkg.Handle(0)
N = 1 << 20
f32 = sys.argv[1] == "f32"
variants = sys.argv[2:]
rhs = pl.lorenz_source(f32); p = pl.lorenz_params(N, f32=f32)
grid = pkg.ranges.saveat_grid(0.1, (0.0, 10.0))
for v in variants:
    prog = h.compile(pkg.ALG_TSIT5, pkg.F32 if f32 else pkg.F64, 3, 3, rhs[0], rhs[1], extra_options=v or None)
    res = []
    for sv in (None, grid):
        nslots = ll.nslots_for((0.0, 10.0), sv) if sv else 0
        b = ll.DeviceBuffers(prog, N, nslots, "cuda:0", u0_shared=True)
        b.u0.copy_(torch.tensor([1.0, 0, 0], dtype=b.u0.dtype)); b.p.copy_(torch.from_numpy(p))
        for _ in range(3):
            ll.solve_device(prog, b, (0.0, 10.0), saveat=sv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ll.solve_device(prog, b, (0.0, 10.0), saveat=sv)
        e1.record(); torch.cuda.synchronize()
        res.append(round(e0.elapsed_time(e1) / 10, 3))
        del b
    print("f32" if f32 else "f64", repr(v), "regs", prog.info["regs_integrate"], "local", prog.info["local_bytes_integrate"],
          "blocks/SM", prog.info["blocks_per_sm"], "ms (final, saveat)", res, flush=True)
    prog.close()
