"""ncu driver, device-resident launch (one full 2^20-trajectory launch per solve): python scripts/prof_dev.py {f64|f32} [saveat]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200_import
pkg = b200_import.load()
pl, ll = pkg.problems_library, pkg.lowlevel
f32 = sys.argv[1] == "f32"
sv = pkg.ranges.saveat_grid(0.1, (0.0, 10.0)) if len(sys.argv) > 2 and sys.argv[2] == "saveat" else None
N = 1 << 20
h = pkg.Handle(0)
rhs = pl.lorenz_source(f32); p = pl.lorenz_params(N, f32=f32)
prog = h.compile(pkg.ALG_TSIT5, pkg.F32 if f32 else pkg.F64, 3, 3, rhs[0], rhs[1])
nslots = ll.nslots_for((0.0, 10.0), sv) if sv else 0
b = ll.DeviceBuffers(prog, N, nslots, "cuda:0", u0_shared=True)
b.u0.copy_(torch.tensor([1.0, 0, 0], dtype=b.u0.dtype)); b.p.copy_(torch.from_numpy(p))
for _ in range(3):
    ll.solve_device(prog, b, (0.0, 10.0), saveat=sv)
torch.cuda.synchronize()
print("ok", int(b.naccept.sum()))
