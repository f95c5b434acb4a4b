"""Writes tests/golden/oracle_pins.json: step counts and final-state bits of the CPU oracle on the
BASELINE configurations (small N).  These are regression pins of the restatement, produced by the
oracle itself — NOT vectors from a Julia run of the reference (none is possible here)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np  # noqa: E402
import b200_import  # noqa: E402
from test_oracle_properties import _run_pin_case  # noqa: E402

pkg = b200_import.load()
cases = [
    dict(name="cfg1 lorenz tsit5 reltol 1e-8", problem="lorenz", alg="tsit5", f32=False, N=16, kw=dict(reltol=1e-8)),
    dict(name="cfg2 lorenz tsit5 defaults", problem="lorenz", alg="tsit5", f32=False, N=16, kw=dict()),
    dict(name="cfg2 lorenz tsit5 f32", problem="lorenz", alg="tsit5", f32=True, N=16, kw=dict()),
    dict(name="lorenz vern7", problem="lorenz", alg="vern7", f32=False, N=8, kw=dict()),
    dict(name="cfg3 robertson rodas5p", problem="robertson", alg="rodas5p", f32=False, N=8, tf=1e5,
         kw=dict(reltol=1e-6, abstol=1e-8)),
    dict(name="cfg3 robertson rosenbrock23", problem="robertson", alg="ros23", f32=False, N=8, tf=1e5,
         kw=dict(reltol=1e-6, abstol=1e-8)),
    dict(name="cfg4 pleiades vern7", problem="pleiades", alg="vern7", f32=False, N=4, kw=dict(reltol=1e-6, abstol=1e-8)),
]
# steppers added for SURVEY §8(f) row 3
for a in ("dp5", "bs3", "vern6", "vern8", "vern9"):
    cases.append(dict(name="lorenz %s" % a, problem="lorenz", alg=a, f32=False, N=8, kw=dict()))
cases.append(dict(name="lorenz dp5 f32", problem="lorenz", alg="dp5", f32=True, N=8, kw=dict()))
cases.append(dict(name="robertson ros32 (mildly stiff span)", problem="robertson", alg="ros32", f32=False, N=4, tf=10.0,
                  kw=dict(reltol=1e-6, abstol=1e-8)))
for a in ("rodas5", "rodas4", "rodas42", "rodas4p", "rodas4p2"):
    cases.append(dict(name="robertson %s" % a, problem="robertson", alg=a, f32=False, N=4, tf=1e4,
                      kw=dict(reltol=1e-6, abstol=1e-8)))
# reverse time (tspan[2] < tspan[1])
cases.append(dict(name="lorenz tsit5 backwards", problem="lorenz", alg="tsit5", f32=False, N=8, tspan=[1.0, 0.0], kw=dict()))
cases.append(dict(name="lorenz vern7 backwards f32", problem="lorenz", alg="vern7", f32=True, N=8, tspan=[0.5, 0.0], kw=dict()))
cases.append(dict(name="robertson rodas5p backwards", problem="robertson", alg="rodas5p", f32=False, N=4, tspan=[1.0e-3, 0.0],
                  kw=dict(reltol=1e-6, abstol=1e-8)))
for c in cases:
    o = _run_pin_case(pkg.problems_library, c)
    c["naccept"] = [int(x) for x in o["naccept"]]
    c["nreject"] = [int(x) for x in o["nreject"]]
    c["u_final_hex"] = [float(x).hex() for x in o["u_final"].astype(np.float64).ravel()]
json.dump({"note": "oracle-generated regression pins; parity with Julia unpinned", "cases": cases},
          open(os.path.join(HERE, "oracle_pins.json"), "w"), indent=1)
print("wrote", len(cases), "cases")
