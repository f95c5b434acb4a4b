"""Randomised differential test: option combinations drawn from a seeded generator, every case compared bit for bit
with the CPU oracle.  The hand-written parity tests cover each feature on its own; this one covers their products —
stepper x precision x (saveat | ragged rows | final states) x tstops x d_discontinuities x save_start / save_end x
tolerances x dt / dtmax / maxiters — on small ensembles (Lorenz for the explicit steppers, Robertson for the
Rosenbrock-type ones).  A failing case prints its seed and options."""
import numpy as np
import pytest

from helpers import assert_same_result, bits

pytestmark = pytest.mark.gpu

EXPLICIT = ["TSIT5", "VERN7", "DP5", "BS3", "VERN6", "VERN9"]
STIFF = ["ROSENBROCK23", "RODAS5P", "RODAS4", "RODAS3P", "ROSENBROCK32"]


def _draw(rng, stiff, f32):
    tf = float(rng.choice([1.0, 2.5, 10.0])) if not stiff else float(rng.choice([1.0, 40.0, 1e3]))
    kw = {}
    mode = rng.choice(["final", "saveat", "ragged", "ragged_saveat"])
    if mode in ("saveat", "ragged_saveat"):
        kind = rng.integers(0, 3)
        if kind == 0:
            m = int(rng.integers(1, 40))
            grid = [tf * (k + 1) / m for k in range(m)]
        elif kind == 1:
            grid = sorted(float(x) for x in rng.uniform(0.0, tf, size=int(rng.integers(1, 12))) if x > 0.0)
        else:
            grid = sorted(set([tf * 0.5, tf] + [float(x) for x in rng.uniform(0.0, tf, size=3) if x > 0.0]))
        if grid:
            kw["saveat"] = grid
    if rng.random() < 0.4:
        kw["save_start"] = bool(rng.integers(0, 2))
    if rng.random() < 0.4:
        kw["save_end"] = bool(rng.integers(0, 2))
    if rng.random() < 0.35:
        kw["tstops"] = sorted(float(x) for x in rng.uniform(-0.1 * tf, 1.1 * tf, size=int(rng.integers(1, 4))))
    if rng.random() < 0.3:
        kw["d_discontinuities"] = [float(x) for x in rng.choice([0.0, 0.25 * tf, 0.5 * tf, tf, 1.5 * tf], size=int(rng.integers(1, 3)), replace=False)]
    if rng.random() < 0.5:
        if f32:
            kw["reltol"], kw["abstol"] = float(rng.choice([1e-3, 1e-4])), float(rng.choice([1e-5, 1e-6]))
        else:
            kw["reltol"], kw["abstol"] = float(rng.choice([1e-3, 1e-6, 1e-9])), float(rng.choice([1e-6, 1e-8, 1e-11]))
    if rng.random() < 0.25:
        kw["dt"] = float(rng.choice([1e-3, 0.01, 0.2]))
    if rng.random() < 0.25:
        kw["dtmax"] = float(rng.choice([0.05, 0.3])) * (1.0 if not stiff else tf)
    if rng.random() < 0.2:
        kw["maxiters"] = int(rng.choice([5, 30, 200]))
    return mode, tf, kw


@pytest.mark.parametrize("seed", range(6))
def test_random_option_products(pkg, handle, seed):
    from oracle import oracle
    pl = pkg.problems_library
    rng = np.random.default_rng(1000 + seed)
    N = 96
    programs = {}
    try:
        for case in range(14):
            f32 = bool(rng.integers(0, 2))
            stiff = bool(rng.integers(0, 2))
            name = str(rng.choice(STIFF if stiff else EXPLICIT))
            mode, tf, kw = _draw(rng, stiff, f32)
            ragged = mode in ("ragged", "ragged_saveat")
            want_stops = ("tstops" in kw) or ("d_discontinuities" in kw)
            if ragged and name == "ROSENBROCK32":
                name = "ROSENBROCK23"
            rdt = np.float32 if f32 else np.float64
            if stiff:
                r, j, tg = pl.robertson_sources(f32)
                n, np_, u0, p = 3, 3, np.array([1.0, 0.0, 0.0], dtype=rdt), pl.robertson_params(N, f32=f32)
                okw = dict(jac=j, tgrad=tg)
                srcs = (r[0], r[1], j[0], j[1], tg[0], tg[1])
            else:
                r = pl.lorenz_source(f32)
                n, np_, u0, p = 3, 3, np.array([1.0, 0.0, 0.0], dtype=rdt), pl.lorenz_params(N, f32=f32)
                okw = {}
                srcs = (r[0], r[1])
            opts = []
            if want_stops:
                opts.append(pkg._lib.OPT_TSTOPS)
            if ragged:
                opts.append(pkg._lib.OPT_EVERYSTEP)
            key = (name, f32, tuple(opts))
            if key not in programs:
                programs[key] = handle.compile(getattr(pkg, "ALG_" + name), pkg.F32 if f32 else pkg.F64, n, np_, *srcs,
                                               extra_options=" ".join(opts) or None)
            prog = programs[key]
            tag = "seed %d case %d: %s f32=%s mode=%s tf=%g %r" % (seed, case, name, f32, mode, tf, kw)
            oalg = getattr(oracle, "ALG_" + name)
            try:
                if ragged:
                    g = pkg.lowlevel.solve_host_everystep(prog, u0, p, (0.0, tf), **kw)
                    o = oracle.solve(oalg, r, u0, p, (0.0, tf), n, np_, f32=f32, save_everystep=True, **dict(kw, **okw))
                    assert np.array_equal(g["row_offsets"], o["row_offsets"]), "row_offsets"
                    assert np.array_equal(g["ts"], o["ts"]), "ts"
                    assert np.array_equal(bits(g["us"]), bits(o["us"])), "ragged rows"
                    for k in ("naccept", "nreject", "nf", "retcode"):
                        assert np.array_equal(g[k], o[k]), k
                    assert np.array_equal(bits(g["u_final"]), bits(o["u_final"])), "u_final"
                else:
                    g = pkg.lowlevel.solve_host(prog, u0, p, (0.0, tf), **kw)
                    o = oracle.solve(oalg, r, u0, p, (0.0, tf), n, np_, f32=f32, **dict(kw, **okw))
                    assert_same_result(g, o)
            except AssertionError as e:
                raise AssertionError(tag + " -> " + str(e))
    finally:
        for pr in programs.values():
            pr.close()
