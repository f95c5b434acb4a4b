"""Oracle (and, with -m gpu, the CUDA path) against golden vectors produced by the REAL reference
(scripts/julia_golden.jl → tests/golden/julia/*.json).

No Julia runtime exists in the build container or on the GPU boxes (`which julia` is empty on both), so the
directory is empty in this repository and every case below reports XFAIL "parity unpinned" — deliberately not a
silent skip.  As soon as someone runs the generator with a Julia install and drops the JSON files in place, the same
cases become hard checks with north_star's bar: accepted/rejected step counts and nf identical per trajectory, final
and saveat states within 1e-10 relative (FP64) / 50·reltol (FP32)."""
import glob
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "julia")

CASES = {
    # file stem: (oracle alg attr, problem, f32, solve kwargs)
    "cfg1_lorenz_tsit5_reltol1e-8": ("ALG_TSIT5", "lorenz", False, dict(reltol=1e-8)),
    "cfg2_lorenz_tsit5_saveat_f64": ("ALG_TSIT5", "lorenz", False, dict(saveat=0.1)),
    "cfg2_lorenz_tsit5_saveat_f32": ("ALG_TSIT5", "lorenz", True, dict(saveat=0.1)),
    "lorenz_dp5_saveat_f64": ("ALG_DP5", "lorenz", False, dict(saveat=0.5)),
    "lorenz_bs3_saveat_f64": ("ALG_BS3", "lorenz", False, dict(saveat=0.5)),
    "lorenz_vern6_saveat_f64": ("ALG_VERN6", "lorenz", False, dict(saveat=0.5)),
    "lorenz_vern7_saveat_f64": ("ALG_VERN7", "lorenz", False, dict(saveat=0.5)),
    "lorenz_vern8_saveat_f64": ("ALG_VERN8", "lorenz", False, dict(saveat=0.5)),
    "lorenz_vern9_saveat_f64": ("ALG_VERN9", "lorenz", False, dict(saveat=0.5)),
    "cfg3_robertson_rodas5p": ("ALG_RODAS5P", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rosenbrock23": ("ALG_ROSENBROCK23", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas5": ("ALG_RODAS5", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas4": ("ALG_RODAS4", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas42": ("ALG_RODAS42", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas4p": ("ALG_RODAS4P", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas4p2": ("ALG_RODAS4P2", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas5pe": ("ALG_RODAS5PE", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas3p": ("ALG_RODAS3P", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas23w": ("ALG_RODAS23W", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_autotsit5_rosenbrock23": ("ALG_AUTOTSIT5_ROSENBROCK23", "robertson", False, dict(reltol=1e-6, abstol=1e-8)),
    "cfg3_robertson_rodas5p_saveat": ("ALG_RODAS5P", "robertson", False,
                                      dict(reltol=1e-6, abstol=1e-8, saveat=[100.0, 1000.0, 5.0e4])),
    "cfg4_pleiades_vern7": ("ALG_VERN7", "pleiades", False, dict(reltol=1e-6, abstol=1e-8)),
    "lorenz_tsit5_d_discontinuities": ("ALG_TSIT5", "lorenz", False, dict(d_discontinuities=[2.5, 5.0], tstops=[7.5], saveat=0.5)),
    "lorenz_vern7_d_discontinuities": ("ALG_VERN7", "lorenz", False, dict(d_discontinuities=[0.0, 2.5])),
    "lorenz_tsit5_tspans": ("ALG_TSIT5", "lorenz", False, dict(_tspans=True)),
    # reverse time (tspan[2] < tspan[1]); _tspan overrides the problem's span
    "lorenz_tsit5_reverse": ("ALG_TSIT5", "lorenz", False, dict(_tspan=(1.0, 0.0), saveat=0.1, tstops=[0.5])),
    "lorenz_vern7_reverse": ("ALG_VERN7", "lorenz", False, dict(_tspan=(1.0, 0.0), reltol=1e-8, abstol=1e-10)),
    "robertson_rodas5p_reverse": ("ALG_RODAS5P", "robertson", False, dict(_tspan=(1e-3, 0.0), reltol=1e-6, abstol=1e-8)),
}


def _load(stem):
    path = os.path.join(GOLD, stem + ".json")
    if not os.path.exists(path):
        pytest.xfail("parity unpinned: no Julia-generated golden vectors (%s); run scripts/julia_golden.jl with a Julia "
                     "install — none exists in this container or on the GPU boxes" % os.path.relpath(path, HERE))
    return json.load(open(path))


def _setup(pkg, problem, f32, N):
    pl = pkg.problems_library
    if problem == "lorenz":
        return pl.lorenz_source(f32), None, None, np.array([1.0, 0.0, 0.0]), pl.lorenz_params(N, f32=f32), (0.0, 10.0), 3, 3
    if problem == "robertson":
        r, j, tg = pl.robertson_sources(f32)
        return r, j, tg, np.array([1.0, 0.0, 0.0]), pl.robertson_params(N, f32=f32), (0.0, 1e5), 3, 3
    return pl.pleiades_source(f32), None, None, pl.pleiades_u0(N, f32=f32), None, (0.0, 3.0), 28, 0


def _spans(pkg, kw, tspan, N):
    """tspan argument of the case: the shared span, or (cases with _tspans) the per-trajectory spans the generator's
    prob_func builds: (0, 5 + 5 U(i, 1))."""
    if kw.get("_tspan"):
        return kw["_tspan"]
    if not kw.get("_tspans"):
        return tspan
    idx = np.arange(N, dtype=np.uint64)
    return np.stack([np.zeros(N), 5.0 + 5.0 * pkg.problems_library.splitmix64_uniform(idx, 1)], axis=1)


def _options(pkg, kw):
    """program options a case needs on the CUDA path"""
    opts = []
    if "tstops" in kw or "d_discontinuities" in kw:
        opts.append(pkg._lib.OPT_TSTOPS)
    if kw.get("_tspans"):
        opts.append(pkg._lib.OPT_TSPANS)
    if kw.get("_tspan") and kw["_tspan"][1] < kw["_tspan"][0]:
        opts.append(pkg._lib.OPT_REVERSE_TIME)
    return " ".join(opts) or None


def _grid(pkg, kw, tspan):
    kw = dict(kw)
    if "saveat" in kw and not isinstance(kw["saveat"], list):
        kw["saveat"] = pkg.ranges.saveat_grid(kw["saveat"], tspan)
    return kw


def _compare(res, gold, f32, reltol):
    N = gold["trajectories"]
    for k in ("naccept", "nreject", "nf"):
        assert np.array_equal(res[k], np.asarray(gold[k])), "%s differs from the reference" % k
    if max(gold["njacs"]) > 0:
        for k in ("njacs", "nw", "nsolve"):
            assert np.array_equal(res[k], np.asarray(gold[k])), "%s differs from the reference" % k
    assert all(rc == "Success" for rc in gold["retcode"]) and (res["retcode"] == 1).all()
    tol = 50 * reltol if f32 else 1e-10
    uf = np.asarray([u[-1] for u in gold["u"]], dtype=np.float64)
    scale = np.maximum(np.abs(uf), 1e-300)
    assert (np.abs(res["u_final"].astype(np.float64) - uf) / scale).max() <= tol
    if res.get("us") is not None:
        rows = np.asarray(gold["u"], dtype=np.float64)          # [N][nrows][n]
        assert rows.shape == res["us"].shape
        assert np.array_equal(np.asarray(gold["t"][0], dtype=np.float64), np.asarray(res["ts"], dtype=np.float64))
        assert (np.abs(res["us"].astype(np.float64) - rows) / np.maximum(np.abs(rows), 1e-300)).max() <= tol
    assert N == res["u_final"].shape[0]


def test_golden_directory_is_reported():
    """One line in every test report that says whether the oracle is pinned to the real reference."""
    files = glob.glob(os.path.join(GOLD, "*.json"))
    if not files:
        pytest.xfail("parity unpinned: tests/golden/julia/ holds no Julia-generated vectors")
    assert set(os.path.splitext(os.path.basename(f))[0] for f in files) <= set(CASES)


@pytest.mark.parametrize("stem", sorted(CASES))
def test_oracle_matches_julia_reference(pkg, stem):
    gold = _load(stem)
    from oracle import oracle
    algname, problem, f32, kw = CASES[stem]
    rhs, jac, tg, u0, p, tspan, n, np_ = _setup(pkg, problem, f32, gold["trajectories"])
    span = _spans(pkg, kw, tspan, gold["trajectories"])
    kw = {k: v for k, v in _grid(pkg, kw, kw.get("_tspan", tspan)).items() if not k.startswith("_")}
    o = oracle.solve(getattr(oracle, algname), rhs, u0, p, span, n, np_, f32=f32, jac=jac, tgrad=tg, **kw)
    _compare(o, gold, f32, kw.get("reltol", 1e-3))


@pytest.mark.gpu
@pytest.mark.parametrize("stem", sorted(CASES))
def test_cuda_path_matches_julia_reference(pkg, handle, stem):
    gold = _load(stem)
    algname, problem, f32, kw = CASES[stem]
    rhs, jac, tg, u0, p, tspan, n, np_ = _setup(pkg, problem, f32, gold["trajectories"])
    span = _spans(pkg, kw, tspan, gold["trajectories"])
    extra = _options(pkg, kw)
    kw = {k: v for k, v in _grid(pkg, kw, kw.get("_tspan", tspan)).items() if not k.startswith("_")}
    prog = handle.compile(getattr(pkg, algname), pkg.F32 if f32 else pkg.F64, n, np_, rhs[0], rhs[1],
                          jac[0] if jac else None, jac[1] if jac else None, tg[0] if tg else None, tg[1] if tg else None,
                          extra_options=extra)
    try:
        g = pkg.lowlevel.solve_host(prog, u0, p, span, **kw)
    finally:
        prog.close()
    _compare(g, gold, f32, kw.get("reltol", 1e-3))


def test_checker_accepts_matching_and_rejects_perturbed_vectors(pkg, tmp_path, monkeypatch):
    """The comparison itself is not vacuous: a file in the generator's format built from the oracle's own output passes,
    and the same file with one step count or one state perturbed beyond the tolerance fails."""
    import sys
    from oracle import oracle
    mod = sys.modules[__name__]
    stem = "cfg2_lorenz_tsit5_saveat_f64"
    algname, problem, f32, kw = CASES[stem]
    N = 8
    rhs, jac, tg, u0, p, tspan, n, np_ = _setup(pkg, problem, f32, N)
    kwg = _grid(pkg, kw, tspan)
    o = oracle.solve(getattr(oracle, algname), rhs, u0, p, tspan, n, np_, **kwg)
    gold = {"case": stem, "trajectories": N, "naccept": o["naccept"].tolist(), "nreject": o["nreject"].tolist(),
            "nf": o["nf"].tolist(), "njacs": [0] * N, "nw": [0] * N, "nsolve": [0] * N, "retcode": ["Success"] * N,
            "t": [list(map(float, o["ts"]))] * N, "u": o["us"].tolist()}
    monkeypatch.setattr(mod, "GOLD", str(tmp_path))
    json.dump(gold, open(os.path.join(str(tmp_path), stem + ".json"), "w"))
    _compare(o, _load(stem), f32, 1e-3)
    bad = dict(gold, naccept=[gold["naccept"][0] + 1] + gold["naccept"][1:])
    with pytest.raises(AssertionError):
        _compare(o, bad, f32, 1e-3)
    u = json.loads(json.dumps(gold["u"]))
    u[3][50][1] *= 1 + 1e-8
    with pytest.raises(AssertionError):
        _compare(o, dict(gold, u=u), f32, 1e-3)
