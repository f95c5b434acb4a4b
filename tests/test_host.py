"""Host-side logic that needs no GPU: Julia range arithmetic, C code generation, the keyword
contract of solve(), synthetic input tables."""
import numpy as np
import pytest


def test_julia_range_matches_twice_precision_values(pkg):
    r = pkg.ranges
    g = r.julia_range(0.1, 0.1, 10.0)
    assert len(g) == 100 and g[-1] == 10.0
    assert all(g[i] == (i + 1) / 10 for i in range(100))          # correctly rounded i/10
    assert g[2] != 3 * 0.1
    assert r.julia_range(0.0, 0.25, 1.0) == [0.0, 0.25, 0.5, 0.75, 1.0]
    assert r.julia_range(1 / 3, 1 / 3, 2.0) == [1 / 3, 2 / 3, 1.0, 4 / 3, 5 / 3, 2.0]
    assert r.julia_range(0.3, 0.3, 1.0) == [0.3, 0.6, 0.9]
    assert r.julia_range(1.0, 1.0, 0.5) == []
    assert r.saveat_grid([0.0, 0.5, 1.0, 2.0], (0.0, 1.0)) == [0.5, 1.0]      # t0 < t <= tf
    assert r.saveat_grid(None, (0.0, 1.0)) == []
    # reversed spans (tdir = -1, solve.jl:1107-1120): the range runs downwards in the same twice-precision arithmetic — each
    # element is the negation of the mirrored forward range's — and lists come back in the order the integrator meets them
    down = r.julia_range(0.9, -0.1, 0.0)
    assert down == [0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2, 0.1, 0.0] and down == [-x for x in r.julia_range(-0.9, 0.1, 0.0)]
    assert r.saveat_grid(0.1, (1.0, 0.0)) == down and r.saveat_grid(-0.1, (1.0, 0.0)) == down      # tdir * abs(saveat)
    assert r.saveat_grid([0.0, 0.5, 1.0, 2.0], (1.0, 0.0)) == [0.5, 0.0]                           # tdir t0 < tdir t <= tdir tf
    assert r.saveat_grid(0.3, (1.0, 0.0)) == [-x for x in r.julia_range(-0.7, 0.3, 0.0)] and len(r.saveat_grid(0.3, (1.0, 0.0))) == 3
    assert r.resolve_save_flags([0.0, 0.5, 1.0], (1.0, 0.0), False) == (True, None)
    assert r.resolve_save_flags([0.5], (1.0, 0.0), False) == (False, False)


def test_codegen_matches_handwritten_lorenz(pkg):
    from oracle import oracle
    cg = pkg.codegen
    src, name = cg.build_function_c(lambda u, p, t: [p[0] * (u[1] - u[0]), u[0] * (p[1] - u[2]) - u[1],
                                                     u[0] * u[1] - p[2] * u[2]], 3, 3)
    assert "void diffeqf(double* du, const double* RHS1, const double* RHS2, const double RHS3)" in src
    assert "pow(" not in src
    hand = pkg.problems_library.lorenz_source()
    p = pkg.problems_library.lorenz_params(64)
    a = oracle.solve(oracle.ALG_TSIT5, (src, name), np.array([1.0, 0, 0]), p, (0.0, 2.0), 3, 3)
    b = oracle.solve(oracle.ALG_TSIT5, hand, np.array([1.0, 0, 0]), p, (0.0, 2.0), 3, 3)
    assert np.array_equal(a["naccept"], b["naccept"])
    assert np.allclose(a["u_final"], b["u_final"], rtol=1e-9)


def test_codegen_jacobian_and_powers(pkg):
    cg = pkg.codegen
    f = lambda u, p, t: [-p[0] * u[0] + p[2] * u[1] * u[2], p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2],
                         p[1] * u[1] ** 2]
    jac, jn = cg.build_jacobian_c(f, 3, 3)
    tg, tn = cg.build_tgrad_c(f, 3, 3)
    assert "pow(" not in jac and jac.count("J[") == 9 and "dT[2] = 0" in tg
    src32, _ = cg.build_function_c(f, 3, 3, f32=True)
    assert "float* du" in src32 and "double" not in src32
    # the generated Jacobian agrees numerically with the handwritten one
    from oracle import oracle
    import ctypes as C
    (r, rn), (hj, hjn), _ = pkg.problems_library.robertson_sources()
    lib = oracle.compile_user([jac, hj])
    u = np.array([0.7, 2e-5, 0.3]); p = np.array([0.04, 3e7, 1e4]); J1 = np.zeros(9); J2 = np.zeros(9)
    for nm, J in ((jn, J1), (hjn, J2)):
        getattr(lib, nm)(C.c_void_p(J.ctypes.data), C.c_void_p(u.ctypes.data), C.c_void_p(p.ctypes.data), C.c_double(0.0))
    assert np.allclose(J1, J2, rtol=1e-14)


def test_solve_keyword_contract(pkg):
    P = pkg
    prob = P.ODEProblem(P.CSource(*P.problems_library.lorenz_source()), np.array([1.0, 0, 0]), (0.0, 1.0),
                        np.array([10.0, 28.0, 8 / 3]))
    ep = P.EnsembleProblem(prob)
    with pytest.raises(NotImplementedError):
        P.EnsembleThreads()                                         # no CPU path in this package
    with pytest.raises(NotImplementedError):
        P.solve(ep, P.Tsit5(), None, trajectories=4, saveat=0.1)
    with pytest.raises(TypeError):                                  # unknown keyword => error, like the reference
        P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=4, saveat=0.1, internalnorm=None)
    with pytest.raises(TypeError):
        P.solve(ep, P.Tsit5(), P.EnsembleB200(), saveat=0.1)        # trajectories missing
    with pytest.raises(TypeError):                                  # logging switches are accepted (and ignored) ...
        P.solve(ep, P.Tsit5(), P.EnsembleB200(), saveat=0.1, verbose=False, progress=True)    # ... trajectories still missing
    with pytest.raises(NotImplementedError):                        # dense sol(t) objects are not on this path
        P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=4, saveat=0.1, dense=True)
    with pytest.raises(ValueError):                                 # solve.jl:277-280: fixed step needs dt (or tstops)
        P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=4, saveat=0.1, adaptive=False)


def test_synthetic_tables_are_deterministic(pkg):
    pl = pkg.problems_library
    p = pl.lorenz_params(1000)
    assert p.shape == (1000, 3) and (p[:, 1] >= 14).all() and (p[:, 1] < 42).all()
    assert np.array_equal(p[10:20], pl.lorenz_params(10, offset=10))
    assert np.array_equal(pl.lorenz_params(5, f32=True), pl.lorenz_params(5).astype(np.float32))
    # SplitMix64 known answer: first output of seed 0x9E3779B97F4A7C15 stream position (i=0,j=0)
    u = pl.splitmix64_uniform(np.array([0]), 0)[0]
    x = (0x9E3779B97F4A7C15 + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
    z = z ^ (z >> 31)
    assert u == (z >> 11) / 2 ** 53
    k = pl.robertson_params(100)
    assert (k[:, 0] >= 0.02).all() and (k[:, 0] < 0.06).all() and (k[:, 1] >= 1.5e7).all()
    u0 = pl.pleiades_u0(50)
    assert u0.shape == (50, 28) and np.abs(u0[:, :14] - pl.PLEIADES_U0[:14]).max() <= 0.01
    assert np.array_equal(u0[:, 14:], np.tile(pl.PLEIADES_U0[14:], (50, 1)))


def test_save_start_and_save_end_defaults(pkg):
    # solve.jl:141-143,596-599; exercised by test/InterfaceI/ode_saveat_tests.jl:11-32
    # (saveat = [1/2] alone gives sol.t == [1/2]; saveat = 1/2 as a Number keeps both end points)
    f = pkg.ranges.resolve_save_flags
    span = (0.0, 1.0)
    assert f(None, span, True) == (True, None)                 # save_everystep: both ends
    assert f(None, span, False) == (True, None)                # isempty(saveat)
    assert f(0.5, span, False) == (True, None)                 # saveat isa Number
    assert f([0.5], span, False) == (False, False)             # neither end point is in the vector
    assert f([0.0, 0.5, 1.0], span, False) == (True, None)
    assert f([0.5, 1.0], span, False) == (False, None)
    assert f([0.5], span, True) == (True, None)
    assert f([0.5], span, False, save_start=True, save_end=True) == (True, True)
    assert f(0.5, span, False, save_end=False) == (True, False)
