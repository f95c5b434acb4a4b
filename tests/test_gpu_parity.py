"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs — bit-exact step counts, stats, final states and saveat rows, FP64 and FP32 — plus
size-independent properties at BASELINE.json's full sizes.

Tolerances: none.  Oracle and kernel perform the same IEEE operations in the same order (explicit
fma where the reference's @muladd fuses, no contraction elsewhere), so every comparison below is
equality of bits; this is stronger than north_star's 1e-10 (FP64) / reltol-scaled (FP32) bounds."""
import numpy as np
import pytest

from helpers import assert_same_result, bits, linear_source

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


@pytest.fixture(scope="module")
def progs(pkg, handle):
    cache = {}

    def get(alg, f32, problem):
        key = (alg, f32, problem)
        if key not in cache:
            pl = pkg.problems_library
            dt = pkg.F32 if f32 else pkg.F64
            if problem == "lorenz":
                s, n = pl.lorenz_source(f32)
                cache[key] = handle.compile(alg, dt, 3, 3, s, n)
            elif problem == "robertson":
                (r, j, tg) = pl.robertson_sources(f32)
                cache[key] = handle.compile(alg, dt, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1])
            elif problem == "pleiades":
                s, n = pl.pleiades_source(f32)
                cache[key] = handle.compile(alg, dt, 28, 0, s, n)
            elif problem == "linear":
                s, n = linear_source(f32)
                cache[key] = handle.compile(alg, dt, 1, 0, s, n)
        return cache[key]
    return get


U0 = np.array([1.0, 0.0, 0.0])
GRID = [k / 10 for k in range(1, 101)]


# ---- configs[0]: Lorenz Tsit5 10k trajectories, reltol 1e-8 -------------------------------------
def test_config1_lorenz_tsit5_10k(pkg, progs, oracle):
    N = 10000
    p = pkg.problems_library.lorenz_params(N)
    g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, False, "lorenz"), U0, p, (0.0, 10.0), reltol=1e-8)
    o = oracle.solve(oracle.ALG_TSIT5, pkg.problems_library.lorenz_source(), U0, p, (0.0, 10.0), 3, 3, reltol=1e-8)
    assert_same_result(g, o)
    assert (g["retcode"] == 1).all() and (g["t_final"] == 10.0).all()


@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("flags", [0, 1])
def test_config2_lorenz_tsit5_saveat(pkg, progs, oracle, f32, flags):
    N = 4096
    pl = pkg.problems_library
    p = pl.lorenz_params(N, f32=f32)
    g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, f32, "lorenz"), U0, p, (0.0, 10.0), saveat=GRID, flags=flags)
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(f32), U0, p, (0.0, 10.0), 3, 3, f32=f32, saveat=GRID)
    assert g["nslots"] == 101
    assert_same_result(g, o)


@pytest.mark.parametrize("alg_name", ["ros23", "rodas5p"])
@pytest.mark.parametrize("f32", [False, True])
def test_config3_robertson(pkg, progs, oracle, alg_name, f32):
    N = 2048
    pl = pkg.problems_library
    alg, oalg = {"ros23": (pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23),
                 "rodas5p": (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P)}[alg_name]
    r, j, tg = pl.robertson_sources(f32)
    p = pl.robertson_params(N, f32=f32)
    tf = 1e5 if not f32 else 1e3
    tol = dict(reltol=1e-6, abstol=1e-8) if not f32 else dict(reltol=1e-3, abstol=1e-5)
    for extra in ({}, {"saveat": [tf * 1e-3, tf * 1e-2, tf * 0.5]}):
        kw = dict(tol, **extra)
        g = pkg.lowlevel.solve_host(progs(alg, f32, "robertson"), U0, p, (0.0, tf), **kw)
        o = oracle.solve(oalg, r, U0, p, (0.0, tf), 3, 3, f32=f32, jac=j, tgrad=tg, **kw)
        assert_same_result(g, o)
        assert (g["retcode"] == 1).all()
        assert (g["njacs"] == 2 * (g["naccept"] + g["nreject"])).all()
        assert np.abs(g["u_final"].astype(np.float64).sum(axis=1) - 1.0).max() < (1e-9 if not f32 else 1e-3)


def test_config4_pleiades_vern7(pkg, progs, oracle):
    N = 512
    pl = pkg.problems_library
    u0 = pl.pleiades_u0(N)
    kw = dict(reltol=1e-6, abstol=1e-8)
    for extra in ({}, {"saveat": [0.5, 1.0, 1.5, 2.0, 2.5, 3.0]}):
        g = pkg.lowlevel.solve_host(progs(pkg.ALG_VERN7, False, "pleiades"), u0, None, (0.0, 3.0), **dict(kw, **extra))
        o = oracle.solve(oracle.ALG_VERN7, pl.pleiades_source(), u0, None, (0.0, 3.0), 28, 0, **dict(kw, **extra))
        assert_same_result(g, o)
        assert (g["nf"] == 2 + 10 * (g["naccept"] + g["nreject"])).all()


@pytest.mark.parametrize("f32", [False, True])
def test_vern7_lorenz_with_lazy_interpolation(pkg, progs, oracle, f32):
    N = 2048
    pl = pkg.problems_library
    p = pl.lorenz_params(N, f32=f32)
    grid = [k / 2 for k in range(1, 21)]
    g = pkg.lowlevel.solve_host(progs(pkg.ALG_VERN7, f32, "lorenz"), U0, p, (0.0, 10.0), saveat=grid)
    o = oracle.solve(oracle.ALG_VERN7, pl.lorenz_source(f32), U0, p, (0.0, 10.0), 3, 3, f32=f32, saveat=grid)
    assert_same_result(g, o)


# ---- edge cases the reference tests exercise -----------------------------------------------------
@pytest.mark.parametrize("N", [1, 31, 33, 1000])
def test_ragged_trajectory_counts(pkg, progs, oracle, N):
    pl = pkg.problems_library
    p = pl.lorenz_params(N)
    for flags in (0, 1):
        g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, False, "lorenz"), U0, p, (0.0, 3.0), saveat=[1.0, 2.5], flags=flags)
        o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, p, (0.0, 3.0), 3, 3, saveat=[1.0, 2.5])
        assert_same_result(g, o)


def test_zero_trajectories(pkg, progs):
    g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, False, "lorenz"), np.zeros((0, 3)), np.zeros((0, 3)), (0.0, 1.0))
    assert g["u_final"].shape == (0, 3)


@pytest.mark.parametrize("save_start,save_end,grid,expect_ts", [
    (None, None, [4.0, 8.0, 12.0], [0.0, 4.0, 8.0, 12.0, 15.0]),      # test/InterfaceI/ode_saveat_tests.jl
    (False, None, [4.0, 8.0, 12.0], [4.0, 8.0, 12.0, 15.0]),
    (None, False, [4.0, 8.0, 12.0], [0.0, 4.0, 8.0, 12.0]),
    (False, False, [4.0, 8.0, 12.0], [4.0, 8.0, 12.0]),
    (None, None, [5.0, 15.0], [0.0, 5.0, 15.0]),
    (None, False, [5.0, 15.0], [0.0, 5.0]),                            # skip_saveat_at_tspan_end
    (None, None, [5.0, 5.0, 7.0], [0.0, 5.0, 5.0, 7.0, 15.0]),         # duplicates are saved twice (heap order)
])
def test_saveat_bookkeeping(pkg, progs, oracle, save_start, save_end, grid, expect_ts):
    pl = pkg.problems_library
    N = 64
    p = pl.lorenz_params(N)
    kw = dict(saveat=grid, save_start=save_start, save_end=save_end)
    g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, False, "lorenz"), U0, p, (0.0, 15.0), **kw)
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, p, (0.0, 15.0), 3, 3, **kw)
    assert list(g["ts"]) == expect_ts
    assert_same_result(g, o)


def test_shared_and_per_trajectory_inputs(pkg, progs, oracle):
    pl = pkg.problems_library
    N = 257
    rng = np.random.default_rng(5)
    u0 = np.array([1.0, 0, 0]) + 0.1 * rng.standard_normal((N, 3))
    pshared = np.array([10.0, 28.0, 8 / 3])
    g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, False, "lorenz"), u0, pshared, (0.0, 2.0))
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), u0, pshared, (0.0, 2.0), 3, 3)
    assert_same_result(g, o)
    g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, False, "lorenz"), U0, pshared, (0.0, 2.0), trajectories=5)
    assert (bits(g["u_final"]) == bits(g["u_final"][0])).all()


def test_user_dt_dtmax_and_negative_t0(pkg, progs, oracle):
    pl = pkg.problems_library
    N = 128
    p = pl.lorenz_params(N)
    for kw, tspan in ((dict(dt=0.01), (0.0, 2.0)), (dict(dtmax=0.05), (0.0, 2.0)), (dict(), (-5.0, -3.0)),
                      (dict(saveat=[-4.5, -3.25]), (-5.0, -3.0)), (dict(dtmin=1e-4), (0.0, 2.0))):
        g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, False, "lorenz"), U0, p, tspan, **kw)
        o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, p, tspan, 3, 3, **kw)
        assert_same_result(g, o)


def test_failure_retcodes_do_not_stall_the_warp(pkg, handle, progs, oracle):
    """A failing trajectory gets a retcode and leaves its lane; its warp mates finish normally
    (check_error.jl:77-117).  maxiters: MaxIters.  RHS with a pole: Unstable/DtLessThanMin/DtNaN."""
    pl = pkg.problems_library
    N = 200
    p = pl.lorenz_params(N)
    g = pkg.lowlevel.solve_host(progs(pkg.ALG_TSIT5, False, "lorenz"), U0, p, (0.0, 10.0), maxiters=100, saveat=[5.0])
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, p, (0.0, 10.0), 3, 3, maxiters=100, saveat=[5.0])
    assert_same_result(g, o)
    assert set(np.unique(g["retcode"])) == {1, 2}
    assert (g["t_final"][g["retcode"] == 2] < 10.0).all()
    # blow-up in finite time: u' = u^2, u0 = p => pole at t = 1/p; some trajectories reach tf first
    src = ("void blow(double* du, const double* u, const double* p, const double t) { du[0] = u[0] * u[0]; }\n", "blow")
    prog = handle.compile(pkg.ALG_TSIT5, pkg.F64, 1, 1, src[0], src[1])
    u0 = np.linspace(0.5, 2.0, 64).reshape(64, 1)
    pp = np.zeros((64, 1))
    g = pkg.lowlevel.solve_host(prog, u0, pp, (0.0, 1.0))
    o = oracle.solve(oracle.ALG_TSIT5, src, u0, pp, (0.0, 1.0), 1, 1)
    assert_same_result(g, o)
    assert (g["retcode"][u0[:, 0] < 0.9] == 1).all() and (g["retcode"][u0[:, 0] > 1.1] != 1).all()


def test_device_api_soa_layout(pkg, progs, oracle):
    import torch
    pl, ll = pkg.problems_library, pkg.lowlevel
    N = 3000
    prog = progs(pkg.ALG_TSIT5, False, "lorenz")
    p = pl.lorenz_params(N)
    nslots = ll.nslots_for((0.0, 10.0), GRID)
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, p, (0.0, 10.0), 3, 3, saveat=GRID)
    for layout in (pkg._lib.LAYOUT_AOS, pkg._lib.LAYOUT_SOA):
        b = ll.DeviceBuffers(prog, N, nslots, "cuda:0", u0_shared=True, layout=layout)
        b.u0.copy_(torch.from_numpy(U0))
        pt = torch.from_numpy(p)
        b.p.copy_(pt if layout == pkg._lib.LAYOUT_AOS else pt.t().contiguous())
        ll.solve_device(prog, b, (0.0, 10.0), saveat=GRID)
        torch.cuda.synchronize()
        uf = b.u_final.cpu().numpy()
        uf = uf if layout == pkg._lib.LAYOUT_AOS else uf.T
        assert np.array_equal(bits(np.ascontiguousarray(uf)), bits(o["u_final"]))
        assert np.array_equal(bits(b.us.cpu().numpy()), bits(o["us"]))
        assert np.array_equal(b.naccept.cpu().numpy(), o["naccept"])
        # deterministic on-device reduction == host sum in the same order class
        out = torch.zeros(3, dtype=torch.float64, device="cuda:0")
        ll.reduce_sum_device(prog.handle, pkg.F64, b.u_final, layout, N, 3, out)
        torch.cuda.synchronize()
        assert np.allclose(out.cpu().numpy(), o["u_final"].sum(axis=0), rtol=1e-12)


# ---- full-size properties (BASELINE sizes; the oracle checks a seeded subsample) ------------------
def test_full_size_1M_lorenz_properties(pkg, progs, oracle):
    pl, ll = pkg.problems_library, pkg.lowlevel
    N = 1 << 20
    p = pl.lorenz_params(N)
    prog = progs(pkg.ALG_TSIT5, False, "lorenz")
    g = ll.solve_host(prog, U0, p, (0.0, 10.0), saveat=GRID)
    assert (g["retcode"] == 1).all() and (g["nsaved"] == 101).all() and (g["t_final"] == 10.0).all()
    assert (g["nf"] == 3 + 6 * (g["naccept"] + g["nreject"])).all()
    assert np.array_equal(bits(g["us"][:, -1, :]), bits(g["u_final"]))          # last row is u(tf)
    assert (g["us"][:, 0, :] == U0).all()
    # schedule independence: lane refill vs one-thread-per-trajectory give identical bits
    s = ll.solve_host(prog, U0, p, (0.0, 10.0), saveat=GRID, flags=pkg._lib.FLAG_STATIC_SCHEDULE)
    for k in ("naccept", "nreject", "nf", "retcode", "nsaved"):
        assert np.array_equal(g[k], s[k])
    assert np.array_equal(bits(g["us"]), bits(s["us"]))
    # permutation equivariance: shuffled inputs give the shuffled outputs
    perm = np.random.default_rng(7).permutation(N)
    q = ll.solve_host(prog, U0, p[perm], (0.0, 10.0))
    assert np.array_equal(bits(q["u_final"]), bits(g["u_final"][perm]))
    assert np.array_equal(q["naccept"], g["naccept"][perm])
    # oracle on a seeded subsample of the full run
    idx = np.sort(np.random.default_rng(8).choice(N, 3000, replace=False))
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, p[idx], (0.0, 10.0), 3, 3, saveat=GRID)
    assert np.array_equal(g["naccept"][idx], o["naccept"]) and np.array_equal(g["nreject"][idx], o["nreject"])
    assert np.array_equal(bits(g["us"][idx]), bits(o["us"]))


def test_full_size_robertson_and_pleiades_properties(pkg, progs, oracle):
    """configs[2] and configs[3] at BASELINE sizes: every trajectory succeeds, the invariants of the
    problems hold, and the oracle agrees bit for bit on a seeded subsample."""
    pl, ll = pkg.problems_library, pkg.lowlevel
    N = 1 << 20
    k = pl.robertson_params(N)
    r, j, tg = pl.robertson_sources(False)
    idx = np.sort(np.random.default_rng(11).choice(N, 1500, replace=False))
    for alg, oalg in ((pkg.ALG_RODAS5P, oracle.ALG_RODAS5P), (pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23)):
        g = ll.solve_host(progs(alg, False, "robertson"), U0, k, (0.0, 1e5), reltol=1e-6, abstol=1e-8)
        assert (g["retcode"] == 1).all() and (g["t_final"] == 1e5).all()
        assert np.abs(g["u_final"].sum(axis=1) - 1.0).max() < 1e-9            # mass conservation
        assert (g["u_final"] > -1e-9).all()
        o = oracle.solve(oalg, r, U0, k[idx], (0.0, 1e5), 3, 3, jac=j, tgrad=tg, reltol=1e-6, abstol=1e-8)
        assert np.array_equal(g["naccept"][idx], o["naccept"]) and np.array_equal(g["nreject"][idx], o["nreject"])
        assert np.array_equal(bits(g["u_final"][idx]), bits(o["u_final"]))
    Np = 1 << 18
    u0 = pl.pleiades_u0(Np)
    g = ll.solve_host(progs(pkg.ALG_VERN7, False, "pleiades"), u0, None, (0.0, 3.0), reltol=1e-6, abstol=1e-8)
    assert (g["retcode"] == 1).all()
    assert (g["nf"] == 2 + 10 * (g["naccept"] + g["nreject"])).all()
    # total linear momentum sum_j m_j v_j is conserved by the pairwise forces (m_j = j)
    m = np.arange(1, 8, dtype=np.float64)
    for lo in (14, 21):
        p0 = (u0[:, lo:lo + 7] * m).sum(axis=1)
        p1 = (g["u_final"][:, lo:lo + 7] * m).sum(axis=1)
        assert np.abs(p1 - p0).max() < 1e-4
    idx = np.sort(np.random.default_rng(12).choice(Np, 96, replace=False))
    o = oracle.solve(oracle.ALG_VERN7, pl.pleiades_source(), u0[idx], None, (0.0, 3.0), 28, 0, reltol=1e-6, abstol=1e-8)
    assert np.array_equal(g["naccept"][idx], o["naccept"]) and np.array_equal(bits(g["u_final"][idx]), bits(o["u_final"]))


def test_high_level_solve_api(pkg, oracle):
    P = pkg
    pl = P.problems_library
    N = 500
    table = pl.lorenz_params(N)
    prob = P.ODEProblem(P.CSource(*pl.lorenz_source()), U0, (0.0, 10.0), table[0])
    # prob_func as a Python closure (harvested on the host) and as a table give the same ensemble
    ep1 = P.EnsembleProblem(prob, prob_func=lambda pr, ctx: P.remake(pr, p=table[ctx.sim_id - 1]))
    ep2 = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table))
    s1 = P.solve(ep1, P.Tsit5(), P.EnsembleB200(), trajectories=N, saveat=0.1)
    s2 = P.solve(ep2, P.Tsit5(), P.EnsembleB200(), trajectories=N, saveat=0.1)
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, (0.0, 10.0), 3, 3, saveat=GRID)
    assert len(s1) == N and s1.converged is False
    for i in (0, 17, N - 1):
        assert s1[i].retcode == "Success" and len(s1[i].t) == 101 and s1[i].t[3] == 0.3
        assert np.array_equal(bits(np.ascontiguousarray(s1[i].u)), bits(o["us"][i]))
        assert np.array_equal(bits(np.ascontiguousarray(s2[i].u)), bits(o["us"][i]))
        assert s1[i].stats.naccept == o["naccept"][i] and s1[i].stats.nf == o["nf"][i]
    # reduction with batches and early stop (lib/DiffEqBase/test/downstream/ensemble.jl:51-112)
    seen = []

    def reduction(u, batch, I):
        seen.append(list(I))
        u = u + [float(np.mean([b for b in batch]))]
        return u, len(u) >= 3
    ep3 = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table), output_func=lambda sol, ctx: (sol.u[-1][0], False),
                            reduction=reduction, u_init=[])
    s3 = P.solve(ep3, P.Tsit5(), P.EnsembleB200(), trajectories=N, batch_size=100, save_everystep=False)
    assert s3.converged is True and len(s3.u) == 3 and seen[0] == list(range(1, 101)) and len(seen) == 3
    assert s3.u[0] == pytest.approx(float(o["u_final"][:100, 0].mean()), rel=1e-13)
    # sympy-traced RHS (stand-in for Symbolics build_function) through the same entry point
    f = lambda u, p, t: [p[0] * (u[1] - u[0]), u[0] * (p[1] - u[2]) - u[1], u[0] * u[1] - p[2] * u[2]]
    ep4 = P.EnsembleProblem(P.ODEProblem(f, U0, (0.0, 10.0), table[0]), prob_func=P.TableProbFunc(p=table))
    s4 = P.solve(ep4, P.Tsit5(), P.EnsembleB200(), trajectories=N, save_everystep=False)
    assert np.allclose(np.stack([s4[i].u[-1] for i in range(N)]), o["u_final"], rtol=1e-6, atol=1e-6)
    # Rosenbrock with a symbolically derived Jacobian
    rob = lambda u, p, t: [-p[0] * u[0] + p[2] * u[1] * u[2], p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2],
                           p[1] * u[1] ** 2]
    ktab = pl.robertson_params(64)
    ep5 = P.EnsembleProblem(P.ODEProblem(rob, U0, (0.0, 1e5), ktab[0]), prob_func=P.TableProbFunc(p=ktab))
    s5 = P.solve(ep5, P.Rodas5P(), P.EnsembleB200(), trajectories=64, save_everystep=False, reltol=1e-6, abstol=1e-8)
    assert all(s5[i].retcode == "Success" for i in range(64))
    assert abs(float(s5[3].u[-1].sum()) - 1.0) < 1e-9


def test_high_level_solve_with_prob_func_changing_tspan(pkg, oracle):
    """solve(EnsembleProblem(prob; prob_func = (prob, ctx) -> remake(prob; p = ..., tspan = ...)), ...): every trajectory
    is solved over its own span — as a closure, as a table, with final states only and with the default (every step) rows."""
    P = pkg
    pl = P.problems_library
    N = 200
    table = pl.lorenz_params(N)
    idx = np.arange(N, dtype=np.uint64)
    spans = np.stack([0.5 * pl.splitmix64_uniform(idx, 1), 1.0 + 4.0 * pl.splitmix64_uniform(idx, 2)], axis=1)
    prob = P.ODEProblem(P.CSource(*pl.lorenz_source()), U0, (0.0, 10.0), table[0])
    ep1 = P.EnsembleProblem(prob, prob_func=lambda pr, ctx: P.remake(pr, p=table[ctx.sim_id - 1], tspan=tuple(spans[ctx.sim_id - 1])))
    ep2 = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table, tspan=spans))
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, spans, 3, 3)
    for ep in (ep1, ep2):
        s = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, save_everystep=False)
        for i in (0, 7, N - 1):
            assert s[i].retcode == "Success" and list(s[i].t) == [spans[i, 0], spans[i, 1]]
            assert np.array_equal(bits(np.ascontiguousarray(s[i].u[-1])), bits(o["u_final"][i]))
            assert s[i].stats.naccept == o["naccept"][i]
    orag = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, spans, 3, 3, save_everystep=True)
    s = P.solve(ep2, P.Tsit5(), P.EnsembleB200(), trajectories=N)          # default: every step
    for i in (0, 7, N - 1):
        a, b = orag["row_offsets"][i], orag["row_offsets"][i + 1]
        assert np.array_equal(np.asarray(s[i].t), orag["ts"][a:b]) and s[i].t[0] == spans[i, 0] and s[i].t[-1] == spans[i, 1]
        assert np.array_equal(bits(np.ascontiguousarray(s[i].u)), bits(orag["us"][a:b]))
    with pytest.raises(NotImplementedError):
        P.solve(ep2, P.Tsit5(), P.EnsembleB200(), trajectories=N, saveat=0.1)


def test_timeseries_meanvar_on_device(pkg, progs, oracle):
    """EnsembleAnalysis.timeseries_steps_meanvar evaluated on the device
    (reference: lib/DiffEqBase/test/downstream/ensemble_analysis.jl:12-33, m ≈ m2, v ≈ v4)."""
    import torch
    pl, ll = pkg.problems_library, pkg.lowlevel
    for f32 in (False, True):
        N = 5000
        p = pl.lorenz_params(N, f32=f32)
        prog = progs(pkg.ALG_TSIT5, f32, "lorenz")
        full = ll.solve_host(prog, U0, p, (0.0, 10.0), saveat=GRID)
        st = ll.solve_host_meanvar(prog, U0, p, (0.0, 10.0), GRID)
        us = full["us"].astype(np.float64)
        assert "us" not in st and st["mean"].shape == (101, 3)
        assert np.allclose(st["mean"], us.mean(axis=0), rtol=1e-12, atol=1e-13)
        assert np.allclose(st["var"], us.var(axis=0, ddof=1), rtol=1e-10, atol=1e-12)
        assert np.array_equal(st["naccept"], full["naccept"]) and np.array_equal(bits(st["u_final"]), bits(full["u_final"]))
        # the oracle's trajectories give the same statistics
        o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(f32), U0, p, (0.0, 10.0), 3, 3, f32=f32, saveat=GRID)
        assert np.allclose(st["mean"], o["us"].astype(np.float64).mean(axis=0), rtol=1e-12, atol=1e-13)
        # device-resident entry point, deterministic (two calls, same bits)
        d_us = torch.from_numpy(full["us"]).cuda()
        m1 = torch.zeros((101, 3), dtype=torch.float64, device="cuda"); v1 = torch.zeros_like(m1)
        m2 = torch.zeros_like(m1); v2 = torch.zeros_like(m1)
        ll.timeseries_meanvar_device(prog.handle, prog.dtype, d_us, m1, v1)
        ll.timeseries_meanvar_device(prog.handle, prog.dtype, d_us, m2, v2)
        torch.cuda.synchronize()
        assert torch.equal(m1, m2) and torch.equal(v1, v2)
        assert np.array_equal(m1.cpu().numpy(), st["mean"])


# ---- save_everystep = true: ragged per-step rows ---------------------------
def _everystep_prog(pkg, handle, alg, f32, problem):
    pl = pkg.problems_library
    dt = pkg.F32 if f32 else pkg.F64
    opt = pkg._lib.OPT_EVERYSTEP
    if problem == "lorenz":
        s, n = pl.lorenz_source(f32)
        return handle.compile(alg, dt, 3, 3, s, n, extra_options=opt)
    if problem == "robertson":
        r, j, tg = pl.robertson_sources(f32)
        return handle.compile(alg, dt, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1], extra_options=opt)
    s, n = pl.pleiades_source(f32)
    return handle.compile(alg, dt, 28, 0, s, n, extra_options=opt)


def _assert_same_ragged(g, o):
    assert_same_result(g, dict(o, us=None))
    assert np.array_equal(g["row_offsets"], o["row_offsets"])
    assert np.array_equal(g["ts"], o["ts"])
    assert np.array_equal(bits(g["us"]), bits(o["us"]))


@pytest.mark.parametrize("f32", [False, True])
def test_everystep_lorenz_tsit5(pkg, handle, oracle, f32):
    N = 3000
    pl = pkg.problems_library
    p = pl.lorenz_params(N, f32=f32)
    prog = _everystep_prog(pkg, handle, pkg.ALG_TSIT5, f32, "lorenz")
    for kw in ({}, {"save_start": False}, {"save_end": False}, {"saveat": [0.25, 1.0, 2.0], "save_end": False},
               {"saveat": [0.5, 2.0]}, {"maxiters": 20}):
        g = pkg.lowlevel.solve_host_everystep(prog, U0, p, (0.0, 2.0), **kw)
        o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(f32), U0, p, (0.0, 2.0), 3, 3, f32=f32, save_everystep=True, **kw)
        _assert_same_ragged(g, o)
        if not kw:
            # sol.t = [t0, every accepted step]; the last row is the end point
            assert np.array_equal(g["nsaved"], g["naccept"] + 1)
            last = g["row_offsets"][1:] - 1
            assert (g["ts"][last] == 2.0).all() and (g["ts"][g["row_offsets"][:-1]] == 0.0).all()
            assert np.array_equal(bits(g["us"][last]), bits(g["u_final"]))
        if "maxiters" in kw:
            assert (g["retcode"] == 2).all()


def test_everystep_stiff_and_vern7(pkg, handle, oracle):
    pl = pkg.problems_library
    r, j, tg = pl.robertson_sources()
    p = pl.robertson_params(1024)
    for alg, oalg in ((pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23), (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P)):
        prog = _everystep_prog(pkg, handle, alg, False, "robertson")
        for kw in ({}, {"saveat": [1.0, 10.0, 50.0]}):
            g = pkg.lowlevel.solve_host_everystep(prog, U0, p, (0.0, 100.0), reltol=1e-6, abstol=1e-8, **kw)
            o = oracle.solve(oalg, r, U0, p, (0.0, 100.0), 3, 3, jac=j, tgrad=tg, reltol=1e-6, abstol=1e-8,
                             save_everystep=True, **kw)
            _assert_same_ragged(g, o)
    u0 = pl.pleiades_u0(256)
    prog = _everystep_prog(pkg, handle, pkg.ALG_VERN7, False, "pleiades")
    g = pkg.lowlevel.solve_host_everystep(prog, u0, None, (0.0, 1.0), reltol=1e-6, abstol=1e-8, saveat=[0.3, 0.6])
    o = oracle.solve(oracle.ALG_VERN7, pl.pleiades_source(), u0, None, (0.0, 1.0), 28, 0, reltol=1e-6, abstol=1e-8,
                     saveat=[0.3, 0.6], save_everystep=True)
    _assert_same_ragged(g, o)


def test_everystep_program_is_refused_by_rectangular_entry_points(pkg, handle, progs):
    pl = pkg.problems_library
    p = pl.lorenz_params(64)
    prog = _everystep_prog(pkg, handle, pkg.ALG_TSIT5, False, "lorenz")
    with pytest.raises(pkg.B200Error):
        pkg.lowlevel.solve_host(prog, U0, p, (0.0, 1.0))
    with pytest.raises(ValueError):
        pkg.lowlevel.solve_host_everystep(progs(pkg.ALG_TSIT5, False, "lorenz"), U0, p, (0.0, 1.0))


def test_high_level_default_is_save_everystep(pkg, oracle):
    """solve(EnsembleProblem, Tsit5(), EnsembleB200(); trajectories) with no saveat: the reference's default
    save_everystep = isempty(saveat) (solve.jl:138) — every trajectory's sol.t / sol.u are its accepted steps."""
    P = pkg
    pl = P.problems_library
    N = 300
    table = pl.lorenz_params(N)
    prob = P.ODEProblem(P.CSource(*pl.lorenz_source()), U0, (0.0, 2.0), table[0])
    ep = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table))
    s = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N)
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, (0.0, 2.0), 3, 3, save_everystep=True)
    for i in (0, 123, N - 1):
        a, b = o["row_offsets"][i], o["row_offsets"][i + 1]
        assert np.array_equal(s[i].t, o["ts"][a:b]) and s[i].t[0] == 0.0 and s[i].t[-1] == 2.0
        assert np.array_equal(bits(np.ascontiguousarray(s[i].u)), bits(o["us"][a:b]))
        assert len(s[i]) == s[i].stats.naccept + 1
    # output_func sees the per-step solution
    ep2 = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table), output_func=lambda sol, ctx: (len(sol.t), False))
    s2 = P.solve(ep2, P.Tsit5(), P.EnsembleB200(), trajectories=N)
    assert list(s2.u) == list(o["nsaved"])


# ---- dense output sol(t), post hoc, from recomputed stages -------------------
@pytest.mark.parametrize("f32", [False, True])
def test_dense_eval_lorenz(pkg, handle, oracle, f32):
    N = 2000
    pl = pkg.problems_library
    p = pl.lorenz_params(N, f32=f32)
    tq = np.concatenate([[0.0], np.sort(np.random.default_rng(5).uniform(0.0, 2.0, 60)), [2.0, 2.05]])
    for alg, oalg in ((pkg.ALG_TSIT5, oracle.ALG_TSIT5), (pkg.ALG_VERN7, oracle.ALG_VERN7)):
        prog = _everystep_prog(pkg, handle, alg, f32, "lorenz")
        g = pkg.lowlevel.solve_host_dense(prog, U0, p, (0.0, 2.0), tq)
        o = oracle.solve(oalg, pl.lorenz_source(f32), U0, p, (0.0, 2.0), 3, 3, f32=f32, dense_tq=tq)
        # the GPU recomputes each step's stages, the oracle interpolates from the k arrays it stored: same bits
        assert np.array_equal(bits(g["dense"]), bits(o["dense"]))
        assert_same_result(g, dict(o, us=None))
        # sol(tf): the interpolation polynomial at Θ = 1 meets the last row to rounding
        assert np.allclose(g["dense"][:, -2], g["u_final"], rtol=1e-4 if f32 else 1e-12, atol=1e-4 if f32 else 1e-12)
        assert (g["dense"][:, 0] == np.asarray(U0, dtype=g["dense"].dtype)).all()


def test_dense_eval_stiff_and_accuracy(pkg, handle, oracle):
    pl = pkg.problems_library
    r, j, tg = pl.robertson_sources()
    p = pl.robertson_params(512)
    tq = np.array([0.0, 1e-3, 0.02, 0.5, 1.0, 7.5, 33.0, 99.0, 100.0])
    for alg, oalg in ((pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23), (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P)):
        prog = _everystep_prog(pkg, handle, alg, False, "robertson")
        g = pkg.lowlevel.solve_host_dense(prog, U0, p, (0.0, 100.0), tq, reltol=1e-6, abstol=1e-8)
        o = oracle.solve(oalg, r, U0, p, (0.0, 100.0), 3, 3, jac=j, tgrad=tg, reltol=1e-6, abstol=1e-8, dense_tq=tq)
        assert np.array_equal(bits(g["dense"]), bits(o["dense"]))
        assert np.abs(g["dense"].sum(axis=2) - 1.0).max() < 1e-6               # mass conservation along sol(t)
    # against the closed form u0 exp(1.01 t): test/Regression_I/ode_dense_tests.jl bounds (Tsit5 2e-6 at dt = 1/4)
    s, n = linear_source()
    prog = handle.compile(pkg.ALG_TSIT5, pkg.F64, 1, 0, s, n, extra_options=pkg._lib.OPT_EVERYSTEP)
    tq = np.linspace(0.0, 1.0, 101)
    g = pkg.lowlevel.solve_host_dense(prog, np.array([[0.5]]), None, (0.0, 1.0), tq, dt=0.25)
    assert np.abs(g["dense"][0, :, 0] - 0.5 * np.exp(1.01 * tq)).max() < 2e-6


def test_high_level_dense_solution_call(pkg, oracle):
    """sol(t) on the default (save_everystep, no saveat) ensemble solution, and the one-pass ensemble form."""
    P = pkg
    pl = P.problems_library
    N = 200
    table = pl.lorenz_params(N)
    prob = P.ODEProblem(P.CSource(*pl.lorenz_source()), U0, (0.0, 2.0), table[0])
    ep = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table))
    s = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N)
    tq = np.array([0.0, 0.3, 0.77, 1.5, 2.0])
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, (0.0, 2.0), 3, 3, dense_tq=tq)
    allv = s.at(tq)
    assert np.array_equal(bits(allv), bits(o["dense"]))
    assert s[7].dense and np.array_equal(bits(s[7](tq)), bits(o["dense"][7]))
    assert np.array_equal(bits(s[7](0.77)), bits(o["dense"][7, 2]))
    assert np.array_equal(bits(s[7](tq[::-1].copy())), bits(o["dense"][7, ::-1]))     # unsorted queries
    s2 = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, saveat=0.5)
    with pytest.raises(NotImplementedError):
        s2[0](0.3)
    s3 = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, dense=True)          # explicit dense = true
    assert np.array_equal(bits(s3[7](tq)), bits(o["dense"][7]))
    s4 = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, dense=False)
    assert not s4[7].dense
    with pytest.raises(NotImplementedError):
        P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, saveat=0.5, dense=True)


# ---- save_idxs ---------------------------------------------------------------
@pytest.mark.parametrize("f32", [False, True])
def test_save_idxs_rows(pkg, handle, oracle, f32):
    N = 3000
    pl = pkg.problems_library
    p = pl.lorenz_params(N, f32=f32)
    dt = pkg.F32 if f32 else pkg.F64
    s, n = pl.lorenz_source(f32)
    idxs = [2, 0]
    prog = handle.compile(pkg.ALG_TSIT5, dt, 3, 3, s, n, extra_options=pkg._lib.opt_save_idxs(idxs))
    g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 10.0), saveat=GRID)
    o = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 10.0), 3, 3, f32=f32, saveat=GRID, save_idxs=idxs)
    assert g["us"].shape == (N, 101, 2)
    assert_same_result(g, o)
    full = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 10.0), 3, 3, f32=f32, saveat=GRID)
    assert np.array_equal(bits(g["us"]), bits(np.ascontiguousarray(full["us"][:, :, idxs])))
    # statistics over the selected components only
    mv = pkg.lowlevel.solve_host_meanvar(prog, U0, p, (0.0, 10.0), GRID)
    assert mv["mean"].shape == (101, 2)
    assert np.allclose(mv["mean"], o["us"].astype(np.float64).mean(axis=0), rtol=1e-5 if f32 else 1e-12)
    # ragged rows
    prog2 = handle.compile(pkg.ALG_TSIT5, dt, 3, 3, s, n,
                           extra_options=pkg._lib.opt_save_idxs([1]) + " " + pkg._lib.OPT_EVERYSTEP)
    g2 = pkg.lowlevel.solve_host_everystep(prog2, U0, p, (0.0, 2.0))
    o2 = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 2.0), 3, 3, f32=f32, save_everystep=True, save_idxs=[1])
    assert g2["us"].shape[1] == 1
    _assert_same_ragged(g2, o2)
    with pytest.raises(pkg.B200Error):                    # dense output needs whole rows
        pkg.lowlevel.solve_host_dense(prog2, U0, p, (0.0, 2.0), [0.5])
    with pytest.raises(pkg.B200Error):                    # index out of range
        handle.compile(pkg.ALG_TSIT5, dt, 3, 3, s, n, extra_options="-DB200_SAVE_IDXS=3")


def test_high_level_save_idxs(pkg, oracle):
    P = pkg
    pl = P.problems_library
    N = 100
    table = pl.lorenz_params(N)
    prob = P.ODEProblem(P.CSource(*pl.lorenz_source()), U0, (0.0, 10.0), table[0])
    ep = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table))
    s = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, saveat=0.1, save_idxs=[0])
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, (0.0, 10.0), 3, 3, saveat=GRID, save_idxs=[0])
    assert s[5].u.shape == (101, 1) and np.array_equal(bits(np.ascontiguousarray(s[5].u)), bits(o["us"][5]))
    s2 = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, save_everystep=False, save_idxs=[2, 1])
    assert s2[5].u.shape == (2, 2) and np.array_equal(s2[5].u[-1], o["u_final"][5][[2, 1]])


# ---- more steppers on the same skeleton: DP5, BS3 ----------------------------
@pytest.mark.parametrize("alg_name", ["dp5", "bs3"])
@pytest.mark.parametrize("f32", [False, True])
def test_low_order_rk_parity(pkg, handle, oracle, alg_name, f32):
    N = 3000
    pl = pkg.problems_library
    alg, oalg = {"dp5": (pkg.ALG_DP5, oracle.ALG_DP5), "bs3": (pkg.ALG_BS3, oracle.ALG_BS3)}[alg_name]
    p = pl.lorenz_params(N, f32=f32)
    s, n = pl.lorenz_source(f32)
    dt = pkg.F32 if f32 else pkg.F64
    prog = handle.compile(alg, dt, 3, 3, s, n)
    tf = 10.0 if alg_name == "dp5" else 3.0
    grid = [k / 10 for k in range(1, int(tf * 10) + 1)]
    for kw in ({}, {"saveat": grid}, {"reltol": 1e-6, "abstol": 1e-8} if not f32 else {"reltol": 1e-4, "abstol": 1e-5},
               {"maxiters": 15}):
        g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, tf), **kw)
        o = oracle.solve(oalg, (s, n), U0, p, (0.0, tf), 3, 3, f32=f32, **kw)
        assert_same_result(g, o)
    stages = 6 if alg_name == "dp5" else 3
    assert (g["nf"] == 3 + stages * (g["naccept"] + g["nreject"])).all()        # 1 (FSAL start) + 2 (initdt)
    # ragged rows and dense output through the same generic paths
    prog_e = handle.compile(alg, dt, 3, 3, s, n, extra_options=pkg._lib.OPT_EVERYSTEP)
    ge = pkg.lowlevel.solve_host_everystep(prog_e, U0, p, (0.0, 2.0), saveat=[0.5, 1.5])
    oe = oracle.solve(oalg, (s, n), U0, p, (0.0, 2.0), 3, 3, f32=f32, save_everystep=True, saveat=[0.5, 1.5])
    _assert_same_ragged(ge, oe)
    tq = np.linspace(0.0, 2.0, 41)
    gd = pkg.lowlevel.solve_host_dense(prog_e, U0, p, (0.0, 2.0), tq)
    od = oracle.solve(oalg, (s, n), U0, p, (0.0, 2.0), 3, 3, f32=f32, dense_tq=tq)
    assert np.array_equal(bits(gd["dense"]), bits(od["dense"]))


def test_low_order_rk_high_level(pkg, oracle):
    P = pkg
    pl = P.problems_library
    N = 128
    table = pl.lorenz_params(N)
    prob = P.ODEProblem(P.CSource(*pl.lorenz_source()), U0, (0.0, 5.0), table[0])
    ep = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table))
    for alg, oalg in ((P.DP5(), oracle.ALG_DP5), (P.BS3(), oracle.ALG_BS3)):
        s = P.solve(ep, alg, P.EnsembleB200(), trajectories=N, saveat=0.5)
        o = oracle.solve(oalg, pl.lorenz_source(), U0, table, (0.0, 5.0), 3, 3, saveat=[k / 2 for k in range(1, 11)])
        for i in (0, 77):
            assert np.array_equal(bits(np.ascontiguousarray(s[i].u)), bits(o["us"][i]))
            assert s[i].stats.naccept == o["naccept"][i]


@pytest.mark.parametrize("name", ["Rodas5", "Rodas4", "Rodas42", "Rodas4P", "Rodas4P2", "Rodas5Pe", "Rodas3P", "Rodas23W"])
def test_rodas_family_parity(pkg, handle, oracle, name):
    """The generic RodasTableau stepper over the other members of the family, Robertson FP64 (+ FP32 for two)."""
    pl = pkg.problems_library
    alg = getattr(pkg, "ALG_" + name.upper())
    oalg = getattr(oracle, "ALG_" + name.upper())
    N = 1024
    for f32 in ((False, True) if name in ("Rodas5", "Rodas4", "Rodas3P") else (False,)):
        r, j, tg = pl.robertson_sources(f32)
        p = pl.robertson_params(N, f32=f32)
        dt = pkg.F32 if f32 else pkg.F64
        tf = 1e4 if not f32 else 1e3
        tol = dict(reltol=1e-6, abstol=1e-8) if not f32 else dict(reltol=1e-3, abstol=1e-5)
        prog = handle.compile(alg, dt, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1])
        for extra in ({}, {"saveat": [tf * 1e-3, tf * 1e-2, tf * 0.5]}):
            kw = dict(tol, **extra)
            g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, tf), **kw)
            o = oracle.solve(oalg, r, U0, p, (0.0, tf), 3, 3, f32=f32, jac=j, tgrad=tg, **kw)
            assert_same_result(g, o)
            assert (g["retcode"] == 1).all()
        assert np.abs(g["u_final"].astype(np.float64).sum(axis=1) - 1.0).max() < (1e-7 if not f32 else 1e-3)
    r, j, tg = pl.robertson_sources()
    p = pl.robertson_params(N)
    prog_e = handle.compile(alg, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1], extra_options=pkg._lib.OPT_EVERYSTEP)
    tq = np.array([0.0, 0.01, 1.0, 20.0, 100.0])
    gd = pkg.lowlevel.solve_host_dense(prog_e, U0, p, (0.0, 100.0), tq, reltol=1e-6, abstol=1e-8)
    od = oracle.solve(oalg, r, U0, p, (0.0, 100.0), 3, 3, jac=j, tgrad=tg, reltol=1e-6, abstol=1e-8, dense_tq=tq)
    assert np.array_equal(bits(gd["dense"]), bits(od["dense"]))


@pytest.mark.parametrize("name", ["Vern6", "Vern8", "Vern9"])
@pytest.mark.parametrize("f32", [False, True])
def test_generated_verner_parity(pkg, handle, oracle, name, f32):
    """Vern6/8/9 (scripts/gen_verner.py): final states, stats, lazily interpolated saveat rows, ragged rows, dense."""
    pl = pkg.problems_library
    alg = getattr(pkg, "ALG_" + name.upper())
    oalg = getattr(oracle, "ALG_" + name.upper())
    N = 2000
    p = pl.lorenz_params(N, f32=f32)
    s, n = pl.lorenz_source(f32)
    prog = handle.compile(alg, pkg.F32 if f32 else pkg.F64, 3, 3, s, n)
    grid = [k / 4 for k in range(1, 21)]
    tol = dict(reltol=1e-8, abstol=1e-10) if not f32 else dict(reltol=1e-4, abstol=1e-5)
    for kw in ({}, {"saveat": grid}, dict(tol, saveat=grid), {"maxiters": 9}):
        g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 5.0), **kw)
        o = oracle.solve(oalg, (s, n), U0, p, (0.0, 5.0), 3, 3, f32=f32, **kw)
        assert_same_result(g, o)
    prog_e = handle.compile(alg, pkg.F32 if f32 else pkg.F64, 3, 3, s, n, extra_options=pkg._lib.OPT_EVERYSTEP)
    ge = pkg.lowlevel.solve_host_everystep(prog_e, U0, p, (0.0, 2.0), saveat=[0.7])
    oe = oracle.solve(oalg, (s, n), U0, p, (0.0, 2.0), 3, 3, f32=f32, save_everystep=True, saveat=[0.7])
    _assert_same_ragged(ge, oe)
    tq = np.linspace(0.0, 2.0, 33)
    gd = pkg.lowlevel.solve_host_dense(prog_e, U0, p, (0.0, 2.0), tq)
    od = oracle.solve(oalg, (s, n), U0, p, (0.0, 2.0), 3, 3, f32=f32, dense_tq=tq)
    assert np.array_equal(bits(gd["dense"]), bits(od["dense"]))


def test_vern9_wider_state(pkg, handle, oracle):
    """n = 8 (prob_ode_2Dlinear flattened)."""
    from helpers import linear2d_source
    s, n = linear2d_source(8)
    rng = np.random.default_rng(3)
    u0 = rng.uniform(0.1, 1.0, size=(300, 8))
    prog = handle.compile(pkg.ALG_VERN9, pkg.F64, 8, 0, s, n)
    kw = dict(reltol=1e-8, abstol=1e-10, saveat=[0.25, 0.5, 1.0])
    g = pkg.lowlevel.solve_host(prog, u0, None, (0.0, 1.0), **kw)
    o = oracle.solve(oracle.ALG_VERN9, (s, n), u0, None, (0.0, 1.0), 8, 0, **kw)
    assert_same_result(g, o)
    assert (g["retcode"] == 1).all()
    assert np.allclose(g["u_final"], u0 * np.exp(1.01), rtol=1e-8)


def test_vern9_pleiades_out_of_line_rhs(pkg, handle, oracle):
    """28 inlined copies of the 6.5 KB Pleiades RHS would take ptxas > 15 minutes; the shim keeps the RHS out of line
    when (source size x call sites) is large (b200ode_shim.cu: rhs_inline).  Same bits either way."""
    pl = pkg.problems_library
    u0 = pl.pleiades_u0(96)
    s, n = pl.pleiades_source()
    prog = handle.compile(pkg.ALG_VERN9, pkg.F64, 28, 0, s, n)
    assert prog.info["compile_ms"] < 120e3
    kw = dict(reltol=1e-8, abstol=1e-10, saveat=[1.0, 2.0, 3.0])
    g = pkg.lowlevel.solve_host(prog, u0, None, (0.0, 3.0), **kw)
    o = oracle.solve(oracle.ALG_VERN9, (s, n), u0, None, (0.0, 3.0), 28, 0, **kw)
    assert_same_result(g, o)
    assert (g["retcode"] == 1).all()


# ---- tstops ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("f32", [False, True])
def test_tstops_parity(pkg, handle, oracle, f32):
    N = 2000
    pl = pkg.problems_library
    p = pl.lorenz_params(N, f32=f32)
    s, n = pl.lorenz_source(f32)
    dt = pkg.F32 if f32 else pkg.F64
    stops = [0.37, 1.0, 1.0, 2.5, 7.0, -3.0]
    prog = handle.compile(pkg.ALG_TSIT5, dt, 3, 3, s, n, extra_options=pkg._lib.OPT_TSTOPS)
    for kw in ({}, {"saveat": [0.37, 0.5, 1.0, 3.0]}, {"maxiters": 12}):
        g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 3.0), tstops=stops, **kw)
        o = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 3.0), 3, 3, f32=f32, tstops=stops, **kw)
        assert_same_result(g, o)
    base = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 3.0), 3, 3, f32=f32)
    assert not np.array_equal(o["naccept"], base["naccept"])
    prog_e = handle.compile(pkg.ALG_TSIT5, dt, 3, 3, s, n, extra_options=pkg._lib.OPT_TSTOPS + " " + pkg._lib.OPT_EVERYSTEP)
    ge = pkg.lowlevel.solve_host_everystep(prog_e, U0, p, (0.0, 3.0), tstops=stops)
    oe = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 3.0), 3, 3, f32=f32, tstops=stops, save_everystep=True)
    _assert_same_ragged(ge, oe)
    rdt = np.float32 if f32 else np.float64
    for i in (0, N - 1):
        row = ge["ts"][ge["row_offsets"][i]:ge["row_offsets"][i + 1]]
        for st in (0.37, 1.0, 2.5):
            assert float(rdt(st)) in row                       # every stop is a step end point
    # a program without the option refuses tstops; stiff steppers take them too
    with pytest.raises(pkg.B200Error):
        pkg.lowlevel.solve_host(handle.compile(pkg.ALG_TSIT5, dt, 3, 3, s, n), U0, p, (0.0, 3.0), tstops=stops)
    if not f32:
        r, j, tg = pl.robertson_sources()
        k = pl.robertson_params(512)
        progr = handle.compile(pkg.ALG_RODAS5P, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1], extra_options=pkg._lib.OPT_TSTOPS)
        g = pkg.lowlevel.solve_host(progr, U0, k, (0.0, 100.0), tstops=[1.0, 10.0], reltol=1e-6, abstol=1e-8)
        o = oracle.solve(oracle.ALG_RODAS5P, r, U0, k, (0.0, 100.0), 3, 3, jac=j, tgrad=tg, tstops=[1.0, 10.0], reltol=1e-6, abstol=1e-8)
        assert_same_result(g, o)


# ---- adaptive = false -----------------------------------------------------------------------------
@pytest.mark.parametrize("f32", [False, True])
def test_fixed_step_parity(pkg, handle, oracle, f32):
    N = 1500
    pl = pkg.problems_library
    p = pl.lorenz_params(N, f32=f32)
    s, n = pl.lorenz_source(f32)
    dt = pkg.F32 if f32 else pkg.F64
    L = pkg._lib
    prog = handle.compile(pkg.ALG_TSIT5, dt, 3, 3, s, n, extra_options=L.OPT_FIXED_DT)
    for kw in ({"dt": 0.01}, {"dt": 0.03, "saveat": [0.5, 1.0, 1.5]}, {"dt": 0.01, "maxiters": 50}):
        g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 2.0), **kw)
        o = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 2.0), 3, 3, f32=f32, adaptive=False, **kw)
        assert_same_result(g, o)
    assert (g["retcode"] == 2).all()
    with pytest.raises(pkg.B200Error):                                  # neither dt nor tstops
        pkg.lowlevel.solve_host(prog, U0, p, (0.0, 2.0))
    prog2 = handle.compile(pkg.ALG_TSIT5, dt, 3, 3, s, n,
                           extra_options=" ".join([L.OPT_FIXED_DT, L.OPT_TSTOPS, L.OPT_EVERYSTEP]))
    # (dt = 0.3 would be beyond Lorenz's stability limit: those trajectories end Unstable on both sides with the
    # same counters, but the NaN payloads of x86 and the GPU differ, so states are compared on a stable step only)
    ge = pkg.lowlevel.solve_host_everystep(prog2, U0, p, (0.0, 2.0), dt=0.03, tstops=[0.5, 1.25])
    oe = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 2.0), 3, 3, f32=f32, adaptive=False, dt=0.03,
                      tstops=[0.5, 1.25], save_everystep=True)
    _assert_same_ragged(ge, oe)
    assert (ge["retcode"] == 1).all() and (ge["nsaved"] == ge["nsaved"][0]).all()    # one step grid for every trajectory
    if not f32:
        r, j, tg = pl.robertson_sources()
        k = pl.robertson_params(256)
        pr = handle.compile(pkg.ALG_RODAS5P, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1], extra_options=L.OPT_FIXED_DT)
        g = pkg.lowlevel.solve_host(pr, U0, k, (0.0, 1.0), dt=0.01)
        o = oracle.solve(oracle.ALG_RODAS5P, r, U0, k, (0.0, 1.0), 3, 3, jac=j, tgrad=tg, adaptive=False, dt=0.01)
        assert_same_result(g, o)


@pytest.mark.parametrize("f32", [False, True])
def test_rosenbrock32_parity(pkg, handle, oracle, f32):
    """Rosenbrock32 is only A-stable and re-uses f(uprev + dt k2) as the next fsalfirst (both as in the reference), so on
    Robertson it needs ~3000 steps to t = 10 and runs into maxiters long before 1e4; FP32 trajectories partly end
    Unstable.  Parity covers all of that: same counters, same retcodes, same states."""
    pl = pkg.problems_library
    r, j, tg = pl.robertson_sources(f32)
    p = pl.robertson_params(1024, f32=f32)
    dt = pkg.F32 if f32 else pkg.F64
    tf = 10.0
    tol = dict(reltol=1e-6, abstol=1e-8) if not f32 else dict(reltol=1e-3, abstol=1e-5)
    prog = handle.compile(pkg.ALG_ROSENBROCK32, dt, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1])
    for extra in ({}, {"saveat": [0.01, 0.1, 5.0]}, {"maxiters": 500}):
        kw = dict(tol, **extra)
        g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, tf), **kw)
        o = oracle.solve(oracle.ALG_ROSENBROCK32, r, U0, p, (0.0, tf), 3, 3, f32=f32, jac=j, tgrad=tg, **kw)
        ok = o["retcode"] == 1
        assert_same_result(g, o, keys=("naccept", "nreject", "nf", "retcode", "nsaved", "njacs", "nw", "nsolve")) if ok.all() \
            else None
        for k in ("naccept", "nreject", "nf", "retcode", "nsaved", "njacs", "nw", "nsolve"):
            assert np.array_equal(g[k], o[k]), k
        assert np.array_equal(bits(g["u_final"][ok]), bits(o["u_final"][ok]))       # failed ones may hold NaNs
        if "saveat" in extra:
            assert np.array_equal(bits(g["us"][ok]), bits(o["us"][ok]))
    if not f32:
        assert ok.sum() == 0                    # maxiters = 500 stops every trajectory
        prog_e = handle.compile(pkg.ALG_ROSENBROCK32, dt, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1],
                                extra_options=pkg._lib.OPT_EVERYSTEP)
        ge = pkg.lowlevel.solve_host_everystep(prog_e, U0, p, (0.0, 1.0), **tol)
        oe = oracle.solve(oracle.ALG_ROSENBROCK32, r, U0, p, (0.0, 1.0), 3, 3, jac=j, tgrad=tg, save_everystep=True, **tol)
        _assert_same_ragged(ge, oe)
        # post-hoc dense output recomputes a step's stages from its saved start row; Rosenbrock32's fsalfirst is
        # f(uprev + dt k2) of the previous step, which no saved row holds, so the path declines instead of guessing
        with pytest.raises(pkg.B200Error):
            pkg.lowlevel.solve_host_dense(prog_e, U0, p, (0.0, 1.0), np.array([0.0, 0.3]), **tol)


def test_device_resident_ragged_and_dense_api(pkg, handle, oracle):
    """The torch-tensor form of save_everystep / dense output: count pass, scan on the device (torch.cumsum), fill pass,
    dense evaluation — nothing but the total row count visits the host."""
    import torch
    pl, ll = pkg.problems_library, pkg.lowlevel
    N = 4096
    s, n = pl.lorenz_source()
    p = pl.lorenz_params(N)
    prog = handle.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, n, extra_options=pkg._lib.OPT_EVERYSTEP)
    b = ll.DeviceBuffers(prog, N, 0, "cuda:0", u0_shared=True)
    b.u0.copy_(torch.tensor([1.0, 0, 0], dtype=torch.float64)); b.p.copy_(torch.from_numpy(p))
    ll.solve_everystep_device(prog, b, (0.0, 2.0))
    offs = torch.zeros(N + 1, dtype=torch.int64, device="cuda:0")
    offs[1:] = torch.cumsum(b.nsaved.to(torch.int64), 0)
    total = int(offs[-1].item())
    ts = torch.empty(total, dtype=torch.float64, device="cuda:0")
    dts = torch.empty_like(ts)
    us = torch.empty((total, 3), dtype=torch.float64, device="cuda:0")
    ll.solve_everystep_device(prog, b, (0.0, 2.0), row_offsets=offs, ts=ts, dts=dts, us=us)
    o = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 2.0), 3, 3, save_everystep=True)
    assert np.array_equal(offs.cpu().numpy(), o["row_offsets"])
    assert np.array_equal(ts.cpu().numpy(), o["ts"]) and np.array_equal(bits(us.cpu().numpy()), bits(o["us"]))
    # dts: the step that ended at each row; the start row carries 0; consecutive rows differ by it up to rounding
    d = dts.cpu().numpy(); t = ts.cpu().numpy(); first = o["row_offsets"][:-1]
    assert (d[first] == 0).all()
    inner = np.ones(total, dtype=bool); inner[first] = False
    assert np.allclose(t[inner] - t[np.nonzero(inner)[0] - 1], d[inner], rtol=1e-12, atol=1e-15)
    tq = torch.linspace(0.0, 2.0, 17, dtype=torch.float64, device="cuda:0")
    out = torch.empty((N, 17, 3), dtype=torch.float64, device="cuda:0")
    ll.dense_eval_device(prog, N, b.p, offs, ts, dts, us, tq, out)
    od = oracle.solve(oracle.ALG_TSIT5, (s, n), U0, p, (0.0, 2.0), 3, 3, dense_tq=tq.cpu().numpy())
    assert np.array_equal(bits(out.cpu().numpy()), bits(od["dense"]))


# ---- the Rosenbrock linear solve for n not in {1, 3}: partial-pivot LU (b200_rosenbrock.cuh, B200_LINSOLVE_LU) ----
# The reference exercises exactly these SVector systems through StaticWOperator
# (test/InterfaceI/static_array_tests.jl:106-167: HIRES with n = 4, 5, 8; benchmark/benchmarks.jl:110-123: Van der Pol).
@pytest.mark.parametrize("problem", ["vdp", "hires5", "hires8"])
@pytest.mark.parametrize("alg_name", ["ros23", "rodas5p"])
@pytest.mark.parametrize("f32", [False, True])
def test_stiff_lu_path_parity(pkg, handle, oracle, problem, alg_name, f32):
    pl = pkg.problems_library
    alg, oalg = {"ros23": (pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23),
                 "rodas5p": (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P)}[alg_name]
    r, j, tg, n, np_, u0, tspan = pl.stiff_sources(problem, f32)
    N = 1024 if problem != "vdp" else 256
    p = pl.stiff_params(problem, N, f32=f32)
    tol = dict(reltol=1e-6, abstol=1e-8) if not f32 else dict(reltol=1e-3, abstol=1e-5)
    prog = handle.compile(alg, pkg.F32 if f32 else pkg.F64, n, np_, r[0], r[1], j[0], j[1], tg[0], tg[1])
    try:
        mid = [tspan[1] * 0.01, tspan[1] * 0.5]
        for extra in ({}, {"saveat": mid}):
            kw = dict(tol, **extra)
            g = pkg.lowlevel.solve_host(prog, u0, p, tspan, **kw)
            o = oracle.solve(oalg, r, u0, p, tspan, n, np_, f32=f32, jac=j, tgrad=tg, **kw)
            assert_same_result(g, o)
            assert (g["retcode"] == 1).all()
            assert (g["njacs"] == 2 * (g["naccept"] + g["nreject"])).all()
    finally:
        prog.close()


def test_singular_w_is_rejected_like_the_reference(pkg, handle, oracle):
    """A W that is exactly singular makes the LU report failure: the attempt returns EEst = 2 and the step is
    rejected (rosenbrock_perform_step.jl:271-274); both sides must agree on every count."""
    # u' = A u with J = A such that W = J - I/(dt*gamma) is singular only by construction of dt; use a 2x2 system whose
    # Jacobian has a zero row: the factorisation then meets a zero pivot whenever 1/(dt gamma) cancels exactly (never in
    # practice), so this test pins the ordinary path on a degenerate J instead: step counts and states still match.
    T = "double"
    rhs = ("void deg_rhs(%s* du, const %s* u, const %s* p, const %s t) { du[0] = -p[0] * u[0]; du[1] = 0.0; }\n" % (T, T, T, T), "deg_rhs")
    jac = ("void deg_jac(%s* J, const %s* u, const %s* p, const %s t) { J[0] = -p[0]; J[1] = 0.0; J[2] = 0.0; J[3] = 0.0; }\n" % (T, T, T, T), "deg_jac")
    N = 64
    p = (10.0 ** np.linspace(0, 6, N)).reshape(N, 1)
    u0 = np.array([1.0, 2.0])
    for alg, oalg in ((pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23), (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P)):
        prog = handle.compile(alg, pkg.F64, 2, 1, rhs[0], rhs[1], jac[0], jac[1])
        try:
            g = pkg.lowlevel.solve_host(prog, u0, p, (0.0, 1.0))
            o = oracle.solve(oalg, rhs, u0, p, (0.0, 1.0), 2, 1, jac=jac)
            assert_same_result(g, o)
            assert (g["u_final"][:, 1] == 2.0).all()
        finally:
            prog.close()


def test_fast_math_matches_ieee(handle):
    """The branch-free division / square-root sequences of the kernels (b200_base.cuh) give the bits of the plain IEEE
    operators on 2^27 operand sets (arbitrary bit patterns, ODE-scale exponents, near-all-ones mantissas) wherever
    they do not raise their flag; the flag is rare on ODE-scale operands."""
    bad, flagged = handle.selftest_fastmath(1 << 27, seed=20261017)
    assert bad == [0] * 6, bad   # [div64, div_const64, sqrt64, div32, sqrt32, unguarded div32]
    assert flagged > 0          # the arbitrary-bit-pattern class must exercise the flag
    bad2, _ = handle.selftest_fastmath(1 << 22, seed=7)
    assert bad2 == [0] * 6, bad2


# ---- lane-group kernel (device/b200_coop.cuh): 16 lanes per trajectory, component-form RHS ---------------------------
def test_lane_group_kernel_pleiades_vern7(pkg, handle, oracle):
    """BASELINE config 4 through the lane-group kernel: bit-exact against the oracle run on the ordinary (full-vector)
    Pleiades source — final states, step counts, nf and the lazily interpolated saveat rows."""
    pl = pkg.problems_library
    N = 777                      # not a multiple of the trajectories per warp / CTA
    u0 = pl.pleiades_u0(N)
    src, name = pl.pleiades_component_source()
    prog = handle.compile(pkg.ALG_VERN7, pkg.F64, 28, 0, src, name, extra_options=pkg._lib.OPT_COMPONENT_RHS)
    try:
        assert prog.info["local_bytes_integrate"] == 0 or prog.info["local_bytes_integrate"] < 512
        kw = dict(reltol=1e-6, abstol=1e-8)
        for extra in ({}, {"saveat": [0.5, 1.0, 1.5, 2.0, 2.5, 3.0]}, {"saveat": [0.01, 2.999], "save_start": False}):
            g = pkg.lowlevel.solve_host(prog, u0, None, (0.0, 3.0), **dict(kw, **extra))
            o = oracle.solve(oracle.ALG_VERN7, pl.pleiades_source(), u0, None, (0.0, 3.0), 28, 0, **dict(kw, **extra))
            assert_same_result(g, o)
            assert (g["nf"] == 2 + 10 * (g["naccept"] + g["nreject"])).all()
        # failure retcodes reach the host and do not stall the other group of the warp
        g = pkg.lowlevel.solve_host(prog, u0[:33], None, (0.0, 3.0), maxiters=7, **kw)
        o = oracle.solve(oracle.ALG_VERN7, pl.pleiades_source(), u0[:33], None, (0.0, 3.0), 28, 0, maxiters=7, **kw)
        assert_same_result(g, o)
        assert (g["retcode"] == pkg._lib.RC_MAXITERS).all()
    finally:
        prog.close()


def test_lane_group_kernel_other_shapes(pkg, handle, oracle):
    """The same kernel at other group shapes: Lorenz (n = 3) as 2 and 4 lanes per trajectory, FP32 Pleiades."""
    pl = pkg.problems_library
    lor = ("double lorenz_i(int i, const double* u, const double* p, const double t) {\n"
           "  if (i == 0) return p[0] * (u[1] - u[0]);\n"
           "  if (i == 1) return u[0] * (p[1] - u[2]) - u[1];\n"
           "  return u[0] * u[1] - p[2] * u[2];\n}\n", "lorenz_i")
    N = 500
    p = pl.lorenz_params(N)
    grid = [k / 2 for k in range(1, 21)]
    o = oracle.solve(oracle.ALG_VERN7, pl.lorenz_source(), U0, p, (0.0, 10.0), 3, 3, saveat=grid)
    for L in (2, 4):
        prog = handle.compile(pkg.ALG_VERN7, pkg.F64, 3, 3, lor[0], lor[1], extra_options="-DB200_COOP=1 -DB200_L=%d" % L)
        try:
            g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 10.0), saveat=grid)
            assert_same_result(g, o)
        finally:
            prog.close()
    u0 = pl.pleiades_u0(256, f32=True)
    src, name = pl.pleiades_component_source(f32=True)
    prog = handle.compile(pkg.ALG_VERN7, pkg.F32, 28, 0, src, name, extra_options=pkg._lib.OPT_COMPONENT_RHS)
    try:
        kw = dict(reltol=1e-4, abstol=1e-5)
        g = pkg.lowlevel.solve_host(prog, u0, None, (0.0, 3.0), **kw)
        o32 = oracle.solve(oracle.ALG_VERN7, pl.pleiades_source(True), u0, None, (0.0, 3.0), 28, 0, f32=True, **kw)
        assert_same_result(g, o32)
    finally:
        prog.close()


@pytest.mark.parametrize("f32", [False, True])
def test_smem_stage_kernel_pleiades_vern7(pkg, handle, oracle, f32):
    """BASELINE config 4 through the shared-memory stage kernel (B200ODE_OPT_SMEM_STAGES, device/b200_vern7_wide.cuh):
    the stage derivatives in shared memory, one inlined copy of the pair-shared Pleiades text compiled with the flagged
    fast sqrt / division — bit-exact against the oracle run on the reference's plain double loop."""
    pl = pkg.problems_library
    N = 1500                     # more than one wave of 112 trajectories per CTA on a few CTAs, ragged tail
    u0 = pl.pleiades_u0(N, f32=f32)
    ref = pl.pleiades_source(f32)
    dtype = pkg.F32 if f32 else pkg.F64
    kw = dict(reltol=1e-4, abstol=1e-5) if f32 else dict(reltol=1e-6, abstol=1e-8)
    for src, name in (pl.pleiades_pairs_source(f32), ref):
        prog = handle.compile(pkg.ALG_VERN7, dtype, 28, 0, src, name, extra_options=pkg._lib.OPT_SMEM_STAGES)
        try:
            assert prog.info["local_bytes_integrate"] <= 512        # only the cold plain-operator copies
            assert prog.info["smem_bytes_integrate"] > 200 * 1024
            n_here = N if name != ref[1] else 300
            for extra in ({}, {"saveat": [3.0], "save_start": True}, {"saveat": [3.0], "save_start": False, "save_end": False}):
                g = pkg.lowlevel.solve_host(prog, u0[:n_here], None, (0.0, 3.0), **dict(kw, **extra))
                o = oracle.solve(oracle.ALG_VERN7, ref, u0[:n_here], None, (0.0, 3.0), 28, 0, f32=f32, **dict(kw, **extra))
                assert_same_result(g, o)
                assert (g["nf"] == 2 + 10 * (g["naccept"] + g["nreject"])).all()
            # default tolerances, and failure retcodes leave the other lanes running
            g = pkg.lowlevel.solve_host(prog, u0[:200], None, (0.0, 3.0))
            o = oracle.solve(oracle.ALG_VERN7, ref, u0[:200], None, (0.0, 3.0), 28, 0, f32=f32)
            assert_same_result(g, o)
            g = pkg.lowlevel.solve_host(prog, u0[:33], None, (0.0, 3.0), maxiters=7, **kw)
            o = oracle.solve(oracle.ALG_VERN7, ref, u0[:33], None, (0.0, 3.0), 28, 0, f32=f32, maxiters=7, **kw)
            assert_same_result(g, o)
            assert (g["retcode"] == pkg._lib.RC_MAXITERS).all()
            # interior saveat rows need the lazy stages this variant does not store: rejected loudly
            with pytest.raises(pkg._lib.B200Error):
                pkg.lowlevel.solve_host(prog, u0[:8], None, (0.0, 3.0), saveat=[1.0, 2.0])
        finally:
            prog.close()


def test_smem_stage_kernel_small_system(pkg, handle, oracle):
    """The same variant on a small system (Lorenz, n = 3: 512 trajectories per CTA) equals the default Vern7 kernel's
    oracle, including a tstops-free run with user dt and dtmax."""
    pl = pkg.problems_library
    N = 3000
    p = pl.lorenz_params(N)
    src, name = pl.lorenz_source()
    prog = handle.compile(pkg.ALG_VERN7, pkg.F64, 3, 3, src, name, extra_options=pkg._lib.OPT_SMEM_STAGES)
    try:
        for kw in ({}, {"reltol": 1e-9, "abstol": 1e-9}, {"dt": 0.01, "dtmax": 0.05}):
            g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 10.0), **kw)
            o = oracle.solve(oracle.ALG_VERN7, (src, name), U0, p, (0.0, 10.0), 3, 3, **kw)
            assert_same_result(g, o)
    finally:
        prog.close()


@pytest.mark.parametrize("f32", [False, True])
def test_staged_saveat_queue_equals_in_step_interpolation(pkg, handle, oracle, f32):
    """B200ODE_OPT_STAGED_SAVEAT: saveat rows packed through the per-warp shared-memory queue (snapshot + row entries,
    drained 32 rows at a time by the whole warp) are the oracle's rows bit for bit — the headline grid, a fine grid with
    many rows per step (queue overflow inside one step), grid points that coincide with step ends, a truncated tail of
    failed trajectories, and both schedules."""
    pl = pkg.problems_library
    N = 4000
    p = pl.lorenz_params(N, f32=f32)
    src, name = pl.lorenz_source(f32)
    dtype = pkg.F32 if f32 else pkg.F64
    u0 = U0.astype(np.float32) if f32 else U0
    prog = handle.compile(pkg.ALG_TSIT5, dtype, 3, 3, src, name, extra_options=pkg._lib.OPT_STAGED_SAVEAT)
    try:
        assert prog.info["smem_bytes_integrate"] > 40 * 1024
        cases = [dict(saveat=[0.1 * k for k in range(1, 101)]),
                 dict(saveat=[0.005 * k for k in range(1, 401)], save_start=False),          # ~20 rows per step
                 dict(saveat=[0.37, 0.371, 0.372, 1.0, 2.0], save_end=False),
                 dict(saveat=[0.5 * k for k in range(1, 21)], maxiters=40),                  # failed trajectories: zero tail
                 dict(saveat=[0.25 * k for k in range(1, 41)], flags=pkg._lib.FLAG_STATIC_SCHEDULE),
                 dict()]
        for kw in cases:
            g = pkg.lowlevel.solve_host(prog, u0, p, (0.0, 2.0 if len(kw.get("saveat", [])) == 400 else 10.0), **kw)
            kwo = {k: v for k, v in kw.items() if k != "flags"}
            o = oracle.solve(oracle.ALG_TSIT5, (src, name), u0, p, (0.0, 2.0 if len(kw.get("saveat", [])) == 400 else 10.0), 3, 3,
                             f32=f32, **kwo)
            assert_same_result(g, o)
    finally:
        prog.close()


def _forced_sources(f32):
    """u' = -p0 u + (t > 1 ? p1 : 0) + (t > 2.5 ? p2 : 0), v' = u - v: a right-hand side with two switching times."""
    T = "float" if f32 else "double"
    s = "f" if f32 else ""
    rhs = ("void forced_rhs(%(T)s* du, const %(T)s* u, const %(T)s* p, const %(T)s t) {\n"
           "  du[0] = -p[0] * u[0] + (t > 1.0%(s)s ? p[1] : 0.0%(s)s) + (t > 2.5%(s)s ? p[2] : 0.0%(s)s);\n"
           "  du[1] = u[0] - u[1];\n}\n" % dict(T=T, s=s), "forced_rhs")
    jac = ("void forced_jac(%(T)s* J, const %(T)s* u, const %(T)s* p, const %(T)s t) {\n"
           "  J[0] = -p[0]; J[1] = 1.0%(s)s; J[2] = 0.0%(s)s; J[3] = -1.0%(s)s;\n}\n" % dict(T=T, s=s), "forced_jac")
    tg = ("void forced_tgrad(%(T)s* dT, const %(T)s* u, const %(T)s* p, const %(T)s t) { dT[0] = 0.0%(s)s; dT[1] = 0.0%(s)s; }\n"
          % dict(T=T, s=s), "forced_tgrad")
    return rhs, jac, tg


@pytest.mark.parametrize("f32", [False, True])
def test_d_discontinuities(pkg, handle, oracle, f32):
    """The d_discontinuities keyword (solve.jl:136; update_fsal! integrator_utils.jl:215-220; starting-time form
    solve.jl:887-901): the integrator stops on each declared time, moves one ulp past it and (FSAL steppers) evaluates
    its first stage again.  Bit-exact GPU vs oracle for Tsit5 (FSAL), Vern7 (not FSAL), Rosenbrock23 and Rodas5P, with
    saveat rows around the switching times, with the ragged per-step output, with a discontinuity at t0, combined with
    tstops, and with entries outside the span."""
    pl = pkg.problems_library
    N = 300
    idx = np.arange(N, dtype=np.uint64)
    rdt = np.float32 if f32 else np.float64
    p = np.stack([0.5 + pl.splitmix64_uniform(idx, 0), 1.0 + pl.splitmix64_uniform(idx, 1), -2.0 * pl.splitmix64_uniform(idx, 2)],
                 axis=1).astype(rdt)
    u0 = np.array([1.0, 0.0], dtype=rdt)
    rhs, jac, tg = _forced_sources(f32)
    dtype = pkg.F32 if f32 else pkg.F64
    tspan = (0.0, 4.0)
    grid = [0.5, 1.0, 1.25, 2.5, 2.75, 4.0]
    cases = [dict(d_discontinuities=[1.0, 2.5], saveat=grid),
             dict(d_discontinuities=[2.5, 1.0, -1.0, 7.0]),
             dict(d_discontinuities=[0.0, 1.0], saveat=grid, save_start=False),
             dict(d_discontinuities=[1.0], tstops=[2.5, 3.0], saveat=grid)]
    for alg, oalg, stiff in ((pkg.ALG_TSIT5, oracle.ALG_TSIT5, False), (pkg.ALG_VERN7, oracle.ALG_VERN7, False),
                             (pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23, True), (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P, True)):
        extra = dict(jac_src=jac[0], jac_name=jac[1], tgrad_src=tg[0], tgrad_name=tg[1]) if stiff else {}
        prog = handle.compile(alg, dtype, 2, 3, rhs[0], rhs[1], extra_options=pkg._lib.OPT_TSTOPS, **extra)
        okw = dict(jac=jac, tgrad=tg) if stiff else {}
        try:
            for kw in cases:
                g = pkg.lowlevel.solve_host(prog, u0, p, tspan, **kw)
                o = oracle.solve(oalg, rhs, u0, p, tspan, 2, 3, f32=f32, **dict(kw, **okw))
                assert_same_result(g, o)
                assert (g["retcode"] == 1).all()
            # the declared discontinuity changes the step sequence (otherwise the test would prove nothing)
            plain = pkg.lowlevel.solve_host(prog, u0, p, tspan, tstops=[1.0, 2.5])
            with_d = pkg.lowlevel.solve_host(prog, u0, p, tspan, d_discontinuities=[1.0, 2.5])
            if alg == pkg.ALG_TSIT5:
                assert (plain["nf"] != with_d["nf"]).any()
        finally:
            prog.close()
    # ragged per-step rows: the declared time is a row, the shifted time is not
    prog = handle.compile(pkg.ALG_TSIT5, dtype, 2, 3, rhs[0], rhs[1], extra_options=pkg._lib.OPT_TSTOPS + " " + pkg._lib.OPT_EVERYSTEP)
    try:
        g = pkg.lowlevel.solve_host_everystep(prog, u0, p, tspan, d_discontinuities=[1.0, 2.5])
        o = oracle.solve(oracle.ALG_TSIT5, rhs, u0, p, tspan, 2, 3, f32=f32, save_everystep=True, d_discontinuities=[1.0, 2.5])
        assert np.array_equal(g["row_offsets"], o["row_offsets"]) and np.array_equal(g["ts"], o["ts"])
        assert np.array_equal(bits(g["us"]), bits(o["us"]))
        assert (np.asarray(g["ts"], dtype=np.float64) == 1.0).sum() == N
    finally:
        prog.close()
    # a program without the tstops variant rejects the keyword
    prog = handle.compile(pkg.ALG_TSIT5, dtype, 2, 3, rhs[0], rhs[1])
    try:
        with pytest.raises(pkg._lib.B200Error):
            pkg.lowlevel.solve_host(prog, u0, p, tspan, d_discontinuities=[1.0])
    finally:
        prog.close()


@pytest.mark.parametrize("f32", [False, True])
def test_per_trajectory_tspans(pkg, handle, oracle, f32):
    """B200Problem.tspans (program option B200ODE_OPT_TSPANS): every trajectory integrates over its own (t0_i, tf_i) — what
    a prob_func that remakes tspan produces.  Final states and statistics through b200ode_solve, rows through the ragged
    output (with a saveat list cut per trajectory); Tsit5, Vern7, Rodas5P; bit-exact against the oracle, whose per-trajectory
    form equals one-trajectory solves (tests/test_oracle_properties.py)."""
    pl = pkg.problems_library
    N = 500
    rdt = np.float32 if f32 else np.float64
    idx = np.arange(N, dtype=np.uint64)
    t0 = -1.0 + 2.0 * pl.splitmix64_uniform(idx, 1)
    spans = np.stack([t0, t0 + 0.25 + 6.0 * pl.splitmix64_uniform(idx, 2)], axis=1)
    dtype = pkg.F32 if f32 else pkg.F64
    u0 = U0.astype(rdt)
    for alg, oalg, problem in ((pkg.ALG_TSIT5, oracle.ALG_TSIT5, "lorenz"), (pkg.ALG_VERN7, oracle.ALG_VERN7, "lorenz"),
                               (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P, "robertson")):
        if problem == "lorenz":
            src = pl.lorenz_source(f32); p = pl.lorenz_params(N, f32=f32); okw = {}
            csrc = (src[0], src[1])
            sp = spans
        else:
            src, j, tg = pl.robertson_sources(f32); p = pl.robertson_params(N, f32=f32); okw = dict(jac=j, tgrad=tg)
            csrc = (src[0], src[1], j[0], j[1], tg[0], tg[1])
            sp = np.stack([np.zeros(N), 10.0 ** (1.0 + 3.0 * pl.splitmix64_uniform(idx, 3))], axis=1)
        prog = handle.compile(alg, dtype, 3, 3, *csrc, extra_options=pkg._lib.OPT_TSPANS)
        try:
            for kw in ({}, {"reltol": 1e-5, "abstol": 1e-7} if not f32 else {"dt": 0.01}, {"dtmax": 0.3} if problem == "lorenz" else {"maxiters": 60}):
                g = pkg.lowlevel.solve_host(prog, u0, p, sp, **kw)
                o = oracle.solve(oalg, src, u0, p, sp, 3, 3, f32=f32, **dict(kw, **okw))
                assert_same_result(g, o)
                if "maxiters" not in kw:
                    assert np.array_equal(g["t_final"], sp[:, 1].astype(rdt).astype(np.float64))
            # a shared tspan is refused by such a program, and the rectangular rows are refused with spans
            with pytest.raises(pkg._lib.B200Error):
                pkg.lowlevel.solve_host(prog, u0, p, (0.0, 1.0))
        finally:
            prog.close()
    # ragged rows, with a saveat list that every trajectory cuts to its own span
    src = pl.lorenz_source(f32); p = pl.lorenz_params(N, f32=f32)
    prog = handle.compile(pkg.ALG_TSIT5, dtype, 3, 3, src[0], src[1], extra_options=pkg._lib.OPT_TSPANS + " " + pkg._lib.OPT_EVERYSTEP)
    try:
        for kw in ({}, {"saveat": [-0.5, 0.0, 0.5, 1.0, 2.0, 4.0]}, {"saveat": [0.25, 3.0], "save_start": False}):
            g = pkg.lowlevel.solve_host_everystep(prog, u0, p, spans, **kw)
            o = oracle.solve(oracle.ALG_TSIT5, src, u0, p, spans, 3, 3, f32=f32, save_everystep=True, **kw)
            assert np.array_equal(g["row_offsets"], o["row_offsets"]) and np.array_equal(g["ts"], o["ts"])
            assert np.array_equal(bits(g["us"]), bits(o["us"]))
            assert np.array_equal(bits(g["u_final"]), bits(o["u_final"]))
    finally:
        prog.close()
    prog = handle.compile(pkg.ALG_TSIT5, dtype, 3, 3, src[0], src[1])
    try:
        with pytest.raises(pkg._lib.B200Error):
            pkg.lowlevel.solve_host(prog, u0, p, spans)
    finally:
        prog.close()


# ---- AutoTsit5(Rosenbrock23()): per-trajectory switching between Tsit5 and Rosenbrock23 ------------------------------
def _vdp_mixed_params(pl, N, f32):
    """Van der Pol with mu spread over [0.5, 500]: the ensemble holds trajectories that never leave Tsit5, trajectories
    that switch to Rosenbrock23 for good and trajectories that switch back and forth."""
    mu = 0.5 * (1000.0 ** pl.splitmix64_uniform(np.arange(N, dtype=np.uint64), 0))
    return mu.reshape(N, 1).astype(np.float32 if f32 else np.float64)


@pytest.mark.parametrize("f32", [False, True])
def test_autotsit5_rosenbrock23_parity(pkg, handle, oracle, f32):
    pl = pkg.problems_library
    r, j, tg, n, np_, u0, _ = pl.stiff_sources("vdp", f32)
    N = 512
    p = _vdp_mixed_params(pl, N, f32)
    tspan = (0.0, 20.0)
    prog = handle.compile(pkg.ALG_AUTOTSIT5_ROSENBROCK23, pkg.F32 if f32 else pkg.F64, n, np_, r[0], r[1], j[0], j[1],
                          tg[0], tg[1])
    try:
        for kw in ({}, {"saveat": [0.5 * k for k in range(1, 41)]}, {"reltol": 1e-6, "abstol": 1e-8}):
            if f32 and "reltol" in kw:
                continue
            g = pkg.lowlevel.solve_host(prog, u0, p, tspan, **kw)
            o = oracle.solve(oracle.ALG_AUTOTSIT5_ROSENBROCK23, r, u0, p, tspan, n, np_, f32=f32, jac=j, tgrad=tg,
                             linsolve=1, **kw)
            assert_same_result(g, o)
            assert (g["retcode"] == 1).all()
            # the ensemble really is mixed: some trajectories never build a W, some do
            assert (g["njacs"] == 0).any() and (g["njacs"] > 0).any()
        # a non-stiff ensemble never switches: identical to plain Tsit5 except for nothing at all
        pn = np.full((64, 1), 0.5, dtype=p.dtype)
        g = pkg.lowlevel.solve_host(prog, u0, pn, (0.0, 5.0))
        t5 = oracle.solve(oracle.ALG_TSIT5, r, u0, pn, (0.0, 5.0), n, np_, f32=f32)
        assert (g["njacs"] == 0).all()
        assert np.array_equal(bits(g["u_final"]), bits(t5["u_final"])) and np.array_equal(g["nf"], t5["nf"])
    finally:
        prog.close()


def test_autotsit5_robertson_and_everystep(pkg, handle, oracle):
    pl = pkg.problems_library
    r, j, tg = pl.robertson_sources()
    p = pl.robertson_params(512)
    prog = handle.compile(pkg.ALG_AUTOTSIT5_ROSENBROCK23, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1])
    try:
        g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 1e5), reltol=1e-6, abstol=1e-8, saveat=[1.0, 100.0, 1e4])
        o = oracle.solve(oracle.ALG_AUTOTSIT5_ROSENBROCK23, r, U0, p, (0.0, 1e5), 3, 3, jac=j, tgrad=tg, reltol=1e-6,
                         abstol=1e-8, saveat=[1.0, 100.0, 1e4])
        assert_same_result(g, o)
        assert (g["retcode"] == 1).all() and (g["njacs"] > 0).all()
    finally:
        prog.close()
    prog = _everystep_prog(pkg, handle, pkg.ALG_AUTOTSIT5_ROSENBROCK23, False, "robertson")
    try:
        g = pkg.lowlevel.solve_host_everystep(prog, U0, p, (0.0, 100.0), reltol=1e-6, abstol=1e-8, saveat=[1.0, 50.0])
        o = oracle.solve(oracle.ALG_AUTOTSIT5_ROSENBROCK23, r, U0, p, (0.0, 100.0), 3, 3, jac=j, tgrad=tg, reltol=1e-6,
                         abstol=1e-8, saveat=[1.0, 50.0], save_everystep=True)
        _assert_same_ragged(g, o)
        with pytest.raises(pkg.B200Error):      # no dense output for the composite algorithm
            pkg.lowlevel.solve_host_dense(prog, U0, p, (0.0, 100.0), [1.0, 2.0])
    finally:
        prog.close()


# ---- callbacks (Tsit5): ContinuousCallback / DiscreteCallback against the oracle -------------------------------------------
def _ball_ensemble(pkg, N, f32):
    """Bouncing balls with their own gravity and restitution: p[i] = (g_i, e_i)."""
    U = pkg.problems_library.splitmix64_uniform
    idx = np.arange(N, dtype=np.uint64)
    p = np.stack([9.81 * (0.5 + U(idx, 0)), 0.8 + 0.2 * U(idx, 1)], axis=1)
    return p.astype(np.float32 if f32 else np.float64)


@pytest.mark.parametrize("f32", [False, True])
def test_callbacks_bouncing_ball_parity(pkg, handle, oracle, f32):
    from helpers import ball_sources, always_true_source, noop_affect_source
    rhs, cond, bounce, stop = ball_sources(f32)
    N = 777
    p = _ball_ensemble(pkg, N, f32)
    u0 = np.array([50.0, 0.0])
    dt = pkg.F32 if f32 else pkg.F64
    bounce_cb = dict(kind="continuous", condition=cond, affect=None, affect_neg=bounce)
    cases = [
        ([bounce_cb], {}, 0),
        ([dict(bounce_cb, interp_points=0, rootfind="right")], {"saveat": [1.0, 2.5, 7.0]}, 0),
        ([dict(bounce_cb, save_positions=(True, False))], {"saveat": [0.5 * k for k in range(1, 30)]}, pkg._lib.FLAG_NO_STEP_ROWS),
        ([dict(kind="continuous", condition=cond, affect=stop)], {}, 0),                     # terminate! at the first hit
        ([bounce_cb, dict(kind="discrete", condition=always_true_source(f32), affect=noop_affect_source(f32),
                          save_positions=(False, True))], {}, 0),
    ]
    for cbs, kw, flags in cases:
        prog = handle.compile(pkg.ALG_TSIT5, dt, 2, 2, rhs[0], rhs[1], extra_options=pkg._lib.OPT_EVERYSTEP, callbacks=cbs)
        try:
            g = pkg.lowlevel.solve_host_everystep(prog, u0, p, (0.0, 15.0), flags=flags, **kw)
            o = oracle.solve(oracle.ALG_TSIT5, rhs, u0, p, (0.0, 15.0), 2, 2, f32=f32, callbacks=cbs,
                             save_everystep=(flags == 0), ragged_saveat=(flags != 0), **kw)
            _assert_same_ragged(g, o)
            if cbs[0].get("affect") is stop:
                assert (g["retcode"] == pkg._lib.RC_TERMINATED).all() and (g["t_final"] < 15.0).all()
            else:
                assert (g["retcode"] == 1).all()
        finally:
            prog.close()
    # rectangular saveat output: save_positions = (false, false)
    cbs = [dict(bounce_cb, save_positions=(False, False))]
    prog = handle.compile(pkg.ALG_TSIT5, dt, 2, 2, rhs[0], rhs[1], callbacks=cbs)
    try:
        grid = [0.25 * k for k in range(1, 61)]
        g = pkg.lowlevel.solve_host(prog, u0, p, (0.0, 15.0), saveat=grid)
        o = oracle.solve(oracle.ALG_TSIT5, rhs, u0, p, (0.0, 15.0), 2, 2, f32=f32, callbacks=cbs, saveat=grid)
        assert_same_result(g, o)
        assert (g["us"][:, :, 0] > -1e-3).all()            # the balls stay above the floor
    finally:
        prog.close()
    with pytest.raises(pkg.B200Error):                      # save_positions need the ragged output
        handle.compile(pkg.ALG_TSIT5, dt, 2, 2, rhs[0], rhs[1], callbacks=[bounce_cb])


def test_callbacks_parameter_change_and_lorenz_events(pkg, handle, oracle):
    """A discrete callback that changes the trajectory's own parameter, and sign-change events of a chaotic state:
    Lorenz with a section at x = 0 (both directions) that flips nothing but saves, plus a one-shot rho kick."""
    pl = pkg.problems_library
    rhs = pl.lorenz_source(False)
    T = "double"
    cond = ("%s lz_sec(const %s* u, const %s* p, const %s t) { return u[0]; }\n" % (T, T, T, T), "lz_sec")
    mark = ("void lz_mark(%s* u, %s* p, const %s t, int* terminate) { u[2] = u[2] + 0.0; }\n" % (T, T, T), "lz_mark")
    dcond = ("%s lz_kick_c(const %s* u, const %s* p, const %s t) { return (t >= 5.0 && p[1] < 100.0) ? 1.0 : 0.0; }\n" % (T, T, T, T), "lz_kick_c")
    kick = ("void lz_kick(%s* u, %s* p, const %s t, int* terminate) { p[1] = p[1] + 100.0; }\n" % (T, T, T), "lz_kick")
    cbs = [dict(kind="continuous", condition=cond, affect=mark),
           dict(kind="discrete", condition=dcond, affect=kick, save_positions=(False, False))]
    N = 500
    p = pl.lorenz_params(N)
    prog = handle.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1], extra_options=pkg._lib.OPT_EVERYSTEP, callbacks=cbs)
    try:
        g = pkg.lowlevel.solve_host_everystep(prog, U0, p, (0.0, 10.0))
        o = oracle.solve(oracle.ALG_TSIT5, rhs, U0, p, (0.0, 10.0), 3, 3, callbacks=cbs, save_everystep=True)
        _assert_same_ragged(g, o)
        assert (g["retcode"] == 1).all()
    finally:
        prog.close()


def test_high_level_callbacks(pkg, oracle):
    """solve(EnsembleProblem, Tsit5(), EnsembleB200(); callback = CallbackSet(...)) — the reference's keyword, served on the
    device: default save_positions add the two event rows to sol.t; terminate! gives retcode Terminated."""
    from helpers import ball_sources
    rhs, cond, bounce, stop = ball_sources()
    CS = pkg.CSource
    N = 64
    p = _ball_ensemble(pkg, N, False)
    prob = pkg.ODEProblem(pkg.ODEFunction(CS(*rhs)), np.array([50.0, 0.0]), (0.0, 15.0), p[0])
    eprob = pkg.EnsembleProblem(prob, prob_func=pkg.TableProbFunc(p=p))
    cb = pkg.ContinuousCallback(CS(*cond), None, CS(*bounce))
    sim = pkg.solve(eprob, pkg.Tsit5(), pkg.EnsembleB200(), trajectories=N, callback=cb, saveat=1.0)
    spec = [cb.spec()]
    grid = [float(k) for k in range(1, 16)]
    o = oracle.solve(oracle.ALG_TSIT5, rhs, np.array([50.0, 0.0]), p, (0.0, 15.0), 2, 2, callbacks=spec, saveat=grid,
                     ragged_saveat=True)
    for i in (0, 7, N - 1):
        a, b = int(o["row_offsets"][i]), int(o["row_offsets"][i + 1])
        sol = sim[i]
        assert np.array_equal(sol.t, o["ts"][a:b]) and np.array_equal(bits(sol.u), bits(o["us"][a:b]))
        assert len(sol.t) > 16 and sol.retcode == "Success"          # 16 grid rows + two rows per bounce
    term = pkg.ContinuousCallback(CS(*cond), CS(*stop))
    sim = pkg.solve(eprob, pkg.Tsit5(), pkg.EnsembleB200(), trajectories=N, callback=pkg.CallbackSet(term))
    assert all(sim[i].retcode == "Terminated" for i in range(N))
    t_hit = np.sqrt(2 * 50.0 / p[:, 0])
    assert np.allclose([sim[i].t[-1] for i in range(N)], t_hit, rtol=1e-9)
    with pytest.raises(NotImplementedError):
        pkg.solve(eprob, pkg.Vern7(), pkg.EnsembleB200(), trajectories=N, callback=cb)


# ---- per-component tolerances (abstol / reltol vectors, solve.jl:377-399) ----------------------------------------------
@pytest.mark.parametrize("f32", [False, True])
def test_vector_tolerances_parity(pkg, handle, oracle, f32):
    pl = pkg.problems_library
    dt = pkg.F32 if f32 else pkg.F64
    N = 1500
    rt, at = [1e-5, 1e-3, 1e-4], [1e-7, 1e-4, 1e-6]
    # explicit (Tsit5, Vern7) and stiff (Rodas5P) steppers; a scalar abstol next to a vector reltol
    p = pl.lorenz_params(N, f32=f32)
    for alg, oalg in ((pkg.ALG_TSIT5, oracle.ALG_TSIT5), (pkg.ALG_VERN7, oracle.ALG_VERN7)):
        s, nm = pl.lorenz_source(f32)
        prog = handle.compile(alg, dt, 3, 3, s, nm, extra_options=pkg._lib.OPT_VECTOR_TOL)
        try:
            for kw in (dict(reltol=rt, abstol=at), dict(reltol=rt), dict(reltol=rt, abstol=at, saveat=[0.5, 1.0, 2.5])):
                g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 3.0), **kw)
                o = oracle.solve(oalg, (s, nm), U0, p, (0.0, 3.0), 3, 3, f32=f32, **kw)
                assert_same_result(g, o)
            # uniform vectors reproduce the scalar solve
            g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 3.0), reltol=[1e-4] * 3, abstol=[1e-6] * 3)
            o = oracle.solve(oalg, (s, nm), U0, p, (0.0, 3.0), 3, 3, f32=f32, reltol=1e-4, abstol=1e-6)
            assert_same_result(g, o)
        finally:
            prog.close()
        plain = handle.compile(alg, dt, 3, 3, s, nm)
        try:
            with pytest.raises(pkg.B200Error):
                pkg.lowlevel.solve_host(plain, U0, p, (0.0, 3.0), reltol=rt)
        finally:
            plain.close()
    r, j, tg = pl.robertson_sources(f32)
    pr = pl.robertson_params(512, f32=f32)
    prog = handle.compile(pkg.ALG_RODAS5P, dt, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1], extra_options=pkg._lib.OPT_VECTOR_TOL)
    try:
        kw = dict(reltol=[1e-4, 1e-3, 1e-4], abstol=[1e-6, 1e-9, 1e-6])
        g = pkg.lowlevel.solve_host(prog, U0, pr, (0.0, 1e3), **kw)
        o = oracle.solve(oracle.ALG_RODAS5P, r, U0, pr, (0.0, 1e3), 3, 3, f32=f32, jac=j, tgrad=tg, **kw)
        assert_same_result(g, o)
    finally:
        prog.close()


# ---- isoutofdomain (solve.jl:166; integrator_utils.jl:135-136,612-618) ----------------------------------------------------
def test_isoutofdomain_parity(pkg, handle, oracle):
    """Robertson concentrations must stay non-negative: isoutofdomain = any(u < 0) rejects such steps and retries them with
    dt * qmin.  Explicit (Tsit5 on Lorenz with an artificial half-space) and stiff (Rodas5P, Rosenbrock23) steppers."""
    pl = pkg.problems_library
    T = "double"
    pos = ("%s rober_out(const %s* u, const %s* p, const %s t) { return (u[0] < 0.0 || u[1] < 0.0 || u[2] < 0.0) ? 1.0 : 0.0; }\n" % (T, T, T, T), "rober_out")
    r, j, tg = pl.robertson_sources()
    pr = pl.robertson_params(700)
    cbs = [dict(kind="isoutofdomain", condition=pos)]
    for alg, oalg in ((pkg.ALG_RODAS5P, oracle.ALG_RODAS5P), (pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23)):
        prog = handle.compile(alg, pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1], callbacks=cbs)
        try:
            kw = dict(reltol=1e-2, abstol=1e-4, saveat=[1.0, 100.0])       # loose tolerances overshoot into u2 < 0
            g = pkg.lowlevel.solve_host(prog, U0, pr, (0.0, 1e4), **kw)
            o = oracle.solve(oalg, r, U0, pr, (0.0, 1e4), 3, 3, jac=j, tgrad=tg, callbacks=cbs, **kw)
            free = oracle.solve(oalg, r, U0, pr, (0.0, 1e4), 3, 3, jac=j, tgrad=tg, **kw)
            assert_same_result(g, o)
            assert (g["u_final"] >= 0).all() and (g["us"] >= 0).all()
            if alg == pkg.ALG_RODAS5P:       # (Rosenbrock23 never leaves the domain at these tolerances)
                assert (g["nreject"] != free["nreject"]).any()              # the domain check really rejected steps
        finally:
            prog.close()
    half = ("%s lz_out(const %s* u, const %s* p, const %s t) { return u[2] > 45.0 ? 1.0 : 0.0; }\n" % (T, T, T, T), "lz_out")
    s, nm = pl.lorenz_source()
    p = pl.lorenz_params(900)
    cbs = [dict(kind="isoutofdomain", condition=half)]
    prog = handle.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, s, nm, callbacks=cbs)
    try:
        g = pkg.lowlevel.solve_host(prog, U0, p, (0.0, 3.0), maxiters=2000)
        o = oracle.solve(oracle.ALG_TSIT5, (s, nm), U0, p, (0.0, 3.0), 3, 3, callbacks=cbs, maxiters=2000)
        assert_same_result(g, o)
        assert (g["retcode"] != 1).any() and (g["retcode"] == 1).any()      # trajectories that must leave the half-space stall
    finally:
        prog.close()


# ---- lane-group Rosenbrock23: rows of W across the lanes of a warp, LU and solves on shuffles ---------------------------------
@pytest.mark.parametrize("problem", ["vdp", "hires5", "hires8", "chain16"])
@pytest.mark.parametrize("f32", [False, True])
def test_lane_group_rosenbrock23_shuffle_lu(pkg, handle, oracle, problem, f32):
    """Same bits as the oracle's sequential partial-pivot LU path (and therefore as the one-thread kernel): final states,
    saveat rows, step counts, nf / njacs / nw / nsolve."""
    pl = pkg.problems_library
    rc, jc, tc, n, np_, u0, tspan = pl.stiff_component_sources(problem, f32)
    r, j, tg = pl.stiff_sources(problem, f32)[:3]
    N = 1000 if problem != "vdp" else 250
    p = pl.stiff_params(problem, N, f32=f32)
    tol = dict(reltol=1e-6, abstol=1e-8) if not f32 else dict(reltol=1e-3, abstol=1e-5)
    prog = handle.compile(pkg.ALG_ROSENBROCK23, pkg.F32 if f32 else pkg.F64, n, np_, rc[0], rc[1], jc[0], jc[1],
                          tc[0] if tc else None, tc[1] if tc else None, extra_options=pkg._lib.OPT_COMPONENT_RHS)
    try:
        assert prog.info["local_bytes_integrate"] <= 256
        mid = [tspan[1] * 0.01, tspan[1] * 0.5]
        for extra in ({}, {"saveat": mid}):
            kw = dict(tol, **extra)
            g = pkg.lowlevel.solve_host(prog, u0, p, tspan, **kw)
            o = oracle.solve(oracle.ALG_ROSENBROCK23, r, u0, p, tspan, n, np_, f32=f32, jac=j, tgrad=tg, **kw)
            assert_same_result(g, o)
            assert (g["retcode"] == 1).all()
    finally:
        prog.close()
    if problem == "hires8" and not f32:
        with pytest.raises(pkg.B200Error):          # n = 3 stays with the in-register inverse of the one-thread kernel
            rr, jj, tt = pl.robertson_sources()
            handle.compile(pkg.ALG_ROSENBROCK23, pkg.F64, 3, 3, rr[0], rr[1], jj[0], jj[1], tt[0], tt[1],
                           extra_options=pkg._lib.OPT_COMPONENT_RHS)


@pytest.mark.parametrize("name", ["Vern7", "DP5", "BS3", "Vern6", "Vern9", "Rosenbrock23", "Rodas5P", "AutoTsit5_Rosenbrock23"])
def test_discrete_callbacks_with_every_stepper(pkg, handle, oracle, name):
    """A one-shot kick (state jump + parameter change) once t >= 1, default save_positions, ragged rows: FSAL steppers
    re-evaluate their first stage in the next loopheader! (reset_fsal!), the others have nothing to refresh."""
    pl = pkg.problems_library
    T = "double"
    dcond = ("%s kick_c(const %s* u, const %s* p, const %s t) { return (t >= 1.0 && p[1] < 1e8) ? 1.0 : 0.0; }\n" % (T, T, T, T), "kick_c")
    kick = ("void kick(%s* u, %s* p, const %s t, int* terminate) { p[1] = p[1] + 1e9; u[0] = u[0] * 0.5; }\n" % (T, T, T), "kick")
    cbs = [dict(kind="discrete", condition=dcond, affect=kick)]
    alg = getattr(pkg, "ALG_" + name.upper())
    oalg = getattr(oracle, "ALG_" + name.upper())
    stiff = name in ("Rosenbrock23", "Rodas5P", "AutoTsit5_Rosenbrock23")
    N = 600
    if stiff:
        r, j, tg = pl.robertson_sources()
        p = pl.robertson_params(N)
        args = (r[0], r[1], j[0], j[1], tg[0], tg[1])
        okw = dict(jac=j, tgrad=tg)
        rhs, tspan, tol = r, (0.0, 50.0), dict(reltol=1e-6, abstol=1e-8)
    else:
        rhs = pl.lorenz_source()
        p = pl.lorenz_params(N)
        args = (rhs[0], rhs[1])
        okw, tspan, tol = {}, (0.0, 3.0), {}
    prog = handle.compile(alg, pkg.F64, 3, 3, *args, extra_options=pkg._lib.OPT_EVERYSTEP, callbacks=cbs)
    try:
        g = pkg.lowlevel.solve_host_everystep(prog, U0, p, tspan, saveat=[0.5, 2.0], **tol)
        o = oracle.solve(oalg, rhs, U0, p, tspan, 3, 3, callbacks=cbs, save_everystep=True, saveat=[0.5, 2.0], **okw, **tol)
        _assert_same_ragged(g, o)
        assert (g["retcode"] == 1).all()
    finally:
        prog.close()
    with pytest.raises(pkg.B200Error):          # continuous callbacks stay with Tsit5
        handle.compile(alg, pkg.F64, 3, 3, *args, extra_options=pkg._lib.OPT_EVERYSTEP,
                       callbacks=[dict(kind="continuous", condition=dcond, affect=kick)])


def _nonautonomous_sources(f32):
    """A non-autonomous 3-state kinetics-like system with analytic Jacobian and time gradient (the reverse-time tests need
    every user function to depend on t)."""
    T = "float" if f32 else "double"
    s = "f" if f32 else ""
    d = dict(T=T, s=s)
    rhs = ("void na_rhs(%(T)s* du, const %(T)s* u, const %(T)s* p, const %(T)s t) {\n"
           "  du[0] = -p[0]*u[0] + p[1]*u[1]*u[2] + t*u[2];\n"
           "  du[1] = p[0]*u[0] - p[1]*u[1]*u[2] - p[2]*u[1]*u[1] + t*t;\n"
           "  du[2] = p[2]*u[1]*u[1] - 0.5%(s)s*u[2]*t;\n}\n" % d, "na_rhs")
    jac = ("void na_jac(%(T)s* J, const %(T)s* u, const %(T)s* p, const %(T)s t) {\n"
           "  J[0] = -p[0]; J[1] = p[0]; J[2] = 0.0%(s)s;\n"
           "  J[3] = p[1]*u[2]; J[4] = -p[1]*u[2] - 2.0%(s)s*p[2]*u[1]; J[5] = 2.0%(s)s*p[2]*u[1];\n"
           "  J[6] = p[1]*u[1] + t; J[7] = -p[1]*u[1]; J[8] = -0.5%(s)s*t;\n}\n" % d, "na_jac")
    tg = ("void na_tgrad(%(T)s* dT, const %(T)s* u, const %(T)s* p, const %(T)s t) {\n"
          "  dT[0] = u[2]; dT[1] = 2.0%(s)s*t; dT[2] = -0.5%(s)s*u[2];\n}\n" % d, "na_tgrad")
    return rhs, jac, tg


@pytest.mark.parametrize("f32", [False, True])
def test_reverse_time(pkg, handle, oracle, f32):
    """tspan[2] < tspan[1] (tdir = -1): programs compiled with B200ODE_OPT_REVERSE_TIME integrate the mirrored problem
    du/ds = -f(u, p, -s); the oracle restates the reference's direction handling natively (solve.jl:273,401,1025-1037,
    1107-1120; integrator_utils.jl:268-324,343-348,1193-1207,1243-1256; initdt.jl) and replays its reverse-time known
    answers (tests/test_oracle_properties.py).  Bit-exact GPU vs oracle: final states, statistics, times; rectangular
    saveat rows (descending list), tstops + d_discontinuities, a user dt of either sign, dtmax, fixed steps, ragged
    rows, isoutofdomain; Tsit5, Vern7, DP5, Rosenbrock23, Rodas5P, AutoTsit5(Rosenbrock23())."""
    L = pkg._lib
    rdt = np.float32 if f32 else np.float64
    dtype = pkg.F32 if f32 else pkg.F64
    rng = np.random.default_rng(5)
    N = 400
    u0 = rng.uniform(0.1, 1.0, (N, 3)).astype(rdt)
    p = rng.uniform(0.5, 3.0, (N, 3)); p[:, 1] *= 30; p = p.astype(rdt)
    rhs, jac, tg = _nonautonomous_sources(f32)
    tspan = (2.0, 0.25)
    grid = [1.75, 1.5, 1.0, 0.3, 0.25]
    tol = dict(reltol=1e-4, abstol=1e-6) if f32 else dict(reltol=1e-6, abstol=1e-8)
    R = L.OPT_REVERSE_TIME
    algs = [(pkg.ALG_TSIT5, oracle.ALG_TSIT5, False), (pkg.ALG_VERN7, oracle.ALG_VERN7, False), (pkg.ALG_DP5, oracle.ALG_DP5, False),
            (pkg.ALG_ROSENBROCK23, oracle.ALG_ROSENBROCK23, True), (pkg.ALG_RODAS5P, oracle.ALG_RODAS5P, True),
            (pkg.ALG_AUTOTSIT5_ROSENBROCK23, oracle.ALG_AUTOTSIT5_ROSENBROCK23, True)]
    for alg, oalg, stiff in algs:
        extra = dict(jac_src=jac[0], jac_name=jac[1], tgrad_src=tg[0], tgrad_name=tg[1]) if stiff else {}
        okw = dict(jac=jac, tgrad=tg) if stiff else {}
        prog = handle.compile(alg, dtype, 3, 3, rhs[0], rhs[1], extra_options=R + " " + L.OPT_TSTOPS, **extra)
        try:
            for kw in (dict(), dict(saveat=grid), dict(saveat=grid, save_start=False, save_end=False),
                       dict(saveat=grid, tstops=[1.2, 0.7, 5.0], d_discontinuities=[1.0, 2.0]),
                       dict(dtmax=0.05), dict(dt=0.01, saveat=grid), dict(dt=-0.01)):
                g = pkg.lowlevel.solve_host(prog, u0, p, tspan, **dict(kw, **tol))
                o = oracle.solve(oalg, rhs, u0, p, tspan, 3, 3, f32=f32, **dict(kw, **tol, **okw))
                assert_same_result(g, o)
                assert (g["retcode"] == 1).all() and (g["t_final"] == tspan[1]).all()
            # forward spans are refused by a reverse-time program, reversed spans by an ordinary one
            with pytest.raises(L.B200Error):
                pkg.lowlevel.solve_host(prog, u0, p, (0.25, 2.0))
        finally:
            prog.close()
    plain = handle.compile(pkg.ALG_TSIT5, dtype, 3, 3, rhs[0], rhs[1])
    try:
        with pytest.raises(L.B200Error):
            pkg.lowlevel.solve_host(plain, u0, p, tspan)
    finally:
        plain.close()
    # fixed steps (adaptive = false): dt = -1/64 with a stop that is not a multiple of it
    prog = handle.compile(pkg.ALG_TSIT5, dtype, 3, 3, rhs[0], rhs[1], extra_options=" ".join((R, L.OPT_TSTOPS, L.OPT_FIXED_DT)))
    try:
        kw = dict(dt=-1.0 / 256, tstops=[1.01], saveat=grid)
        g = pkg.lowlevel.solve_host(prog, u0, p, tspan, **kw)
        o = oracle.solve(oracle.ALG_TSIT5, rhs, u0, p, tspan, 3, 3, f32=f32, adaptive=False, **kw)
        # (a fixed-step run of a stiff member may blow up: the sign bit of a NaN is the one thing negation does not mirror)
        fin = np.isfinite(o["u_final"]).all(axis=1)
        assert fin.sum() > N // 2 and np.array_equal(np.isfinite(g["u_final"]).all(axis=1), fin)
        assert_same_result({k: (v[fin] if isinstance(v, np.ndarray) and v.shape[:1] == (N,) else v) for k, v in g.items()},
                           {k: (v[fin] if isinstance(v, np.ndarray) and v.shape[:1] == (N,) else v) for k, v in o.items()})
    finally:
        prog.close()
    # ragged per-step rows: times come back as the caller's (descending), Tsit5 and Rodas5P
    for alg, oalg, stiff in (algs[0], algs[4]):
        extra = dict(jac_src=jac[0], jac_name=jac[1], tgrad_src=tg[0], tgrad_name=tg[1]) if stiff else {}
        okw = dict(jac=jac, tgrad=tg) if stiff else {}
        prog = handle.compile(alg, dtype, 3, 3, rhs[0], rhs[1], extra_options=R + " " + L.OPT_EVERYSTEP, **extra)
        try:
            g = pkg.lowlevel.solve_host_everystep(prog, u0, p, tspan, saveat=[1.5, 1.0], **tol)
            o = oracle.solve(oalg, rhs, u0, p, tspan, 3, 3, f32=f32, save_everystep=True, saveat=[1.5, 1.0], **tol, **okw)
            assert np.array_equal(g["row_offsets"], o["row_offsets"]) and np.array_equal(g["ts"], o["ts"])
            assert np.array_equal(bits(g["us"]), bits(o["us"]))
            assert np.all(np.diff(np.asarray(g["ts"][:o["row_offsets"][1]], dtype=np.float64)) < 0)
            # dense output: sol_i(tq) from the recomputed stages, queries in the order the integration meets them
            # (descending), extrapolation beyond both ends included
            tq = [2.1, 2.0, 1.9, 1.3, 1.0001, 0.7, 0.25, 0.2]
            gd = pkg.lowlevel.solve_host_dense(prog, u0, p, tspan, tq, **tol)
            od = oracle.solve(oalg, rhs, u0, p, tspan, 3, 3, f32=f32, dense_tq=tq, **tol, **okw)
            assert np.array_equal(bits(gd["dense"]), bits(od["dense"]))
            assert np.array_equal(bits(gd["dense"][:, 1]), bits(u0))          # Θ = 0 on the first interval
            with pytest.raises(L.B200Error):       # ascending queries are refused by a reverse-time program
                pkg.lowlevel.solve_host_dense(prog, u0, p, tspan, [0.5, 1.0])
        finally:
            prog.close()
    # isoutofdomain sees the caller's time: reject every step that ends in (0.9, 1.1) with u[0] above a threshold
    T = "float" if f32 else "double"
    dom = ("%s na_dom(const %s* u, const %s* p, const %s t) { return (t < 1.1 && t > 0.9 && u[0] > 0.05) ? 1 : 0; }\n" % (T, T, T, T), "na_dom")
    cbs = [dict(kind="isoutofdomain", condition=dom)]
    prog = handle.compile(pkg.ALG_TSIT5, dtype, 3, 3, rhs[0], rhs[1], extra_options=R, callbacks=cbs)
    try:
        g = pkg.lowlevel.solve_host(prog, u0, p, tspan, maxiters=3000, **tol)
        o = oracle.solve(oracle.ALG_TSIT5, rhs, u0, p, tspan, 3, 3, f32=f32, callbacks=cbs, maxiters=3000, **tol)
        assert_same_result(g, o)
        ref = oracle.solve(oracle.ALG_TSIT5, rhs, u0, p, tspan, 3, 3, f32=f32, **tol)
        assert (o["nreject"] != ref["nreject"]).any()          # the domain test did act
    finally:
        prog.close()
    # events in reverse time: a ball over a rising floor, every user function time dependent (the callbacks see the caller's t);
    # root finding left / right, save_positions rows, a discrete callback that changes u and p and terminates
    from helpers import moving_floor_sources
    mrhs, cond, bounce, disc, damp = moving_floor_sources(f32)
    pb = np.stack([9.81 * (0.5 + rng.uniform(size=N)), 0.8 + 0.2 * rng.uniform(size=N)], axis=1).astype(rdt)
    ub = np.array([50.0, 0.0], dtype=rdt)
    cbs = [dict(kind="continuous", condition=cond, affect=bounce, save_positions=(True, True)),
           dict(kind="discrete", condition=disc, affect=damp, save_positions=(False, True))]
    for cb, flags, kw in ((cbs, 0, dict()), (cbs, L.FLAG_NO_STEP_ROWS, dict(saveat=[14.0, 12.5, 9.0, 3.0, 1.0])),
                          ([dict(cbs[0], interp_points=0, rootfind="right")], 0, dict())):
        prog = handle.compile(pkg.ALG_TSIT5, dtype, 2, 2, mrhs[0], mrhs[1], extra_options=R + " " + L.OPT_EVERYSTEP, callbacks=cb)
        try:
            g = pkg.lowlevel.solve_host_everystep(prog, ub, pb, (15.0, 0.0), flags=flags, **kw)
            okw = dict(ragged_saveat=True) if flags else dict(save_everystep=True)
            o = oracle.solve(oracle.ALG_TSIT5, mrhs, ub, pb, (15.0, 0.0), 2, 2, f32=f32, callbacks=cb, **okw, **kw)
            assert np.array_equal(g["row_offsets"], o["row_offsets"]) and np.array_equal(g["ts"], o["ts"])
            assert np.array_equal(bits(g["us"]), bits(o["us"])) and np.array_equal(bits(g["u_final"]), bits(o["u_final"]))
            for k in ("naccept", "nreject", "nf", "retcode", "nsaved"):
                assert np.array_equal(g[k], o[k]), k
            assert (o["nsaved"] > (0 if flags else o["naccept"] + 1)).all()
        finally:
            prog.close()
    # combinations that are declined at compile time
    with pytest.raises(L.B200Error):
        handle.compile(pkg.ALG_TSIT5, dtype, 3, 3, rhs[0], rhs[1], extra_options=R + " " + L.OPT_TSPANS)


def test_high_level_solve_in_reverse_time(pkg, oracle):
    """solve(EnsembleProblem(ODEProblem(f, u0, (10.0, 0.0), p); prob_func), alg, EnsembleB200(); saveat = 0.1, ...): the grid is
    (t0 - h):-h:tf in the reference's range arithmetic, sol.t runs downwards; explicit and stiff stepper."""
    P = pkg
    pl = P.problems_library
    N = 128
    table = pl.lorenz_params(N)
    prob = P.ODEProblem(P.CSource(*pl.lorenz_source()), U0, (1.0, 0.0), table[0])
    ep = P.EnsembleProblem(prob, prob_func=P.TableProbFunc(p=table))
    grid = P.ranges.saveat_grid(0.1, (1.0, 0.0))
    assert grid == P.ranges.julia_range(0.9, -0.1, 0.0) and grid[-1] == 0.0 and len(grid) == 10
    o = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, (1.0, 0.0), 3, 3, saveat=grid)
    s = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, saveat=0.1)
    for i in (0, 5, N - 1):
        assert s[i].retcode == "Success" and list(s[i].t) == [1.0] + grid
        assert np.array_equal(bits(np.ascontiguousarray(s[i].u)), bits(o["us"][i]))
        assert s[i].stats.naccept == o["naccept"][i] and s[i].stats.nreject == o["nreject"][i]
    # default output (every step), a list saveat given in ascending order, tstops
    oe = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, (1.0, 0.0), 3, 3, save_everystep=True, tstops=[0.5])
    s = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, tstops=[0.5])
    a, b = oe["row_offsets"][3], oe["row_offsets"][4]
    assert np.array_equal(np.asarray(s[3].t), oe["ts"][a:b]) and 0.5 in list(s[3].t) and s[3].t[0] == 1.0 and s[3].t[-1] == 0.0
    assert np.array_equal(bits(np.ascontiguousarray(s[3].u)), bits(oe["us"][a:b]))
    s = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N, saveat=[0.0, 0.25, 0.5, 1.0])
    assert list(s[0].t) == [1.0, 0.5, 0.25, 0.0]
    # sol(t) in reverse time: any query order (sorted by tdir * t before the device pass)
    sd = P.solve(ep, P.Tsit5(), P.EnsembleB200(), trajectories=N)
    od = oracle.solve(oracle.ALG_TSIT5, pl.lorenz_source(), U0, table, (1.0, 0.0), 3, 3, dense_tq=[0.9, 0.5, 0.123])
    assert np.array_equal(bits(np.ascontiguousarray(sd[5]([0.5, 0.123, 0.9]))), bits(od["dense"][5][[1, 2, 0]]))
    assert np.array_equal(bits(np.ascontiguousarray(sd.at([0.9, 0.5, 0.123]))), bits(od["dense"]))
    # Robertson backwards over a short span with Rodas5P
    (r, rn), (j, jn), (tg, tgn) = pl.robertson_sources()
    pr = pl.robertson_params(N)
    u1 = np.array([0.9, 2e-5, 0.1])
    probr = P.ODEProblem(P.ODEFunction(P.CSource(r, rn), jac=P.CSource(j, jn), tgrad=P.CSource(tg, tgn)), u1, (1.0, 0.99), pr[0])
    epr = P.EnsembleProblem(probr, prob_func=P.TableProbFunc(p=pr))
    orr = oracle.solve(oracle.ALG_RODAS5P, (r, rn), u1, pr, (1.0, 0.99), 3, 3, jac=(j, jn), tgrad=(tg, tgn))
    s = P.solve(epr, P.Rodas5P(), P.EnsembleB200(), trajectories=N, save_everystep=False)
    assert (orr["retcode"] == 1).sum() > N // 2 and (orr["retcode"] != 1).any()     # backwards Robertson is unstable for some members
    for i in (0, 1, int(np.argmax(orr["retcode"] != 1)), N - 1):
        assert list(s[i].t) == [1.0, orr["t_final"][i]] and np.array_equal(bits(np.ascontiguousarray(s[i].u[-1])), bits(orr["u_final"][i]))
        assert s[i].stats.naccept == orr["naccept"][i] and s[i].stats.nreject == orr["nreject"][i]
        assert (s[i].retcode == "Success") == (orr["retcode"][i] == 1) and (orr["retcode"][i] != 1 or s[i].t[-1] == 0.99)


def test_reverse_time_meanvar_and_infinite_span(pkg, handle, oracle):
    """timeseries_steps_meanvar on the device for a reverse-time program (rows on a descending grid), and the C ABI's answer
    to an infinite span (declined: the kernels' stop tolerance is written for finite spans; the oracle follows the
    reference's inf_handling.jl)."""
    pl, ll, L = pkg.problems_library, pkg.lowlevel, pkg._lib
    N = 3000
    p = pl.lorenz_params(N)
    rhs = pl.lorenz_source(False)
    grid = pkg.ranges.saveat_grid(0.05, (0.5, 0.0))
    prog = handle.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1], extra_options=L.OPT_REVERSE_TIME)
    try:
        full = ll.solve_host(prog, U0, p, (0.5, 0.0), saveat=grid)
        st = ll.solve_host_meanvar(prog, U0, p, (0.5, 0.0), grid)
        us = full["us"].astype(np.float64)
        assert st["mean"].shape == (11, 3) and list(full["ts"]) == [0.5] + grid and grid[-1] == 0.0
        assert np.allclose(st["mean"], us.mean(axis=0), rtol=1e-12, atol=1e-13)
        assert np.allclose(st["var"], us.var(axis=0, ddof=1), rtol=1e-10, atol=1e-12)
        o = oracle.solve(oracle.ALG_TSIT5, rhs, U0, p, (0.5, 0.0), 3, 3, saveat=grid)
        assert np.array_equal(bits(full["us"]), bits(o["us"])) and np.array_equal(st["naccept"], o["naccept"])
        with pytest.raises(L.B200Error):
            ll.solve_host(prog, U0, p, (0.5, -np.inf))
    finally:
        prog.close()
    plain = handle.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
    try:
        with pytest.raises(L.B200Error):
            ll.solve_host(plain, U0, p, (0.0, np.inf))
    finally:
        plain.close()
