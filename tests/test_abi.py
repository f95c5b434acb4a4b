"""The C-ABI library loads, exports every symbol include/b200ode.h declares, validates its
arguments, compiles through NVRTC without a GPU, and fails loudly when no device exists."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "b200ode.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ode_\w+)\s*\(", text)))


def test_header_symbols_exported(pkg):
    L = pkg._lib.lib()
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), "libb200ode.so does not export %s" % n
    assert sorted(pkg._lib.EXPORTS) == names, "ctypes binding and header disagree"


def test_version_string(pkg):
    assert b"b200ode" in pkg._lib.lib().b200ode_version()


def test_nvrtc_compiles_all_algorithms_without_gpu(pkg):
    pl = pkg.problems_library
    for f32 in (False, True):
        src, name = pl.lorenz_source(f32)
        for alg in (pkg.ALG_TSIT5, pkg.ALG_VERN7):
            cubin, log = pkg.compile_only(alg, pkg.F32 if f32 else pkg.F64, 3, 3, src, name)
            assert len(cubin) > 1000 and "b200_integrate" in log
        (r, j, tg) = pl.robertson_sources(f32)
        for alg in (pkg.ALG_ROSENBROCK23, pkg.ALG_RODAS5P):
            cubin, log = pkg.compile_only(alg, pkg.F32 if f32 else pkg.F64, 3, 3, r[0], r[1], j[0], j[1], tg[0], tg[1])
            assert len(cubin) > 1000


def test_headline_kernel_has_no_spills(pkg):
    src, name = pkg.problems_library.lorenz_source(False)
    _, log = pkg.compile_only(pkg.ALG_TSIT5, pkg.F64, 3, 3, src, name)
    m = re.search(r"Function properties for b200_integrate\s*\n\s*ptxas\s*\.\s*(\d+) bytes stack frame, (\d+) bytes spill stores", log)
    assert m, log
    assert int(m.group(2)) <= 16
    regs = int(re.search(r"b200_integrate.*?Used (\d+) registers", log, re.S).group(1))
    assert regs <= 128      # 4 CTAs x 128 threads per SM


def test_symbolics_style_source_with_include_line(pkg):
    src = "#include <math.h>\nvoid diffeqf(double* du, const double* RHS1, const double* RHS2, const double RHS3) {\n" \
          "  du[0] = RHS2[0] * RHS1[0] + sqrt(RHS1[0] * RHS1[0]);\n}\n"
    cubin, _ = pkg.compile_only(pkg.ALG_TSIT5, pkg.F64, 1, 1, src, "diffeqf")
    assert len(cubin) > 1000


def test_compile_errors_are_reported(pkg):
    with pytest.raises(pkg.B200Error) as e:
        pkg.compile_only(pkg.ALG_TSIT5, pkg.F64, 3, 3, "void f(double* du, const double* u, const double* p, const double t) { du[0] = nope; }", "f")
    assert e.value.code == pkg._lib.ECOMPILE and "nope" in str(e.value)
    with pytest.raises(pkg.B200Error) as e:
        pkg.compile_only(99, pkg.F64, 3, 3, "x", "f")
    assert e.value.code == pkg._lib.EINVAL
    with pytest.raises(pkg.B200Error) as e:
        pkg.compile_only(pkg.ALG_TSIT5, pkg.F64, 0, 3, "x", "f")
    assert e.value.code == pkg._lib.EINVAL
    with pytest.raises(pkg.B200Error) as e:     # Rosenbrock without a Jacobian
        src, name = pkg.problems_library.lorenz_source(False)
        pkg.compile_only(pkg.ALG_RODAS5P, pkg.F64, 3, 3, src, name)
    assert e.value.code == pkg._lib.EINVAL


def test_nslots_matches_reference_save_rules(pkg):
    ll = pkg.lowlevel
    # saveat=4.0 on (0,15): t == [0,4,8,12,15]  (test/InterfaceI/ode_saveat_tests.jl)
    assert ll.nslots_for((0.0, 15.0), [4.0, 8.0, 12.0]) == 5
    assert ll.nslots_for((0.0, 15.0), [4.0, 8.0, 12.0], save_end=False) == 4
    assert ll.nslots_for((0.0, 15.0), [4.0, 8.0, 12.0], save_start=False, save_end=False) == 3
    assert ll.nslots_for((0.0, 10.0), [5.0, 10.0]) == 3
    assert ll.nslots_for((0.0, 10.0), [5.0, 10.0], save_end=False) == 2     # skip_saveat_at_tspan_end
    assert ll.nslots_for((0.0, 10.0), None) == 0


def test_no_gpu_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B200Error) as e:
        pkg.Handle(0)
    assert e.value.code == pkg._lib.ECUDA and "no CPU fallback" in str(e.value)


def test_package_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "ordinarydiffeq.jl_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_nslots_agrees_with_rows_the_oracle_actually_saves(pkg):
    """b200ode_nslots (pure host arithmetic, no GPU needed) must equal the number of rows a successful trajectory
    saves, for every combination of the save flags and random grids — checked against the oracle's nsaved."""
    import itertools
    from oracle import oracle
    from helpers import linear_source
    ll = pkg.lowlevel
    rng = np.random.default_rng(11)
    s = linear_source()
    u0 = np.array([[0.5]])
    for trial in range(12):
        k = int(rng.integers(1, 6))
        grid = sorted(float(x) for x in rng.uniform(0.05, 1.0, k))
        if trial % 3 == 0:
            grid[-1] = 1.0                                   # grid ends at tf
        if trial % 4 == 1 and k > 1:
            grid[1] = grid[0]                                # duplicate point
        if trial % 6 == 3 and k > 1:
            grid[-2] = grid[-1] = 1.0                        # tf twice: save_end = false skips both copies
        for ss, se in itertools.product((None, True, False), repeat=2):
            n = ll.nslots_for((0.0, 1.0), grid, save_start=ss, save_end=se)
            o = oracle.solve(oracle.ALG_TSIT5, s, u0, None, (0.0, 1.0), 1, 0, saveat=grid, save_start=ss, save_end=se)
            assert o["retcode"][0] == 1
            assert n == o["nslots"] == o["nsaved"][0], (grid, ss, se, n, o["nsaved"][0])
            if n > 0:
                assert len(o["ts"]) == n and list(o["ts"]) == sorted(o["ts"])


def test_binding_structs_have_the_library_sizes(pkg):
    """The ctypes mirrors of the public structs (ordinarydiffeq.jl_b200/_lib.py) are as large as the C structs the library
    was built with (b200ode_struct_size) — a field added on one side only fails here, not as memory corruption."""
    L = pkg._lib.lib()
    import ctypes as C
    L.b200ode_struct_size.argtypes = [C.c_int]
    L.b200ode_struct_size.restype = C.c_int
    mirrors = [pkg._lib.B200Problem, pkg._lib.B200Opts, pkg._lib.B200Result, pkg._lib.B200DeviceProblem,
               pkg._lib.B200DeviceResult, pkg._lib.B200ProgramInfo, pkg._lib.B200CallbackSrc, pkg._lib.B200Ragged]
    for which, cls in enumerate(mirrors):
        assert L.b200ode_struct_size(which) == C.sizeof(cls), cls.__name__
    assert L.b200ode_struct_size(99) == -1


def test_reverse_time_program_variants_compile_without_gpu(pkg):
    """B200ODE_OPT_REVERSE_TIME: the wrapped user functions (RHS, Jacobian, time gradient, callbacks, isoutofdomain) compile for
    sm_100a in both precisions with the variants they combine with; the combinations the header rules out are refused by the
    compile entry point, not at launch."""
    L, pl = pkg._lib, pkg.problems_library
    R = L.OPT_REVERSE_TIME
    for f32 in (False, True):
        dt = pkg.F32 if f32 else pkg.F64
        src, name = pl.lorenz_source(f32)
        (r, rn), (j, jn), (tg, tgn) = pl.robertson_sources(f32)
        for extra in (R, R + " " + L.OPT_TSTOPS, R + " " + L.OPT_EVERYSTEP, R + " " + L.OPT_FIXED_DT, R + " " + L.OPT_VECTOR_TOL,
                      R + " " + L.opt_save_idxs([0, 2])):
            cubin, _ = pkg.compile_only(pkg.ALG_TSIT5, dt, 3, 3, src, name, extra_options=extra)
            assert len(cubin) > 1000
        for alg in (pkg.ALG_ROSENBROCK23, pkg.ALG_RODAS5P, pkg.ALG_AUTOTSIT5_ROSENBROCK23):
            cubin, _ = pkg.compile_only(alg, dt, 3, 3, r, rn, j, jn, tg, tgn, extra_options=R + " " + L.OPT_TSTOPS)
            assert len(cubin) > 1000
        cubin, _ = pkg.compile_only(pkg.ALG_ROSENBROCK23, dt, 3, 3, r, rn, j, jn, extra_options=R)       # no time gradient given
        assert len(cubin) > 1000
    from helpers import moving_floor_sources
    mrhs, cond, bounce, disc, damp = moving_floor_sources(False)
    cbs = [dict(kind="continuous", condition=cond, affect=bounce, save_positions=(True, True)),
           dict(kind="discrete", condition=disc, affect=damp, save_positions=(False, True)),
           dict(kind="isoutofdomain", condition=("double neg(const double* u, const double* p, const double t) { return u[0] < -1.0 && t < 1.0; }\n", "neg"))]
    cubin, _ = pkg.compile_only(pkg.ALG_TSIT5, pkg.F64, 2, 2, mrhs[0], mrhs[1], extra_options=R + " " + L.OPT_EVERYSTEP, callbacks=cbs)
    assert len(cubin) > 1000
    src, name = pl.lorenz_source(False)
    for bad in (R + " " + L.OPT_TSPANS, ):
        with pytest.raises(pkg.B200Error) as e:
            pkg.compile_only(pkg.ALG_TSIT5, pkg.F64, 3, 3, src, name, extra_options=bad)
        assert e.value.code == L.EUNSUPPORTED
    psrc, pname = pl.pleiades_pairs_source(False)
    with pytest.raises(pkg.B200Error) as e:
        pkg.compile_only(pkg.ALG_VERN7, pkg.F64, 28, 0, psrc, pname, extra_options=R + " " + L.OPT_SMEM_STAGES)
    assert e.value.code == L.EUNSUPPORTED
