"""CPU-side checks of the scalar device math (device/b200_base.cuh, b200_detmath.cuh), compiled
for the host by tests/detmath_host.cpp.  These are the routines that make oracle and kernel
agree bit-for-bit on dt0 and on the controller's fastpower."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host():
    out = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "detmath_host.so")
    src = os.path.join(HERE, "detmath_host.cpp")
    inc = os.path.join(ROOT, "ordinarydiffeq.jl_b200", "csrc", "device")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", inc, src, "-o", so], check=True)
    L = C.CDLL(so)
    for f in (L.t_log10_cr, L.t_exp10_cr, L.t_eps):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    L.t_fastpower.restype = C.c_double
    L.t_fastpower.argtypes = [C.c_double, C.c_double]
    L.t_fastpower_f.restype = C.c_float
    L.t_fastpower_f.argtypes = [C.c_float, C.c_float]
    L.t_div_const.restype = C.c_double
    L.t_div_const.argtypes = [C.c_double, C.c_double]
    L.t_div_const_f.restype = C.c_float
    L.t_div_const_f.argtypes = [C.c_float, C.c_float]
    return L


def test_device_log10_exp10_correctly_rounded(host):
    import mpmath as mp
    mp.mp.prec = 300
    rng = np.random.default_rng(2)
    for i in range(4000):
        x = float(10 ** rng.uniform(-15, 12)) if i % 2 else float(rng.uniform(0.5, 2.0))
        if i % 500 == 3:
            x = 10.0 ** int(rng.integers(-10, 10))
        assert host.t_log10_cr(x) == float(mp.log10(mp.mpf(x))), x
        y = float(rng.uniform(-20, 4))
        assert host.t_exp10_cr(y) == float(mp.power(10, mp.mpf(y))), y


def test_device_math_equals_oracle_math(host):
    """Two independent routes (double-double vs libquadmath) to the same bits."""
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for _ in range(20000):
        x = float(10 ** rng.uniform(-15, 12))
        assert host.t_log10_cr(x) == L.oracle_log10(x)
        y = float(rng.uniform(-20, 4))
        assert host.t_exp10_cr(y) == L.oracle_exp10(y)
        e = float(10 ** rng.uniform(-12, 4))
        b = float(rng.choice([7 / 50, 2 / 25, 1 / 10, 2 / 35, 7 / 20, 1 / 5]))
        assert host.t_fastpower(e, b) == L.oracle_fastpower(e, b)
        assert host.t_fastpower_f(np.float32(e), np.float32(b)) == L.oracle_fastpower_f32(np.float32(e), np.float32(b))
        assert host.t_eps(x) == L.oracle_eps(x)


def test_constant_divisor_division_is_exact(host):
    """b200_div_const(a, b) == a / b bit-for-bit (Markstein correction with RN(1/b))."""
    rng = np.random.default_rng(4)
    for b in [0.9, 3.0, 28.0, 1.0, 2.0, 5.0, 7.0, 64.0]:
        a = (rng.uniform(1, 2, 200000) * 2.0 ** rng.integers(-300, 300, 200000)) * rng.choice([-1.0, 1.0], 200000)
        for x in a[:20000]:
            assert host.t_div_const(float(x), b) == float(x) / b
        af = a[:20000].astype(np.float32)
        af = af[np.isfinite(af) & (af != 0)]
        for x in af:
            assert host.t_div_const_f(x, np.float32(b)) == np.float32(x) / np.float32(b)
    # variable divisor with its correctly rounded reciprocal (q11 / fastpower(errold), C ./ dt)
    for _ in range(40000):
        b = float(rng.uniform(1, 2) * 2.0 ** rng.integers(-60, 60))
        x = float(rng.uniform(1, 2) * 2.0 ** rng.integers(-200, 200))
        assert host.t_div_const(x, b) == x / b
        bf, xf = np.float32(b), np.float32(rng.uniform(1, 2) * 2.0 ** rng.integers(-40, 40))
        assert host.t_div_const_f(xf, bf) == xf / bf
    for x in (0.0, float("inf"), 5e-324, 1e-310, 1e305):
        assert host.t_div_const(x, 3.0) == x / 3.0
    assert np.isnan(host.t_div_const(float("nan"), 3.0))
