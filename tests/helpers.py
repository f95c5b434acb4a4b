"""Shared test helpers: small problems written as C source (the text both the oracle's gcc
build and the GPU path's NVRTC build compile)."""
import numpy as np


def linear_source(f32=False, name="lin_rhs"):
    """u' = 1.01 u  (ODEProblemLibrary.prob_ode_linear, u0 = 1/2, tspan (0,1))"""
    T = "float" if f32 else "double"
    lit = "1.01f" if f32 else "1.01"
    return ("void %s(%s* du, const %s* u, const %s* p, const %s t) { du[0] = %s * u[0]; }\n"
            % (name, T, T, T, T, lit)), name


def linear_jac_sources(f32=False):
    T = "float" if f32 else "double"
    lit = "1.01f" if f32 else "1.01"
    z = "0.0f" if f32 else "0.0"
    jac = ("void lin_jac(%s* J, const %s* u, const %s* p, const %s t) { J[0] = %s; }\n" % (T, T, T, T, lit)), "lin_jac"
    tg = ("void lin_tgrad(%s* dT, const %s* u, const %s* p, const %s t) { dT[0] = %s; }\n" % (T, T, T, T, z)), "lin_tgrad"
    return jac, tg


def linear2d_source(n=8, name="lin2d_rhs"):
    """prob_ode_2Dlinear flattened: u' = 1.01 u for a 4x2 matrix of states."""
    body = "".join("  du[%d] = 1.01 * u[%d];\n" % (i, i) for i in range(n))
    return ("void %s(double* du, const double* u, const double* p, const double t) {\n%s}\n" % (name, body)), name


def counting_source(name="cnt_rhs"):
    """RHS that counts its own calls in p-independent static storage (single-threaded oracle use only)."""
    return ("static long b200_test_calls = 0;\n"
            "long b200_test_get_calls(void) { return b200_test_calls; }\n"
            "void b200_test_reset_calls(void) { b200_test_calls = 0; }\n"
            "void %s(double* du, const double* u, const double* p, const double t) {\n"
            "  b200_test_calls++;\n  du[0] = p[0] * (u[1] - u[0]);\n  du[1] = u[0] * (p[1] - u[2]) - u[1];\n"
            "  du[2] = u[0] * u[1] - p[2] * u[2];\n}\n" % name), name


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def assert_same_result(g, o, keys=("naccept", "nreject", "nf", "retcode", "nsaved", "njacs", "nw", "nsolve")):
    """Bit-exact comparison of a GPU result dict with an oracle result dict."""
    for k in keys:
        assert np.array_equal(g[k], o[k]), "%s differs (first at %s)" % (k, np.nonzero(g[k] != o[k])[0][:5])
    assert np.array_equal(bits(g["u_final"]), bits(o["u_final"])), "u_final bits differ"
    assert np.array_equal(np.asarray(g["t_final"], dtype=np.float64), np.asarray(o["t_final"], dtype=np.float64))
    if o.get("us") is not None:
        assert np.array_equal(bits(g["us"]), bits(o["us"])), "saveat rows differ"
        assert np.array_equal(g["ts"], o["ts"]), "ts differ"


# ---- callbacks written as C source (condition: real f(u, p, t); affect: void f(u, p, t, int* terminate)) ----------
def ball_sources(f32=False):
    """Bouncing ball of test/Integrators_I/ode_event_tests.jl:105-126: y'' = -g (g = p[0]), event y = 0."""
    T = "float" if f32 else "double"
    rhs = ("void ball_rhs(%s* du, const %s* u, const %s* p, const %s t) { du[0] = u[1]; du[1] = -p[0]; }\n" % (T, T, T, T), "ball_rhs")
    cond = ("%s ball_cond(const %s* u, const %s* p, const %s t) { return u[0]; }\n" % (T, T, T, T), "ball_cond")
    bounce = ("void ball_bounce(%s* u, %s* p, const %s t, int* terminate) { u[1] = -p[1] * u[1]; }\n" % (T, T, T), "ball_bounce")
    stop = ("void ball_stop(%s* u, %s* p, const %s t, int* terminate) { *terminate = 1; }\n" % (T, T, T), "ball_stop")
    return rhs, cond, bounce, stop


def moving_floor_sources(f32=False, mirror=False):
    """A ball over a floor that rises with time, every user function time dependent: RHS (y' = v + 0.01 t), condition
    (y - 0.1 t), continuous affect (v -> -e v + 0.001 t), discrete condition (t < 7 && v > 3) and affect (halves v, changes
    the trajectory's own restitution, terminates for t < 0.5).  mirror = True gives the same problem in mirrored time s = -t
    (RHS negated, every function reads -s): what a reverse-time program integrates."""
    T = "float" if f32 else "double"
    f = "f" if f32 else ""
    tt = "(-t)" if mirror else "t"
    sg = "-" if mirror else ""
    d = dict(T=T, f=f, tt=tt, sg=sg)
    rhs = ("void mf_rhs(%(T)s* du, const %(T)s* u, const %(T)s* p, const %(T)s t) { du[0] = %(sg)s(u[1] + 0.01%(f)s*%(tt)s); du[1] = %(sg)s(-p[0]); }\n" % d, "mf_rhs")
    cond = ("%(T)s mf_cond(const %(T)s* u, const %(T)s* p, const %(T)s t) { return u[0] - 0.1%(f)s*%(tt)s; }\n" % d, "mf_cond")
    bounce = ("void mf_bounce(%(T)s* u, %(T)s* p, const %(T)s t, int* terminate) { u[1] = -p[1] * u[1] + 0.001%(f)s*%(tt)s; }\n" % d, "mf_bounce")
    disc = ("%(T)s mf_dc(const %(T)s* u, const %(T)s* p, const %(T)s t) { return (%(tt)s < 7.0%(f)s && u[1] > 3.0%(f)s) ? 1 : 0; }\n" % d, "mf_dc")
    damp = ("void mf_damp(%(T)s* u, %(T)s* p, const %(T)s t, int* terminate) { u[1] = 0.5%(f)s * u[1]; p[1] = 0.9%(f)s * p[1]; "
            "if (%(tt)s < 0.5%(f)s) *terminate = 1; }\n" % d, "mf_damp")
    return rhs, cond, bounce, disc, damp


def always_true_source(f32=False, name="cb_true"):
    T = "float" if f32 else "double"
    return ("%s %s(const %s* u, const %s* p, const %s t) { return 1; }\n" % (T, name, T, T, T), name)


def noop_affect_source(f32=False, name="cb_noop"):
    T = "float" if f32 else "double"
    return ("void %s(%s* u, %s* p, const %s t, int* terminate) { }\n" % (name, T, T, T), name)
