"""C code generation for user right-hand sides — the stand-in, in this Julia-less
environment, for `Symbolics.build_function(f_expr, u, p, t; target = CTarget())`.

The generated text has the shape Symbolics' C target emits:

    #include <math.h>
    void diffeqf(double* du, const double* RHS1, const double* RHS2, const double RHS3) {
      du[0] = ...;
    }

Rules that keep CPU-oracle and GPU results bit-identical:
  * integer powers are expanded to products (pow() is not correctly rounded and differs
    between glibc and CUDA), x**0.5 and sqrt(x) print as sqrt(),
  * no common-subexpression elimination or re-association beyond sympy's own
    canonical ordering — both sides compile the *same text* without contraction.
"""
import sympy as sp
from sympy.printing.c import C99CodePrinter


class _B200CPrinter(C99CodePrinter):
    def __init__(self, f32=False):
        settings = {}
        if f32:
            from sympy.codegen.ast import real, float32
            settings["type_aliases"] = {real: float32}
        super().__init__(settings)
        self._f32 = f32

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= abs(int(e)) <= 8:
            bs = self._print(b)
            if not (b.is_Symbol or b.is_Indexed):
                bs = "(" + bs + ")"
            prod = "*".join([bs] * abs(int(e)))
            if int(e) > 0:
                return "(" + prod + ")"
            one = "1.0f" if self._f32 else "1.0"
            return "(" + one + "/(" + prod + "))"
        if e == sp.Rational(3, 2):
            bs = self._print(b)
            fn = "sqrtf" if self._f32 else "sqrt"
            return "((%s)*%s(%s))" % (bs, fn, bs)
        if e == sp.Rational(-3, 2):
            bs = self._print(b)
            fn = "sqrtf" if self._f32 else "sqrt"
            one = "1.0f" if self._f32 else "1.0"
            return "(%s/((%s)*%s(%s)))" % (one, bs, fn, bs)
        return super()._print_Pow(expr)


def trace(f, n, np_, iip=False):
    """Call the user's Python function on symbols (the analogue of Symbolics tracing).

    f(u, p, t) -> sequence of n expressions (out-of-place), or f(du, u, p, t) in place.
    """
    u = [sp.Symbol("RHS1_%d" % i, real=True) for i in range(n)]
    p = [sp.Symbol("RHS2_%d" % i, real=True) for i in range(np_)]
    t = sp.Symbol("RHS3", real=True)
    if iip:
        du = [None] * n
        f(du, u, p, t)
        exprs = du
    else:
        exprs = list(f(u, p, t))
    if len(exprs) != n:
        raise ValueError("right-hand side returned %d components, expected %d" % (len(exprs), n))
    return [sp.sympify(e) for e in exprs], u, p, t


def _emit(name, out_name, exprs, u, p, t, f32):
    pr = _B200CPrinter(f32)
    ty = "float" if f32 else "double"
    sub = {s: sp.Symbol("RHS1[%d]" % i) for i, s in enumerate(u)}
    sub.update({s: sp.Symbol("RHS2[%d]" % i) for i, s in enumerate(p)})
    lines = ["#include <math.h>",
             "void %s(%s* %s, const %s* RHS1, const %s* RHS2, const %s RHS3) {" % (name, ty, out_name, ty, ty, ty)]
    for i, e in enumerate(exprs):
        lines.append("  %s[%d] = %s;" % (out_name, i, pr.doprint(sp.sympify(e).xreplace(sub))))
    lines.append("}")
    return "\n".join(lines) + "\n"


def build_function_c(f, n, np_, fname="diffeqf", f32=False, iip=False):
    """RHS source.  Returns (source, name)."""
    exprs, u, p, t = trace(f, n, np_, iip)
    return _emit(fname, "du", exprs, u, p, t, f32), fname


def build_jacobian_c(f, n, np_, fname="diffeqjac", f32=False, iip=False):
    """∂f/∂u, column-major n×n (Symbolics.jacobian + build_function)."""
    exprs, u, p, t = trace(f, n, np_, iip)
    J = [[sp.diff(exprs[i], u[j]) for j in range(n)] for i in range(n)]
    flat = [J[i][j] for j in range(n) for i in range(n)]   # column major
    return _emit(fname, "J", flat, u, p, t, f32), fname


def build_matrix_c(f, n, np_, fname="diffeqjac", f32=False, iip=False):
    """A user-supplied Jacobian callable jac(u, p, t) -> n x n (rows of rows, numpy array of expressions or sympy Matrix;
    in-place form jac(J, u, p, t) fills a list of rows), emitted column-major like build_jacobian_c."""
    u = [sp.Symbol("RHS1_%d" % i, real=True) for i in range(n)]
    p = [sp.Symbol("RHS2_%d" % i, real=True) for i in range(np_)]
    t = sp.Symbol("RHS3", real=True)
    if iip:
        M = [[0] * n for _ in range(n)]
        f(M, u, p, t)
    else:
        M = f(u, p, t)
    M = sp.Matrix(M)
    if M.shape != (n, n):
        raise ValueError("jac returned a %s matrix, expected %dx%d" % (M.shape, n, n))
    flat = [M[i, j] for j in range(n) for i in range(n)]   # column major
    return _emit(fname, "J", flat, u, p, t, f32), fname


def build_tgrad_c(f, n, np_, fname="diffeqtgrad", f32=False, iip=False):
    """∂f/∂t."""
    exprs, u, p, t = trace(f, n, np_, iip)
    return _emit(fname, "dT", [sp.diff(e, t) for e in exprs], u, p, t, f32), fname


# ---- component form (the lane-group kernels, B200ODE_OPT_COMPONENT_RHS) -------------------------------------------------
# The same expressions, printed by the same printer, behind an index: lane g of a trajectory's lane group evaluates only
# ITS component (row).  A generated function can only be a switch over the index; lanes of one group then take different
# cases (serialised), while the groups of a warp run the same case together.  Hand-written component functions with loops
# over a regular structure (problems_library.pleiades_component_source) avoid that.
def _emit_switch(name, index_args, cases, u, p, t, f32):
    pr = _B200CPrinter(f32)
    ty = "float" if f32 else "double"
    zero = "0.0f" if f32 else "0.0"
    sub = {s: sp.Symbol("RHS1[%d]" % i) for i, s in enumerate(u)}
    sub.update({s: sp.Symbol("RHS2[%d]" % i) for i, s in enumerate(p)})
    lines = ["#include <math.h>",
             "%s %s(%s, const %s* RHS1, const %s* RHS2, const %s RHS3) {" % (ty, name, index_args[0], ty, ty, ty),
             "  switch (%s) {" % index_args[1]]
    for key, e in cases:
        e = sp.sympify(e)
        if e == 0:
            continue                      # falls through to the default
        lines.append("    case %d: return %s;" % (key, pr.doprint(e.xreplace(sub))))
    lines += ["    default: return %s;" % zero, "  }", "}"]
    return "\n".join(lines) + "\n"


def build_function_component_c(f, n, np_, fname="diffeqf_i", f32=False, iip=False):
    """real NAME(int i, u, p, t) -> du_i."""
    exprs, u, p, t = trace(f, n, np_, iip)
    return _emit_switch(fname, ("int i", "i"), list(enumerate(exprs)), u, p, t, f32), fname


def build_jacobian_entry_c(f, n, np_, fname="diffeqjac_ij", f32=False, iip=False):
    """real NAME(int i, int j, u, p, t) -> (∂f/∂u)[i][j]; structural zeros share the default case."""
    exprs, u, p, t = trace(f, n, np_, iip)
    cases = [(i * n + j, sp.diff(exprs[i], u[j])) for i in range(n) for j in range(n)]
    return _emit_switch(fname, ("int i, int j", "i * %d + j" % n), cases, u, p, t, f32), fname


def build_tgrad_component_c(f, n, np_, fname="diffeqtgrad_i", f32=False, iip=False):
    """real NAME(int i, u, p, t) -> (∂f/∂t)_i, or None when the system is autonomous."""
    exprs, u, p, t = trace(f, n, np_, iip)
    cases = [(i, sp.diff(e, t)) for i, e in enumerate(exprs)]
    if all(sp.sympify(e) == 0 for _, e in cases):
        return None
    return _emit_switch(fname, ("int i", "i"), cases, u, p, t, f32), fname
