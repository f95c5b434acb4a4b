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


def build_tgrad_c(f, n, np_, fname="diffeqtgrad", f32=False, iip=False):
    """∂f/∂t."""
    exprs, u, p, t = trace(f, n, np_, iip)
    return _emit(fname, "dT", [sp.diff(e, t) for e in exprs], u, p, t, f32), fname
