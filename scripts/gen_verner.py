"""Generate the Vern6 / Vern8 / Vern9 steppers (device + oracle flavours) from the reference's tableaus.

What is read from the reference (read-only): the Float64 coefficient literals of
lib/OrdinaryDiffEqVerner/src/verner_tableaus.jl and, from verner_rk_perform_step.jl, only the
*structure* of each stage (which coefficient multiplies which k, and which abscissa the stage is
evaluated at).  The structure is cross-checked against the coefficient names (aSSJJ / aSJ, rJJP).
No reference code is copied: the emitted C++ is written by this script.

A Vern7 stepper is generated as well (oracle flavour only, ALG id 102) so that
tests/test_oracle_properties.py can check the generator against the hand-written Vern7.

    python scripts/gen_verner.py [/root/reference]
"""
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAB = os.path.join(ref, "lib/OrdinaryDiffEqVerner/src/verner_tableaus.jl")
STEP = os.path.join(ref, "lib/OrdinaryDiffEqVerner/src/verner_rk_perform_step.jl")
tab_lines = open(TAB).read().split("\n")
step_text = open(STEP).read()


def fmt(v):
    v = v.strip()
    if re.fullmatch(r"[-+]?\d+", v):
        v += ".0"
    return v


def coeffs(func):
    """name -> literal of the CompiledFloats method of `func` (first method after the struct)."""
    start = None
    for i, ln in enumerate(tab_lines):
        if re.match(r"(@fold )?function %s\(" % func, ln):
            sig = " ".join(tab_lines[i:i + 8])
            if "CompiledFloats" in sig.split("convert")[0]:
                start = i
                break
    assert start is not None, func
    out = []
    for ln in tab_lines[start + 1:]:
        if ln.startswith("end"):
            break
        m = re.match(r"\s*(\w+)\s*=\s*convert\((T2?),\s*([-+0-9.eE]+)\)\s*$", ln)
        if m:
            out.append((m.group(1), fmt(m.group(3))))
    return out


def balanced(text, i):
    """text[i] == '(' -> index just past its matching ')'."""
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "(":
            depth += 1
        elif text[j] == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced")


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def terms(expr):
    return [(a, int(k)) for a, k in re.findall(r"(\w+) \* k(\d+)", expr)]


def parse_step(order):
    """Structure of perform_step!(…, ::Vern{order}ConstantCache): stages, solution row, error row."""
    m = re.search(r"@muladd function perform_step!\(integrator, cache::Vern%dConstantCache" % order, step_text)
    body = step_text[m.start():]
    body = body[:body.index("integrator.k[1] = k1")]
    body = re.sub(r"\s+", " ", body)
    gdefs = {}
    for gm in re.finditer(r"\bg(\d+) = uprev \+ dt \* \(", body):
        end = balanced(body, gm.end() - 1)
        gdefs[int(gm.group(1))] = body[gm.end() - 1:end]
    stages = {}
    fsal_first = "k1 = integrator.fsalfirst" in body
    for km in re.finditer(r"\bk(\d+) = f\(", body):
        end = balanced(body, km.end() - 1)
        args = split_args(body[km.end():end - 1])
        s = int(km.group(1))
        state, tm = args[0], args[2]
        if state == "uprev":
            stages[s] = dict(terms=[], time=None, first=False)
            continue
        if re.fullmatch(r"g\d+", state):
            state = "uprev + dt * " + gdefs[int(state[1:])]
        first = bool(re.fullmatch(r"uprev \+ a \* k1", state))
        tt = terms(state)
        if first:
            a21 = re.search(r"\ba = dt \* (\w+)", body).group(1)
            tt = [(a21, 1)]
        cm = re.fullmatch(r"t \+ (\w+) \* dt", tm)
        assert cm or tm == "t + dt", tm
        stages[s] = dict(terms=tt, time=cm.group(1) if cm else None, first=first)
    # FSAL (Vern6): the last stage is `integrator.fsallast = f(u, p, t + dt)`; kS = integrator.fsallast
    fsal = bool(re.search(r"integrator\.fsallast = f\(u, p, t \+ dt\)", body))
    um = re.search(r"\bu = uprev \+ dt \* \(", body)
    uexpr = body[um.end() - 1:balanced(body, um.end() - 1)]
    em = re.search(r"utilde = dt \* \(", body)
    eexpr = body[em.end() - 1:balanced(body, em.end() - 1)]
    nf = int(re.search(r"increment_nf!\(integrator\.stats, (\d+)\)", body).group(1))
    return dict(stages=stages, u=terms(uexpr), err=terms(eexpr), fsal=fsal, fsal_first=fsal_first, nf=nf)


def build(order):
    tabc = coeffs("Vern%dTableau" % order)
    extc = coeffs("Vern%dExtraStages" % order)
    intc = coeffs("Vern%dInterpolationCoefficients" % order)
    st = parse_step(order)
    names = dict(tabc)
    S = max(st["stages"]) if not st["fsal"] else max(st["stages"]) + 1     # FSAL: kS assigned from fsallast
    # cross-check the parsed structure against the coefficient names (aSSJJ or aSJ)
    def row_of(a):
        d = a[1:]
        # Vern6: aSJ, Vern7: aSSJ, Vern8/9: aSSJJ
        return (int(d[0]), int(d[1])) if len(d) == 2 else (int(d[:2]), int(d[2:]))
    for s, info in st["stages"].items():
        for a, k in info["terms"]:
            assert a in names and row_of(a) == (s, k), (order, s, a, k)
        want = sorted(n for n in names if re.fullmatch(r"a\d+", n) and row_of(n)[0] == s)
        assert want == sorted(a for a, _ in info["terms"]), (order, s, want, info["terms"])
        if info["time"]:
            assert info["time"] in names
    for a, k in st["u"]:
        assert a in names and (a == "b%d" % k or row_of(a) == (S, k)), (order, a, k)
    for a, k in st["err"]:
        assert a == "btilde%d" % k and a in names
    # extra stages: structure from the names
    enames = dict(extc)
    extra = {}
    for n in enames:
        if re.fullmatch(r"a\d{4}", n):
            s, k = int(n[1:3]), int(n[3:])
            extra.setdefault(s, []).append((n, k))
    for s in extra:
        extra[s].sort(key=lambda t: t[1])
        assert "c%d" % s in enames
    # interpolant: rJJP
    interp = {}
    for n, _ in intc:
        j, pw = int(n[1:3]), int(n[3:])
        interp.setdefault(j, []).append((pw, n))
    for j in interp:
        interp[j].sort()
        pws = [p for p, _ in interp[j]]
        assert pws == list(range(pws[0], pws[0] + len(pws))) and pws[0] == (1 if j == 1 else 2), (order, j, pws)
    nk = max(max(extra), max(interp))
    return dict(order=order, S=S, NK=nk, tab=tabc, ext=extc, interp_c=intc, st=st, extra=extra, interp=interp)


def emit(M, flavor):
    """C++ text of one stepper.  flavor: 'device' (struct B200Vern{o}) or 'oracle' (template struct Vern{o}[Gen])."""
    o, S, NK, st = M["order"], M["S"], M["NK"], M["st"]
    dev = flavor == "device"
    FMA = "b200_fma" if dev else "jl_fma"
    RT = "real" if dev else "R"
    N = "B200_N" if dev else "n"
    cname = "B200_VERN%d_C" % o
    C = (cname + ".") if dev else ""
    loop = ("#pragma unroll\n        for (int i = 0; i < B200_N; ++i) " if dev else "for (int i = 0; i < n; ++i) ")

    def rhs(dst, src, t):
        return ("B200_RHS(%s, %s, p, %s);" if dev else "P->f(%s, %s, p, %s);") % (dst, src, t)

    def chain(tt, kfmt="k[%d][i]"):
        e = "%s%s * %s" % (C, tt[0][0], kfmt % (tt[0][1] - 1))
        for a, k in tt[1:]:
            e = "%s(%s%s, %s, %s)" % (FMA, C, a, kfmt % (k - 1), e)
        return e

    def tm(c):
        return "%s(%s%s, dt, t)" % (FMA, C, c) if c else "t + dt"

    L = []
    allc = M["tab"] + M["ext"] + M["interp_c"]
    if dev:
        L.append("struct B200Vern%dCoeffs {" % o)
        L.append("    real " + ", ".join(n for n, _ in allc) + ";")
        L.append("};")
        L.append("__constant__ B200Vern%dCoeffs %s = {" % (o, cname))
        L.append("    " + ", ".join("(real)%s" % v for _, v in allc))
        L.append("};")
        L.append("struct B200Vern%d {" % o)
        L.append("    real k[%d][B200_N];     // k1..k%d, then the lazy stages k%d..k%d" % (NK, S, S + 1, NK))
        L.append("    static B200_D int order() { return %d; }" % o)
        L.append("    static B200_D real qsteady_min() { return (real)1; }")
        L.append("    static B200_D real qsteady_max() { return (real)1; }")
        if st["fsal_first"]:
            L.append("    B200_D void init(const real* u, const real* p, real t, int& nf) { B200_RHS(k[0], u, p, t); nf += 1; }")
        else:
            L.append("    B200_D void init(const real*, const real*, real, int&) {}")
        L.append("    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt, real reltol, real abstol, int& nf) {")
        L.append("        real tmp[B200_N];")
    else:
        sname = "Vern%d%s" % (o, "Gen" if o == 7 else "")
        L.append("template <typename R> struct %s {" % sname)
        L.append("    static constexpr int order = %d;" % o)
        L.append("    static constexpr bool is_rosenbrock = false;")
        L.append("    R k[%d][ORACLE_MAXN];" % NK)
        L.append("    const ProblemFns<R>* P;")
        L.append("    static bool fsal_init() { return %s; }" % ("true" if st["fsal_first"] else "false"))
        L.append("    static R qsteady_max() { return (R)1; }")
        if st["fsal_first"]:
            L.append("    void initialize(const R* uprev, const R* p, R t, Stats<R>& st) { P->f(k[0], uprev, p, t); st.nf += 1; }")
            L.append("    void update_fsal() { memcpy(k[0], k[%d], sizeof(R) * P->n); }" % (S - 1))
        else:
            L.append("    void initialize(const R*, const R*, R, Stats<R>&) {}")
            L.append("    void update_fsal() {}")
        L.append("    R perform_step(const R* uprev, R* u, const R* p, R t, R dt, const Opts<R>& o, Stats<R>& st, bool) {")
        L.append("        const int n = P->n;")
        L.append("        " + " ".join("const R %s = (R)%s;" % (n_, v) for n_, v in M["tab"]))
        L.append("        R tmp[ORACLE_MAXN];")
    for s in sorted(st["stages"]):
        info = st["stages"][s]
        if not info["terms"]:
            L.append("        " + rhs("k[%d]" % (s - 1), "uprev", "t"))
            continue
        if info["first"]:
            a = info["terms"][0][0]
            L.append("        { const %s a = dt * %s%s;" % (RT, C, a))
            L.append("        " + loop + "tmp[i] = %s(a, k[0][i], uprev[i]); }" % FMA)
        else:
            L.append("        " + loop + "tmp[i] = %s(dt, %s, uprev[i]);" % (FMA, chain(info["terms"])))
        L.append("        " + rhs("k[%d]" % (s - 1), "tmp", tm(info["time"])))
    L.append("        " + loop + "u[i] = %s(dt, %s, uprev[i]);" % (FMA, chain(st["u"])))
    if st["fsal"]:
        L.append("        " + rhs("k[%d]" % (S - 1), "u", "t + dt"))
    L.append("        %s += %d;" % ("nf" if dev else "st.nf", st["nf"]))
    if dev:
        L.append("        real ut[B200_N];")
        L.append("        " + loop + "ut[i] = dt * (%s);" % chain(st["err"]))
        L.append("        return b200_residual_norm(ut, uprev, u, reltol, abstol);")
        L.append("    }")
        if st["fsal_first"]:
            L.append("    B200_D void accept() {\n#pragma unroll\n        for (int i = 0; i < B200_N; ++i) k[0][i] = k[%d][i];\n    }" % (S - 1))
        else:
            L.append("    B200_D void accept() {}")
        L.append("    B200_D void dense_prepare(const real* uprev, const real* /*u*/, const real* p, real t, real dt) {")
        L.append("        real tmp[B200_N];")
    else:
        L.append("        R atmp[ORACLE_MAXN];")
        L.append("        for (int i = 0; i < n; ++i) {")
        L.append("            R utilde = dt * (%s);" % chain(st["err"]))
        L.append("            atmp[i] = residual(utilde, uprev[i], u[i], o.atol(i), o.rtol(i));")
        L.append("        }")
        L.append("        return rms(atmp, n);")
        L.append("    }")
        L.append("    void addsteps(const R* uprev, const R*, const R* p, R t, R dt) {")
        L.append("        const int n = P->n;")
        L.append("        " + " ".join("const R %s = (R)%s;" % (n_, v) for n_, v in M["ext"]))
        L.append("        R tmp[ORACLE_MAXN];")
    for s in sorted(M["extra"]):
        L.append("        " + loop + "tmp[i] = %s(dt, %s, uprev[i]);" % (FMA, chain(M["extra"][s])))
        L.append("        " + rhs("k[%d]" % (s - 1), "tmp", tm("c%d" % s)))
    L.append("    }")
    if dev:
        L.append("    B200_D void interp(real th, real dt, const real* y0, const real* /*y1*/, real* out) const {")
    else:
        L.append("    void interpolant(R th, R dt, const R* y0, const R*, R* out) const {")
        L.append("        const int n = P->n;")
        L.append("        " + " ".join("const R %s = (R)%s;" % (n_, v) for n_, v in M["interp_c"]))
    L.append("        const %s th2 = th * th;" % RT)
    js = sorted(M["interp"])
    for j in js:
        cs = [nm for _, nm in M["interp"][j]]
        e = C + cs[-1]
        for nm in reversed(cs[:-1]):
            e = "%s(th, %s, %s%s)" % (FMA, e, C, nm)
        L.append("        const %s b%d = %s * %s;" % (RT, j, "th" if j == 1 else "th2", e))
    e = "k[%d][i] * b%d" % (js[0] - 1, js[0])
    for j in js[1:]:
        e = "%s(k[%d][i], b%d, %s)" % (FMA, j - 1, j, e)
    L.append("        " + loop + "out[i] = %s(dt, %s, y0[i]);" % (FMA, e))
    L.append("    }")
    L.append("};")
    return "\n".join(L) + "\n"


HEAD = """// GENERATED by scripts/gen_verner.py — do not edit.
// Vern6 / Vern8 / Vern9 (lib/OrdinaryDiffEqVerner): coefficients are the Float64 literals of
// verner_tableaus.jl (Vern6 :32-69,254-316,430-491; Vern8 :1467-1576,1803-1934,2157-2268;
// Vern9 :2664-2985,3142-3317,3607-3850); the stage structure follows perform_step!
// (verner_rk_perform_step.jl:23-110, 583-765, 1024-1244), the lazy extra stages
// _ode_addsteps! (verner_addsteps.jl) and the interpolants interpolants.jl:25-55, 380-465, 669-770.
// Fusion: MuladdMacro nesting (first product plain, later ones fma'd outward), @evalpoly = Horner with fma.
"""

models = {o: build(o) for o in (6, 7, 8, 9)}
dev = HEAD + "#pragma once\n#include \"b200_base.cuh\"\n\n" + "\n".join(
    "#if B200_ALG == B200_ALG_VERN%d\n%s#endif\n" % (o, emit(models[o], "device")) for o in (6, 8, 9))
open(os.path.join(root, "ordinarydiffeq.jl_b200", "csrc", "device", "b200_verner_gen.cuh"), "w").write(dev)
orc = HEAD + "#define ORACLE_HAVE_VERNER_GEN 1\n\n" + "\n".join(emit(models[o], "oracle") for o in (6, 7, 8, 9))
open(os.path.join(root, "oracle", "oracle_verner_gen.inc"), "w").write(orc)
for o in (6, 7, 8, 9):
    M = models[o]
    print("Vern%d: S=%d NK=%d nf=%d fsal=%s coeffs=%d/%d/%d interp k=%s" % (
        o, M["S"], M["NK"], M["st"]["nf"], M["st"]["fsal"], len(M["tab"]), len(M["ext"]), len(M["interp_c"]), sorted(M["interp"])))
