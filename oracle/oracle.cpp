// oracle.cpp — CPU restatement of the reference's ensemble ODE path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under ordinarydiffeq.jl_b200/ may include, link
// or call this file; it is used by tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py as the checker and the CPU baseline.
//
// PARITY UNPINNED: the reference is Julia and no Julia runtime exists in this
// environment, and the reference's own tests hold no golden vectors for ensemble
// step counts.  What pins this file is (1) every
// known-answer / property test the reference's suite offers for this path
// (tests/test_oracle_properties.py lists them with file:line) and (2) a function-by-
// function correspondence with the reference source, cited below.  Arithmetic that
// lives in un-vendored Julia packages is restated from their published algorithms:
//   FastPower.jl 1.x      fastpower (Float32 pipeline)       -> fastpower()
//   MuladdMacro.jl 0.2.x  @muladd nesting                     -> explicit std::fma
//   StaticArrays.jl 1.9   sum(abs2,·) left fold, inv 3x3       -> rms(), inv3()
//   Julia Base            eps, nextfloat, log10, ^, exp2(Float32)
//
// Each trajectory is the out-of-place / SVector form of the reference
// (ConstantCache steppers), either time direction (tdir = sign(tf - t0) carried the way the reference carries it),
// adaptive PI control or fixed steps, tstops / d_discontinuities, callbacks (Tsit5), post-hoc dense evaluation.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fopenmp ... -lquadmath).

#include <quadmath.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <type_traits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAXN 64
#define ORACLE_MAXNP 256

// ---------------------------------------------------------------------------
// Julia Base scalar helpers
template <typename R> struct Bits;
template <> struct Bits<double> {
    typedef uint64_t U;
    static U to(double x) { U u; memcpy(&u, &x, 8); return u; }
    static double from(U u) { double x; memcpy(&x, &u, 8); return x; }
    static bool finite(double x) { return ((to(x) >> 52) & 0x7FF) != 0x7FF; }
};
template <> struct Bits<float> {
    typedef uint32_t U;
    static U to(float x) { U u; memcpy(&u, &x, 4); return u; }
    static float from(U u) { float x; memcpy(&x, &u, 4); return x; }
    static bool finite(float x) { return ((to(x) >> 23) & 0xFF) != 0xFF; }
};

// eps(x::AbstractFloat) (Base float.jl): ulp above |x|; eps(0)=nextfloat(0); NaN if non-finite
template <typename R> static R jl_eps(R x) {
    R ax = std::fabs(x);
    if (!Bits<R>::finite(ax)) return std::numeric_limits<R>::quiet_NaN();
    return Bits<R>::from(Bits<R>::to(ax) + 1) - ax;
}
template <typename R> static R jl_nextfloat(R x) { return Bits<R>::from(Bits<R>::to(x) + 1); }  // x >= 0 finite
// nextfloat(x) for a finite x of either sign (shift_past_discontinuity!, integrator_utils.jl:1188-1196, forward time)
template <typename R> static R jl_nextfloat_signed(R x) {
    if (x == (R)0) return Bits<R>::from(1);
    return x > (R)0 ? Bits<R>::from(Bits<R>::to(x) + 1) : Bits<R>::from(Bits<R>::to(x) - 1);
}
// prevfloat(x), finite x of either sign; _shift_past_discontinuity! picks by tdir (integrator_utils.jl:1193-1195)
template <typename R> static R jl_prevfloat_signed(R x) { return -jl_nextfloat_signed(-x); }
template <typename R> static R jl_shift_past(R x, R tdir) { return tdir > (R)0 ? jl_nextfloat_signed(x) : jl_prevfloat_signed(x); }
// Base.max/min propagate NaN
template <typename R> static R jl_max(R a, R b) { return std::isnan(a) ? a : (std::isnan(b) ? b : (a > b ? a : b)); }
template <typename R> static R jl_min(R a, R b) { return std::isnan(a) ? a : (std::isnan(b) ? b : (a < b ? a : b)); }
// Base.FastMath.max_fast(x,y) = ifelse(y > x, y, x)
template <typename R> static R jl_max_fast(R x, R y) { return y > x ? y : x; }
static inline double jl_fma(double a, double b, double c) { return std::fma(a, b, c); }
static inline float jl_fma(float a, float b, float c) { return std::fmaf(a, b, c); }

// log10 / 10^x: correctly rounded through binary128 (libquadmath); for Float32 the
// double result is rounded once more (Julia computes these in the working type to
// < 1 ulp; correctly rounded is the modal value — see DESIGN.md "initial dt").
static double cr_log10(double x) { return (double)log10q((__float128)x); }
static double cr_exp10(double x) { return (double)powq((__float128)10, (__float128)x); }

// ---------------------------------------------------------------------------
// FastPower.fastpower — Float32 pipeline (EXT FastPower.jl; call sites
// lib/OrdinaryDiffEqCore/src/integrators/controllers.jl:815-816)
static float fp_fastlog2(float x) {
    const float a = 0.338953f, b = 2.198599f, c = 1.523692f;
    uint32_t ux1i = Bits<float>::to(x);
    uint32_t exp = (ux1i & 0x7F800000u) >> 23;
    uint32_t greater = ux1i & 0x00400000u;
    float signif, fexp;
    if (greater != 0u) {
        uint32_t ux2i = (ux1i & 0x007FFFFFu) | 0x3f000000u;
        signif = Bits<float>::from(ux2i);
        fexp = (float)exp - 126.0f;
    } else {
        uint32_t ux2i = (ux1i & 0x007FFFFFu) | 0x3f800000u;
        signif = Bits<float>::from(ux2i);
        fexp = (float)exp - 127.0f;
    }
    signif = signif - 1.0f;
    float t = a * signif;
    t = t + b;
    t = signif * t;
    float d = signif + c;
    t = t / d;
    return fexp + t;
}
// Base.Math.exp2_fast(::Float32) = exp_impl_fast(x, Val(2)) (base/special/exp.jl)
static float jl_exp2_fast_f32(float x) {
    if (x >= 128.0f) return std::numeric_limits<float>::infinity();
    if (x <= -150.0f) return 0.0f;
    float N_float = std::nearbyintf(x);      // round(x), RoundNearest (ties to even)
    int32_t N = (int32_t)N_float;
    float r = std::fmaf(N_float, -1.0f, x);  // LogBU(Val(2),Float32) = -1
    r = std::fmaf(N_float, 0.0f, r);         // LogBL(Val(2),Float32) = 0
    static const float c[8] = {1.0f, 0.6931472f, 0.2402265f, 0.05550411f, 0.009618025f,
                               0.0013333423f, 0.00015469732f, 1.5316464e-5f};
    float small_part = c[7];
    for (int i = 6; i >= 0; --i) small_part = std::fmaf(r, small_part, c[i]);   // evalpoly = Horner with muladd
    float twopk = Bits<float>::from((uint32_t)(N + 127) << 23);
    return twopk * small_part;
}
static double fastpower(double x, double y) {
    if (x == 0.0) return 0.0;
    if (std::isinf(x) && std::isinf(y)) return std::numeric_limits<double>::infinity();
    return (double)jl_exp2_fast_f32((float)y * fp_fastlog2((float)x));
}
static float fastpower(float x, float y) {
    if (x == 0.0f) return 0.0f;
    if (std::isinf(x) && std::isinf(y)) return std::numeric_limits<float>::infinity();
    return jl_exp2_fast_f32(y * fp_fastlog2(x));
}

// ---------------------------------------------------------------------------
template <typename R> struct Fn {
    typedef void (*rhs_t)(R* du, const R* u, const R* p, const R t);
};

enum { ALG_TSIT5 = 1, ALG_VERN7 = 2, ALG_ROS23 = 3, ALG_RODAS5P = 4, ALG_DP5 = 5, ALG_BS3 = 6,
       ALG_RODAS5 = 7, ALG_RODAS4 = 8, ALG_RODAS42 = 9, ALG_RODAS4P = 10, ALG_RODAS4P2 = 11,
       ALG_VERN6 = 12, ALG_VERN8 = 13, ALG_VERN9 = 14, ALG_ROS32 = 15, ALG_RODAS5PE = 16, ALG_AUTOTSIT5_ROS23 = 17, ALG_RODAS3P = 18, ALG_RODAS23W = 19,
       ALG_VERN7_GENERATED = 102 };
enum { RC_DEFAULT = 0, RC_SUCCESS = 1, RC_MAXITERS = 2, RC_DTLESSTHANMIN = 3, RC_UNSTABLE = 4, RC_DTNAN = 5 };

// One callback of the CallbackSet.  condition: R f(const R* u, const R* p, R t); affect: void f(R* u, R* p, R t, int* terminate)
// (the C rendering of condition(u, t, integrator) / affect!(integrator); *terminate = 1 is terminate!(integrator)).
struct OracleCallback {
    int kind;                 // 0 DiscreteCallback, 1 ContinuousCallback, 2 the isoutofdomain function (condition only)
    void* condition;
    void* affect;             // NULL: nothing
    void* affect_neg;         // continuous only; NULL: nothing
    int rootfind;             // 0 NoRootFind, 1 LeftRootFind (default), 2 RightRootFind
    int interp_points;        // default 10
    double abstol;            // default 10eps(Float64)
    double repeat_nudge;      // default 1//100
    int save_before, save_after;   // save_positions
};
enum { RC_TERMINATED = 6 };

// Element i (1-based) of range(start, stop = stop, length = len) for IEEE floats — Base._linspace + the twice-precision
// getindex (Julia Base twiceprecision.jl, EXT).  The rational shortcut Base.range takes for "simple" end points is not
// reproduced; the sample points only bracket sign changes, so a last-bit difference cannot move an event unless its
// root lies within an ulp of a sample point.
template <typename R> struct JlLinspace {
    R ref_hi, ref_lo, step_hi, step_lo; long long offset, len;
    static void add12(R x, R y, R& hi, R& lo) {
        if (std::fabs(y) > std::fabs(x)) std::swap(x, y);
        hi = x + y; lo = (x - hi) + y;
    }
    static R truncbits(R x, int nb) {
        typename Bits<R>::U m = ~(typename Bits<R>::U)0;
        return Bits<R>::from(Bits<R>::to(x) & (m << nb));
    }
    JlLinspace(R start, R stop, long long len_) : len(len_) {
        R delta = stop - start;                                   // (finite end points, no overflow handling)
        R tmin = -(start / delta);
        // imin = round(Int, tmin*(len-1)+1) (ties to even); clamp before the conversion
        R timin = std::nearbyint(tmin * (R)(len - 1) + (R)1);
        long long imin = timin <= (R)1 ? 1 : (timin >= (R)len ? len : (long long)timin);
        R ref, step;
        if (1 < imin && imin < len) {
            double t = (double)(imin - 1) / (double)(len - 1);
            ref = (R)((1 - t) * (double)start + t * (double)stop);
            step = (imin - 1 < len - imin) ? (ref - start) / (R)(imin - 1) : (stop - ref) / (R)(len - imin);
        } else if (imin <= 1) { imin = 1; ref = start; step = delta / (R)(len - 1); }
        else { imin = len; ref = stop; step = delta / (R)(len - 1); }
        const R m = std::nextafter(std::numeric_limits<R>::max(), (R)0);
        const R k = (R)std::max(imin - 1, len - imin);
        const R lo = std::max(-(m + ref) / k, (-m + ref) / k), hi = std::min((m - ref) / k, (m + ref) / k);
        R step_pre = step < lo ? lo : (step > hi ? hi : step);
        const int prec_half = std::is_same<R, float>::value ? 12 : 27;         // cld(precision(T), 2)
        long long mx = std::max(imin - 1, len - imin);
        int nbl = len < 2 ? 0 : (int)std::ceil(std::log2((double)mx)) + 1;
        const int nb = std::min(prec_half, nbl);
        step_hi = truncbits(step_pre, nb);
        R x1h, x1l, x2h, x2l;
        add12((R)(1 - imin) * step_hi, ref, x1h, x1l);
        add12((R)(len - imin) * step_hi, ref, x2h, x2l);
        R a = (start - x1h) - x1l, b = (stop - x2h) - x2l;
        step_lo = (b - a) / (R)(len - 1);
        ref_hi = ref; ref_lo = a - (R)(1 - imin) * step_lo;
        offset = imin;
    }
    R operator[](long long i) const {
        const R u = (R)(i - offset);
        const R sh = u * step_hi, sl = u * step_lo;
        R xh, xl; add12(ref_hi, sh, xh, xl);
        return xh + (xl + (sl + ref_lo));
    }
};

template <typename R> struct Opts {
    R reltol, abstol, dt, dtmin, dtmax;
    // per-component tolerances (abstol / reltol given as vectors, solve.jl:377-399); NULL: the scalars above
    const R* abstol_v = nullptr; const R* reltol_v = nullptr;
    R atol(int i) const { return abstol_v ? abstol_v[i] : abstol; }
    R rtol(int i) const { return reltol_v ? reltol_v[i] : reltol; }
    long long maxiters;
    const R* saveat; int nsaveat;
    bool save_start, save_end, save_end_user;
    int linsolve;      // 0: StaticWOperator inverse (n<=3), 1: partial-pivot LU
    bool save_everystep = false;   // solve.jl:138 (default isempty(saveat)); ragged rows, see Out::row_offsets
    // opts.tstops as initialize_tstops builds it (solve.jl:1021-1040): ascending, inside (t0, tf), tf last; NULL: {tf}
    const R* tstops = nullptr; int ntstops = 0;
    // opts.d_discontinuities as reinit_d_discontinuities! builds it (solve.jl:1185-1197): entries >= t0, ascending
    const R* disc = nullptr; int ndisc = 0;
    bool adaptive = true;          // false: fixed dt = opts.dt (dtcache), every step accepted
    // callbacks (CallbackSet: continuous callbacks first, then discrete ones, each group in the order given)
    const struct OracleCallback* cbs = nullptr; int ncb = 0;
    // isoutofdomain(u, p, t) (solve.jl:166): R f(const R* u, const R* p, R t), non-zero = outside; NULL: never
    void* isout = nullptr;
    // Test switch (tests/test_oracle_properties.py): this FORWARD run stands for the mirror image of a reverse-time run, the
    // way the CUDA path's B200_REVERSE programs integrate it — the two places where the reference is not symmetric in tdir
    // (fix_dt_at_bounds!'s dtmin clamp, check_error's tstop comparison) then act as they do for tdir < 0.  Lets the CPU suite
    // check "mirrored forward + these two switches == native reverse" where they matter (dtmin > 0).
    bool mirror_of_reverse = false;
};

// ODE_DEFAULT_NORM(u::StaticArray, t) = sqrt_fast(real(sum(abs2,u)) / max(length(u),1))
// (lib/DiffEqBase/src/common_defaults.jl:102-107); StaticArrays mapreduce is a left fold.
template <typename R> static R rms(const R* v, int n) {
    R acc = v[0] * v[0];
    for (int i = 1; i < n; ++i) acc = acc + v[i] * v[i];
    return std::sqrt(acc / (R)(n > 1 ? n : 1));
}

// calculate_residuals(ũ::Number, u₀, u₁, α, ρ, internalnorm, t)
// (lib/DiffEqBase/src/calculate_residuals.jl:9-14): @muladd @fastmath ũ/(α+max(|u₀|,|u₁|)*ρ)
template <typename R> static R residual(R ut, R u0, R u1, R abstol, R reltol) {
    return ut / jl_fma(jl_max_fast(std::fabs(u0), std::fabs(u1)), reltol, abstol);
}

// ---------------------------------------------------------------------------
// Steppers.  Each provides: order, fsal, qsteady_max, init, attempt, accept, dense, interp.
template <typename R> struct Stats { int nf = 0, njacs = 0, nw = 0, nsolve = 0; };

template <typename R> struct ProblemFns {
    typename Fn<R>::rhs_t f = nullptr, jac = nullptr, tgrad = nullptr;
    int n = 0, np = 0;
};

// Tsit5 — lib/OrdinaryDiffEqTsit5/src/tsit_perform_step.jl:125-186 (ConstantCache),
// tableau tsit_tableaus.jl:52-92, interpolant interpolants.jl:32-57 + tsit_tableaus.jl:244-276
template <typename R> struct Tsit5 {
    static constexpr int order = 5;
    static constexpr bool is_rosenbrock = false;
    R k[7][ORACLE_MAXN];
    R g6[ORACLE_MAXN];                // stage state of k6 (read by the composite algorithm's stiffness estimate)
    const ProblemFns<R>* P;

    void initialize(const R* uprev, const R* p, R t, Stats<R>& st) {
        P->f(k[0], uprev, p, t);      // fsalfirst = f(uprev, p, t)
        st.nf += 1;
    }
    R perform_step(const R* uprev, R* u, const R* p, R t, R dt, const Opts<R>& o, Stats<R>& st, bool) {
        const int n = P->n;
        const R c1 = (R)0.161, c2 = (R)0.327, c3 = (R)0.9, c4 = (R)0.9800255409045097;
        const R a21 = (R)0.161, a31 = (R)-0.008480655492356989, a32 = (R)0.335480655492357,
                a41 = (R)2.8971530571054935, a42 = (R)-6.359448489975075, a43 = (R)4.3622954328695815,
                a51 = (R)5.325864828439257, a52 = (R)-11.748883564062828, a53 = (R)7.4955393428898365,
                a54 = (R)-0.09249506636175525, a61 = (R)5.86145544294642, a62 = (R)-12.92096931784711,
                a63 = (R)8.159367898576159, a64 = (R)-0.071584973281401, a65 = (R)-0.028269050394068383,
                a71 = (R)0.09646076681806523, a72 = (R)0.01, a73 = (R)0.4798896504144996,
                a74 = (R)1.379008574103742, a75 = (R)-3.290069515436081, a76 = (R)2.324710524099774;
        const R btilde1 = (R)-0.00178001105222577714, btilde2 = (R)-0.0008164344596567469,
                btilde3 = (R)0.007880878010261995, btilde4 = (R)-0.1447110071732629,
                btilde5 = (R)0.5823571654525552, btilde6 = (R)-0.45808210592918697,
                btilde7 = (R)0.015151515151515152;
        R *k1 = k[0], *k2 = k[1], *k3 = k[2], *k4 = k[3], *k5 = k[4], *k6 = k[5], *k7 = k[6];
        R tmp[ORACLE_MAXN];
        R a = dt * a21;
        // k2 = f(uprev + a*k1, p, t + c1*dt)
        for (int i = 0; i < n; ++i) tmp[i] = jl_fma(a, k1[i], uprev[i]);
        P->f(k2, tmp, p, jl_fma(c1, dt, t));
        // k3 = f(uprev + dt*(a31*k1 + a32*k2), p, t + c2*dt)
        for (int i = 0; i < n; ++i) tmp[i] = jl_fma(dt, jl_fma(a32, k2[i], a31 * k1[i]), uprev[i]);
        P->f(k3, tmp, p, jl_fma(c2, dt, t));
        for (int i = 0; i < n; ++i)
            tmp[i] = jl_fma(dt, jl_fma(a43, k3[i], jl_fma(a42, k2[i], a41 * k1[i])), uprev[i]);
        P->f(k4, tmp, p, jl_fma(c3, dt, t));
        for (int i = 0; i < n; ++i)
            tmp[i] = jl_fma(dt, jl_fma(a54, k4[i], jl_fma(a53, k3[i], jl_fma(a52, k2[i], a51 * k1[i]))), uprev[i]);
        P->f(k5, tmp, p, jl_fma(c4, dt, t));
        for (int i = 0; i < n; ++i)
            g6[i] = jl_fma(dt, jl_fma(a65, k5[i], jl_fma(a64, k4[i], jl_fma(a63, k3[i], jl_fma(a62, k2[i], a61 * k1[i])))),
                           uprev[i]);
        P->f(k6, g6, p, t + dt);
        for (int i = 0; i < n; ++i)
            u[i] = jl_fma(dt,
                          jl_fma(a76, k6[i],
                                 jl_fma(a75, k5[i], jl_fma(a74, k4[i], jl_fma(a73, k3[i], jl_fma(a72, k2[i], a71 * k1[i]))))),
                          uprev[i]);
        P->f(k7, u, p, t + dt);       // fsallast
        st.nf += 6;
        R atmp[ORACLE_MAXN];
        for (int i = 0; i < n; ++i) {
            R utilde = dt * jl_fma(btilde7, k7[i],
                                   jl_fma(btilde6, k6[i],
                                          jl_fma(btilde5, k5[i],
                                                 jl_fma(btilde4, k4[i],
                                                        jl_fma(btilde3, k3[i], jl_fma(btilde2, k2[i], btilde1 * k1[i]))))));
            atmp[i] = residual(utilde, uprev[i], u[i], o.atol(i), o.rtol(i));
        }
        return rms(atmp, n);
    }
    void update_fsal() { memcpy(k[0], k[6], sizeof(R) * P->n); }   // fsalfirst = fsallast
    void addsteps(const R*, const R*, const R*, R, R) {}            // length(k) >= 7: nothing to add
    static constexpr bool supports_callbacks = true;
    // reset_fsal! (integrator_utils.jl:1325-1343): fsalfirst = f(u, p, t), nf += 1
    void reset_fsal(const R* u, const R* p, R t, Stats<R>& st) { P->f(k[0], u, p, t); st.nf += 1; }
    // _ode_addsteps!(k, t, uprev, u, dt, f, p, ::Tsit5ConstantCache, always_calc_begin = true)
    // (lib/OrdinaryDiffEqTsit5/src/tsit_perform_step.jl:40-82): all seven stages again from uprev with the (shortened)
    // dt — note `uprev + dt*(a21*k1)`, not perform_step!'s `a = dt*a21; uprev + a*k1`; stats are not touched
    void addsteps_always(const R* uprev, const R* p, R t, R dt) {
        const int n = P->n;
        const R c1 = (R)0.161, c2 = (R)0.327, c3 = (R)0.9, c4 = (R)0.9800255409045097;
        const R a21 = (R)0.161, a31 = (R)-0.008480655492356989, a32 = (R)0.335480655492357,
                a41 = (R)2.8971530571054935, a42 = (R)-6.359448489975075, a43 = (R)4.3622954328695815,
                a51 = (R)5.325864828439257, a52 = (R)-11.748883564062828, a53 = (R)7.4955393428898365,
                a54 = (R)-0.09249506636175525, a61 = (R)5.86145544294642, a62 = (R)-12.92096931784711,
                a63 = (R)8.159367898576159, a64 = (R)-0.071584973281401, a65 = (R)-0.028269050394068383,
                a71 = (R)0.09646076681806523, a72 = (R)0.01, a73 = (R)0.4798896504144996,
                a74 = (R)1.379008574103742, a75 = (R)-3.290069515436081, a76 = (R)2.324710524099774;
        R *k1 = k[0], *k2 = k[1], *k3 = k[2], *k4 = k[3], *k5 = k[4], *k6 = k[5], *k7 = k[6];
        R tmp[ORACLE_MAXN];
        P->f(k1, uprev, p, t);
        for (int i = 0; i < n; ++i) tmp[i] = jl_fma(dt, a21 * k1[i], uprev[i]);
        P->f(k2, tmp, p, jl_fma(c1, dt, t));
        for (int i = 0; i < n; ++i) tmp[i] = jl_fma(dt, jl_fma(a32, k2[i], a31 * k1[i]), uprev[i]);
        P->f(k3, tmp, p, jl_fma(c2, dt, t));
        for (int i = 0; i < n; ++i) tmp[i] = jl_fma(dt, jl_fma(a43, k3[i], jl_fma(a42, k2[i], a41 * k1[i])), uprev[i]);
        P->f(k4, tmp, p, jl_fma(c3, dt, t));
        for (int i = 0; i < n; ++i)
            tmp[i] = jl_fma(dt, jl_fma(a54, k4[i], jl_fma(a53, k3[i], jl_fma(a52, k2[i], a51 * k1[i]))), uprev[i]);
        P->f(k5, tmp, p, jl_fma(c4, dt, t));
        for (int i = 0; i < n; ++i)
            tmp[i] = jl_fma(dt, jl_fma(a65, k5[i], jl_fma(a64, k4[i], jl_fma(a63, k3[i], jl_fma(a62, k2[i], a61 * k1[i])))), uprev[i]);
        P->f(k6, tmp, p, t + dt);
        for (int i = 0; i < n; ++i)
            tmp[i] = jl_fma(dt, jl_fma(a76, k6[i], jl_fma(a75, k5[i], jl_fma(a74, k4[i], jl_fma(a73, k3[i], jl_fma(a72, k2[i], a71 * k1[i]))))), uprev[i]);
        P->f(k7, tmp, p, t + dt);
    }
    void interpolant(R Theta, R dt, const R* y0, const R*, R* out) const {
        const int n = P->n;
        const R r11 = (R)1.0, r12 = (R)-2.763706197274826, r13 = (R)2.9132554618219126, r14 = (R)-1.0530884977290216,
                r22 = (R)0.13169999999999998, r23 = (R)-0.2234, r24 = (R)0.1017,
                r32 = (R)3.9302962368947516, r33 = (R)-5.941033872131505, r34 = (R)2.490627285651253,
                r42 = (R)-12.411077166933676, r43 = (R)30.33818863028232, r44 = (R)-16.548102889244902,
                r52 = (R)37.50931341651104, r53 = (R)-88.1789048947664, r54 = (R)47.37952196281928,
                r62 = (R)-27.896526289197286, r63 = (R)65.09189467479366, r64 = (R)-34.87065786149661,
                r72 = (R)1.5, r73 = (R)-4.0, r74 = (R)2.5;
        R Theta2 = Theta * Theta;
        R b1 = Theta * jl_fma(Theta, jl_fma(Theta, jl_fma(Theta, r14, r13), r12), r11);
        R b2 = Theta2 * jl_fma(Theta, jl_fma(Theta, r24, r23), r22);
        R b3 = Theta2 * jl_fma(Theta, jl_fma(Theta, r34, r33), r32);
        R b4 = Theta2 * jl_fma(Theta, jl_fma(Theta, r44, r43), r42);
        R b5 = Theta2 * jl_fma(Theta, jl_fma(Theta, r54, r53), r52);
        R b6 = Theta2 * jl_fma(Theta, jl_fma(Theta, r64, r63), r62);
        R b7 = Theta2 * jl_fma(Theta, jl_fma(Theta, r74, r73), r72);
        for (int i = 0; i < n; ++i)
            out[i] = jl_fma(dt,
                            jl_fma(k[6][i], b7,
                                   jl_fma(k[5][i], b6,
                                          jl_fma(k[4][i], b5,
                                                 jl_fma(k[3][i], b4, jl_fma(k[2][i], b3, jl_fma(k[1][i], b2, k[0][i] * b1)))))),
                            y0[i]);
    }
    static R qsteady_max() { return (R)1; }
    static bool fsal_init() { return true; }
};

#if __has_include("oracle_verner_gen.inc")
#include "oracle_verner_gen.inc"
#endif
#if __has_include("oracle_lowrk.inc")
#include "oracle_lowrk.inc"
#endif
#if __has_include("oracle_vern7.inc")
#include "oracle_vern7.inc"
#define ORACLE_HAVE_VERN7 1
#endif
#if __has_include("oracle_rosenbrock.inc")
#include "oracle_rosenbrock.inc"
#define ORACLE_HAVE_ROSENBROCK 1
#if __has_include("oracle_composite.inc")
#include "oracle_composite.inc"
#define ORACLE_HAVE_COMPOSITE 1
#endif
#endif

// ---------------------------------------------------------------------------
// _ode_initdt_oop — lib/OrdinaryDiffEqCore/src/initdt.jl:346-459 (g === nothing); `dtmax` is the signed value of
// auto_dt_reset! (integrator_interface.jl:645-648: tdir * min(|opts.dtmax|, |first_tstop - tdir t|)), the result carries tdir
template <typename R>
static R ode_initdt(const ProblemFns<R>& P, const R* u0, const R* p, R t, R dtmax, R abstol, R reltol, R opts_dtmin,
                    int order, const R* abstol_v = nullptr, const R* reltol_v = nullptr, R tdir = (R)1) {
    const int n = P.n;
    R dtmax_tdir = tdir * dtmax;
    R dtmin = jl_nextfloat(jl_max(opts_dtmin, jl_eps(t)));
    R smalldt = jl_max(dtmin, (R)1e-6);                 // convert(_tType, 1//10^6)
    R sk[ORACLE_MAXN], tmp[ORACLE_MAXN], f0[ORACLE_MAXN], f1[ORACLE_MAXN], u1[ORACLE_MAXN];
    for (int i = 0; i < n; ++i)      // abstol + |u0|*reltol (@muladd), element-wise for vector tolerances
        sk[i] = jl_fma(std::fabs(u0[i]), reltol_v ? reltol_v[i] : reltol, abstol_v ? abstol_v[i] : abstol);
    for (int i = 0; i < n; ++i) tmp[i] = u0[i] / sk[i];
    R d0 = rms(tmp, n);
    P.f(f0, u0, p, t);
    for (int i = 0; i < n; ++i) if (std::isnan(f0[i])) return tdir * dtmin;          // NAN_CHECK(f₀)
    for (int i = 0; i < n; ++i) tmp[i] = f0[i] / sk[i];
    R d1 = rms(tmp, n);
    if (std::isnan(d1)) return tdir * dtmin;
    R dt0;
    // d₀ < 1//10^5: the binary64 literal 1e-5 lies above the rational, so for any
    // binary64/binary32 operand "x < 1//10^5" == "(double)x < 1e-5"
    if ((double)d0 < 1e-5 || (double)d1 < 1e-5) dt0 = smalldt;
    else dt0 = (d0 / d1) / (R)100;
    dt0 = jl_min(dt0, dtmax_tdir);
    R dt0_tdir = tdir * dt0;
    for (int i = 0; i < n; ++i) u1[i] = jl_fma(dt0_tdir, f0[i], u0[i]);
    P.f(f1, u1, p, t + dt0_tdir);
    bool eq = true;
    for (int i = 0; i < n; ++i) eq = eq && (f0[i] == f1[i]);
    if (eq) return tdir * jl_max(dtmin, (R)100 * dt0);
    for (int i = 0; i < n; ++i) tmp[i] = (f1[i] - f0[i]) / sk[i];
    R d2 = rms(tmp, n) / dt0;
    R max_d1d2 = jl_max(d1, d2);
    R dt1;
    // max_d₁d₂ <= 1//Int64(10)^15: binary64 1e-15 lies above the rational => "<" on doubles
    if ((double)max_d1d2 < 1e-15) dt1 = jl_max(smalldt, dt0 * (R)0.001);             // dt₀ * 1//10^3
    else if (std::isinf(max_d1d2)) dt1 = (R)0;
    else {
        R l = (R)cr_log10((double)max_d1d2);
        R e = -((R)2 + l) / (R)order;
        dt1 = (R)cr_exp10((double)e);
    }
    return tdir * jl_max(dtmin, jl_min(jl_min((R)100 * dt0, dt1), dtmax_tdir));
}

// ---------------------------------------------------------------------------
template <typename R> struct Out {
    R* u_final; R* t_final; R* us; int nslots;
    int *nsaved, *naccept, *nreject, *nf, *njacs, *nw, *nsolve, *retcode;
    // ragged output (save_everystep): trajectory i owns rows row_offsets[i]..row_offsets[i+1]-1 of us/ts_rag
    const long long* row_offsets = nullptr; R* ts_rag = nullptr;
    // dense = true: every saved row also keeps the stepper cache (its k array), the way sol.k does
    // (integrator_utils.jl:455-473); points to a DenseSink<R, Alg> owned by dense_one
    void* dense_sink = nullptr;
    // save_idxs (0-based): saved rows hold only these components (integrator_utils.jl:368-375); NULL = all
    const int* save_idxs = nullptr; int nsave = 0;
};

template <typename R, typename Alg> struct DenseSink {
    std::vector<R> ts; std::vector<R> us; std::vector<Alg> ks;
};

template <typename A, typename = void> struct PIBeta {
    static double b2() { return 2.0 / (5.0 * A::order); }
    static double b1() { return 7.0 / (10.0 * A::order); }
};
template <typename A> struct PIBeta<A, std::void_t<decltype(A::beta2())>> {
    static double b2() { return A::beta2(); }
    static double b1() { return A::beta1(); }
};

// CompositeAlgorithm (AutoTsit5(Rosenbrock23()), oracle_composite.inc): one PI controller cache per branch
// (CompositeController, controllers.jl:1254-1338), choose_algorithm! in loopheader! (integrator_utils.jl:121),
// do_error_check (composite_algs.jl:37-42, solve.jl:909)
template <typename A, typename = void> struct SupportsCallbacks { static constexpr bool value = false; };
template <typename A> struct SupportsCallbacks<A, std::void_t<decltype(A::supports_callbacks)>> { static constexpr bool value = A::supports_callbacks; };
template <typename A, typename = void> struct IsComposite { static constexpr bool value = false; };
template <typename A> struct IsComposite<A, std::void_t<decltype(A::is_composite)>> { static constexpr bool value = A::is_composite; };

// One trajectory: __init + solve! + postamble!
template <typename R, typename Alg>
static void solve_one(const ProblemFns<R>& P, const R* u0, const R* p_in, R t0, R tf, const Opts<R>& o, long long idx,
                      const Out<R>& out) {
    const int n = P.n;
    Alg cache; cache.P = &P;
    Stats<R> stats;
    R u[ORACLE_MAXN], uprev[ORACLE_MAXN];
    for (int i = 0; i < n; ++i) { u[i] = u0[i]; uprev[i] = u0[i]; }
    // affect! may change the parameters of its own trajectory: work on a private copy when there are callbacks
    R p_local[ORACLE_MAXNP];
    if (o.ncb > 0 && p_in != nullptr) { for (int i = 0; i < P.np; ++i) p_local[i] = p_in[i]; }
    const R* p = (o.ncb > 0 && p_in != nullptr) ? p_local : p_in;
    R t = t0, tprev = t0;
    // tdir = sign(tspan[end] - tspan[1]) (solve.jl:273).  The internal queues of the reference hold tdir * time
    // (initialize_tstops / initialize_saveat / initialize_d_discontinuities, solve.jl:1021-1197); here the lists keep the
    // times themselves, in the order they are met, and every comparison multiplies both sides by tdir (exact)
    const R tdir = tf > t0 ? (R)1 : (tf < t0 ? (R)-1 : (R)0);
    // dtmax > 0 && tdir < 0 && (dtmax *= tdir) (solve.jl:401); dtmin is all abs
    const R dtmax = (o.dtmax > (R)0 && tdir < (R)0) ? o.dtmax * tdir : o.dtmax, opts_dtmin = o.dtmin;
    int nsaved = 0, save_idx = 0;
    R last_saved_t = t0;
    auto emit = [&](R ts, const R* v) {
        if (out.dense_sink) {
            auto* sink = (DenseSink<R, Alg>*)out.dense_sink;
            sink->ts.push_back(ts);
            for (int i = 0; i < n; ++i) sink->us.push_back(v[i]);
            sink->ks.push_back(cache);
        }
        const int w = out.save_idxs ? out.nsave : n;       // row width
        auto comp = [&](int i) { return out.save_idxs ? v[out.save_idxs[i]] : v[i]; };
        if (out.row_offsets) {
            if (out.us && nsaved < (int)(out.row_offsets[idx + 1] - out.row_offsets[idx])) {
                const size_t row = (size_t)out.row_offsets[idx] + (size_t)nsaved;
                for (int i = 0; i < w; ++i) out.us[row * w + i] = comp(i);
                out.ts_rag[row] = ts;
            }
        } else if (out.us && nsaved < out.nslots) {
            R* dst = out.us + ((size_t)idx * out.nslots + nsaved) * w;
            for (int i = 0; i < w; ++i) dst[i] = comp(i);
        }
        nsaved += 1; last_saved_t = ts;
    };
    int tstop_idx = 0;
    R cur_tstop = (o.ntstops > 0) ? o.tstops[0] : tf;                   // first(opts.tstops)
    // save_start (solve.jl:809-824)
    if (o.save_start) emit(t, u);
    // initialize!(integrator, cache) (solve.jl:831)
    if (Alg::fsal_init()) cache.initialize(uprev, p, t, stats);
    // handle_dt! (solve.jl:968-985): automatic dt when dt == 0 and adaptive
    R dt;
    const R dtcache = o.dt;                                            // solve.jl:699
    if (o.dt == (R)0 && o.adaptive) {
        R dtmax_init = tdir * jl_min(std::fabs(dtmax), std::fabs(tdir * cur_tstop - tdir * t));     // auto_dt_reset!: first_tstop
        dt = ode_initdt(P, u, p, t, dtmax_init, o.abstol, o.reltol, opts_dtmin, Alg::order, o.abstol_v, o.reltol_v, tdir);
        stats.nf += 2;
    } else if (o.adaptive && o.dt > (R)0 && tdir < (R)0) dt = o.dt * tdir;      // allow positive dt, but auto-convert (solve.jl:981-983)
    else dt = o.dt;
    R dtpropose = dt;
    // handle_starting_time_discontinuity! (solve.jl:872-901), the last act of init: a discontinuity at exactly t0 is popped,
    // t moves one ulp into the span and a first-same-as-last stepper re-evaluates its first stage there (reset_fsal!)
    int disc_idx = 0;
    if (o.ndisc > 0 && o.disc[0] == t) {
        disc_idx = 1;
        t = jl_shift_past(t, tdir);
        if constexpr (IsComposite<Alg>::value) cache.reset_fsal(u, p, t, stats);
        else if (Alg::fsal_init()) cache.initialize(u, p, t, stats);
    }
    // PIControllerCache (controllers.jl:793-803) and PIController defaults (alg_utils.jl)
    // beta2_default = 2//(5 order), beta1_default = 7//(10 order) (alg_utils.jl:766,788) unless the algorithm
    // overrides them (DP5); QT(rational) = correctly rounded quotient
    constexpr bool composite = IsComposite<Alg>::value;
    R beta2 = (R)0, beta1 = (R)0;
    const R qmin = (R)0.2, qmax = (R)10, gamma = (R)0.9, qoldinit = (R)1e-4, qmax_first_step = (R)10000;
    const R qsteady_min = (R)1;
    R qsteady_max = (R)1;
    // PIControllerCache state, one per branch of a composite algorithm (q11 = 1, errold = qoldinit)
    R q11_b[2] = {(R)1, (R)1}, errold_b[2] = {qoldinit, qoldinit};
    R EEst = (R)1;
    int br = 0;                       // branch whose controller cache is active (cache.current - 1)
    bool do_error_check = true;
    auto select_branch = [&]() {      // controller parameters of the active branch
        if constexpr (composite) {
            br = cache.current - 1;
            beta2 = (R)cache.beta2_cur(); beta1 = (R)cache.beta1_cur(); qsteady_max = cache.qsteady_max_cur();
        } else {
            beta2 = (R)PIBeta<Alg>::b2(); beta1 = (R)PIBeta<Alg>::b1(); qsteady_max = Alg::qsteady_max();
        }
    };
    select_branch();
#define q11 q11_b[br]
#define errold errold_b[br]
    long long iter = 0; int success_iter = 0, naccept = 0, nreject = 0;
    bool accept_step = false, next_step_tstop = false, isout = false;
    R tstop_target = cur_tstop;
    int retcode = RC_DEFAULT;

    // modify_dt_for_tstops! (integrator_utils.jl:268-324), adaptive branch
    auto modify_dt_for_tstops = [&]() {
        R tdir_t = tdir * t, tdir_tstop = tdir * cur_tstop;             // first_tstop(integrator)
        R distance_to_tstop = std::fabs(tdir_tstop - tdir_t);
        // tstop_tol = 100 eps(max(|t|, |tstop|)) when both are finite, else zero (integrator_utils.jl:277-286)
        R tstop_tol = (Bits<R>::finite(tdir_tstop) && Bits<R>::finite(t)) ? (R)100 * jl_eps(jl_max(std::fabs(t), std::fabs(tdir_tstop))) : (R)0;
        if (o.adaptive) {
            R original_dt = std::fabs(dt);
            dtpropose = tdir * original_dt;
            if (original_dt + tstop_tol < distance_to_tstop) next_step_tstop = false;
            else { next_step_tstop = true; tstop_target = tdir * tdir_tstop; }
            dt = tdir * jl_min(original_dt, distance_to_tstop);
        } else if (dtcache == (R)0) {                                   // (:300-304) step from stop to stop
            dt = tdir * distance_to_tstop;
            next_step_tstop = true; tstop_target = tdir * tdir_tstop;
        } else {                                                        // (:305-316) dtchangeable, !force_stepfail
            if (std::fabs(dtcache) + tstop_tol < distance_to_tstop) next_step_tstop = false;
            else { next_step_tstop = true; tstop_target = tdir * tdir_tstop; }
            dt = tdir * jl_min(std::fabs(dtcache), distance_to_tstop);
        }
    };

    // ---- savevalues! / _savevalues! (integrator_utils.jl:336-414); returns savedexactly
    auto savevalues = [&](bool force_save) -> bool {
        bool savedexactly = false, added = false;
        while (save_idx < o.nsaveat && tdir * o.saveat[save_idx] <= tdir * t) {     // first(saveat) <= tdir_t; curt = tdir * pop!(saveat)
            R curt = o.saveat[save_idx++];
            if (curt != t) {
                R Theta = (curt - tprev) / dt;
                if (!added) { cache.addsteps(uprev, u, p, tprev, dt); added = true; }
                R val[ORACLE_MAXN];
                cache.interpolant(Theta, dt, uprev, u, val);
                emit(curt, val);
            } else {
                if (curt == tf && !o.save_end) continue;           // skip_saveat_at_tspan_end
                savedexactly = true;
                emit(t, u);
            }
        }
        // force_save || save_everystep branch (:385-411)
        if (force_save || (o.save_everystep &&
                           (nsaved == 0 || ((t != last_saved_t || dt == (R)0) && (o.save_end || t != tf))))) {
            savedexactly = true;
            emit(t, u);
        }
        return savedexactly;
    };

    // ---- callbacks (lib/DiffEqBase/src/callbacks.jl) --------------------------------------------------------------
    bool reeval_fsal = false, terminated = false;
    int event_last_time = 0;                 // 1-based index of the continuous callback whose event ended the last step
    R last_event_error = (R)0;
    typedef R (*cond_t)(const R*, const R*, R);
    typedef void (*affect_t)(R*, R*, R, int*);
    auto jl_sign = [](R x) -> R { return x > (R)0 ? (R)1 : (x < (R)0 ? (R)-1 : x); };      // sign(±0) = ±0, sign(NaN) = NaN
    // get_condition (callbacks.jl:91-137): u at t, uprev at tprev, the interpolant in between
    auto get_condition = [&](const OracleCallback& cb, R abst) -> R {
        cond_t f = (cond_t)cb.condition;
        if (abst == t) return f(u, p, abst);
        if (abst == tprev) return f(uprev, p, abst);
        R val[ORACLE_MAXN];
        R Theta = (abst - tprev) / dt;                                 // current_interpolant
        cache.addsteps(uprev, u, p, tprev, dt);
        cache.interpolant(Theta, dt, uprev, u, val);
        return f(val, p, abst);
    };
    // is_event_occurrence (callbacks.jl:523-528)
    auto is_event = [&](const OracleCallback& cb, R prev_sign, R next_sign) -> bool {
        return ((prev_sign < (R)0 && cb.affect != nullptr) || (prev_sign > (R)0 && cb.affect_neg != nullptr)) &&
               prev_sign * next_sign <= (R)0;
    };
    // find_callback_time(integrator, callback::ContinuousCallback, callback_idx) (callbacks.jl:361-403)
    auto find_callback_time = [&](const OracleCallback& cb, int callback_idx, R& callback_t, R& bottom_sign, R& residual) -> bool {
        R bottom_t = tprev;
        R bottom_condition = get_condition(cb, bottom_t);
        if (event_last_time == callback_idx) {
            // nudge_tprev (:413-422): still within abstol of the last root => look just right of tprev
            if ((double)std::fabs(bottom_condition - last_event_error) <= cb.abstol) bottom_t = tprev + dt * (R)cb.repeat_nudge;
            else bottom_t = tprev;
            bottom_condition = get_condition(cb, bottom_t);
        }
        bottom_sign = jl_sign(bottom_condition);
        // check_event_occurrence (:427-446)
        R top_t = t;
        R top_sign = jl_sign(get_condition(cb, top_t));
        bool occurred = is_event(cb, bottom_sign, top_sign);
        if (cb.interp_points >= 2 && !occurred) {
            JlLinspace<R> ts(tprev, t, cb.interp_points);              // range(tprev, stop = t, length = interp_points)
            for (int i = 2; i <= cb.interp_points; ++i) {
                top_t = (i == cb.interp_points) ? t : ts[i];           // the last element is `stop` exactly
                top_sign = jl_sign(get_condition(cb, top_t));
                occurred = is_event(cb, bottom_sign, top_sign);
                if (occurred) break;
            }
        }
        if (!occurred) { callback_t = t; residual = (R)0; }
        else if (cb.rootfind == 0 || top_sign == (R)0) { callback_t = top_t; residual = (R)0; }
        else {
            // find_root (:478-491): IntervalNonlinearProblem solved with ModAB(), abstol = reltol = 0 — EXT
            // (BracketingNonlinearSolve).  With zero tolerances every bracketing method ends on the pair of adjacent
            // floats around the sign change, so plain bisection is used; an exact zero counts as the far side.
            // (backward integration: tup[1] > tup[2], "left" / "right" keep the order of the tuple — find_root's docstring)
            R left = bottom_t, right = top_t;
            for (;;) {
                R mid = left + (right - left) / (R)2;
                if (!(tdir * left < tdir * mid && tdir * mid < tdir * right)) break;
                R sm = jl_sign(get_condition(cb, mid));
                if (sm == bottom_sign) left = mid; else right = mid;
            }
            callback_t = (cb.rootfind == 1) ? left : right;
            residual = get_condition(cb, callback_t);
        }
        return occurred;
    };
    // reeval_internals_due_to_modification! (integrator_interface.jl:54-80)
    auto reeval_internals = [&](bool continuous_modification) {
        if constexpr (SupportsCallbacks<Alg>::value) {
            if (continuous_modification) cache.addsteps_always(uprev, p, tprev, dt);    // opts.calck is true with callbacks
        }
        reeval_fsal = true;
    };
    auto run_affect = [&](void* fn) {
        int term = 0;
        ((affect_t)fn)(u, p_local, t, &term);
        if (term) { terminated = true; retcode = RC_TERMINATED; }      // terminate!(integrator)
    };
    // apply_callback! (callbacks.jl:557-637)
    auto apply_callback = [&](const OracleCallback& cb, R cb_time, R prev_sign, bool& saved_in_cb) -> bool {
        if (o.adaptive) dtpropose = tdir * jl_max(jl_nextfloat(opts_dtmin), tdir * dt);   // set_proposed_dt!(tdir * max(nextfloat(dtmin), tdir * dtrelax * dt)), dtrelax = 1
        // change_t_via_interpolation! (integrator_interface.jl:5-39)
        if (cb_time != t) {
            R val[ORACLE_MAXN];
            R Theta = (cb_time - tprev) / dt;
            cache.addsteps(uprev, u, p, tprev, dt);
            cache.interpolant(Theta, dt, uprev, u, val);
            for (int i = 0; i < n; ++i) u[i] = val[i];
            t = cb_time;
            dt = t - tprev;
            reeval_internals(true);
        }
        bool savedexactly = savevalues(false);
        saved_in_cb = true;
        if (cb.save_before && !savedexactly) savevalues(true);
        void* fn = prev_sign < (R)0 ? cb.affect : (prev_sign > (R)0 ? cb.affect_neg : nullptr);
        if (fn == nullptr) return false;                            // derivative_discontinuity = false
        run_affect(fn);
        reeval_internals(true);
        if (cb.save_after) { savevalues(true); saved_in_cb = true; }
        return true;
    };
    auto handle_callbacks = [&]() {
        bool saved_in_cb = false;
        int ncont = 0;
        for (int k = 0; k < o.ncb; ++k) if (o.cbs[k].kind == 1) ncont += 1;
        if (ncont > 0) {
            // find_first_continuous_callback (:140-226): the earliest event wins, ties keep the first callback
            bool event_occurred = false; R tmin = t, upcrossing = (R)0, residual = (R)0; int identified = 0, ci = 0, evc = 0;
            for (int k = 0; k < o.ncb; ++k) {
                if (o.cbs[k].kind != 1) continue;
                ci += 1;
                R t2, s2, r2;
                bool occ2 = find_callback_time(o.cbs[k], ci, t2, s2, r2);
                if (ci == 1) { tmin = t2; upcrossing = s2; residual = r2; event_occurred = occ2; identified = k; }
                else if (occ2 && (!event_occurred || tdir * t2 < tdir * tmin)) {
                    tmin = t2; upcrossing = s2; residual = r2; event_occurred = true; identified = k;
                }
                if (event_occurred && identified == k) evc = ci;
            }
            if (event_occurred) {
                last_event_error = residual;
                event_last_time = evc;
                apply_callback(o.cbs[identified], tmin, upcrossing, saved_in_cb);
            } else event_last_time = 0;
        }
        // apply_discrete_callback! (:649-690), in order
        for (int k = 0; k < o.ncb; ++k) {
            const OracleCallback& cb = o.cbs[k];
            if (cb.kind != 0) continue;
            if (((cond_t)cb.condition)(u, p, t) != (R)0) {
                bool savedexactly = savevalues(false);
                saved_in_cb = true;
                if (cb.save_before && !savedexactly) savevalues(true);
                if (cb.affect) { run_affect(cb.affect); reeval_internals(false); }
                if (cb.save_after) { savevalues(true); saved_in_cb = true; }
            }
        }
        if (!saved_in_cb) savevalues(false);
    };

    // solve! (solve.jl:904-946): `while !isempty(tstops); while t < first(tstops) ... end; handle_tstop! end`.
    // tf is the last stop, so the two loops collapse into this one plus the pop at the end of an accepted step.
    while (tdir * t < tdir * tf) {
        // ---- loopheader! (integrator_utils.jl:84-127)
        if (iter > 0) {
            if (accept_step || !o.adaptive) {                          // (:98-110)
                success_iter += 1;
                // apply_step! (:175-203)
                for (int i = 0; i < n; ++i) uprev[i] = u[i];
                dt = dtpropose;
                // update_fsal! (:215-239): reeval_fsal / derivative_discontinuity => reset_fsal!
                // (for every FSAL stepper initialize! IS "fsalfirst = f(u, p, t); nf += 1"; steppers that are not FSAL have
                //  nothing to refresh; the composite algorithm re-evaluates in its current branch)
                // first branch (:216-220): has_discontinuity && first_discontinuity == t => pop it, shift t one ulp past
                // it, reset_fsal! for a first-same-as-last stepper
                if (disc_idx < o.ndisc && o.disc[disc_idx] == t) {
                    disc_idx += 1;
                    t = jl_shift_past(t, tdir);
                    if constexpr (IsComposite<Alg>::value) cache.reset_fsal(u, p, t, stats);
                    else if (Alg::fsal_init()) cache.initialize(u, p, t, stats);
                } else if (reeval_fsal) {
                    if constexpr (IsComposite<Alg>::value) cache.reset_fsal(u, p, t, stats);
                    else if (Alg::fsal_init()) cache.initialize(u, p, t, stats);
                } else cache.update_fsal();
                modify_dt_for_tstops();
            } else {
                // handle_step_rejection! (:129-139): isout => dt * qmin, else step_reject_controller! (controllers.jl:838-843)
                if (isout) dt = dt * qmin;
                else dt = dt / jl_min((R)1 / qmin, q11 / gamma);
            }
        }
        iter += 1;
        if constexpr (composite) {      // choose_algorithm!(integrator, integrator.cache) (:121)
            cache.choose_algorithm(dt, uprev, p, t, stats, do_error_check);
            select_branch();
        }
        // fix_dt_at_bounds! (:1243-1256); timedepentdtmin = max(eps(t), dtmin)
        // (for tdir < 0 the reference clamps with max(dtmax, dt) and then takes min(dt, dtmin) against the POSITIVE dtmin —
        //  a no-op on a negative dt; restated as written)
        if (tdir > (R)0) { dt = jl_min(dtmax, dt); if (!o.mirror_of_reverse) dt = jl_max(dt, jl_max(jl_eps(t), opts_dtmin)); }
        else { dt = jl_max(dtmax, dt); dt = jl_min(dt, std::fabs(jl_max(jl_eps(t), opts_dtmin))); }
        modify_dt_for_tstops();
        // ---- check_error (lib/DiffEqBase/src/check_error.jl:70-118); `integrator.do_error_check &&` (solve.jl:909)
        if (do_error_check) {
            int code = RC_SUCCESS;
            if (std::isnan(dt)) code = RC_DTNAN;
            else if (iter > o.maxiters) code = RC_MAXITERS;
            else if (o.adaptive && std::fabs(dt) <= std::fabs(opts_dtmin) && (!accept_step || (o.mirror_of_reverse ? t + dt > cur_tstop : t + dt < cur_tstop))) code = RC_DTLESSTHANMIN;   // t + dt < tdir * first(opts.tstops), as written (check_error.jl:96)
            else if (o.adaptive && !accept_step && std::fabs(dt) <= std::fabs(jl_eps(t))) code = RC_UNSTABLE;
            else if (accept_step) {
                for (int i = 0; i < n; ++i) if (!Bits<R>::finite(u[i])) code = RC_UNSTABLE;
            }
            if (code != RC_SUCCESS) { retcode = code; break; }
        }
        // ---- perform_step! or handle_tstop_step! (:326-333)
        if (next_step_tstop && std::fabs(dt) < jl_eps(std::fabs(t))) {
            accept_step = true;
        } else {
            EEst = cache.perform_step(uprev, u, p, t, dt, o, stats, o.nsaveat > 0 || out.dense_sink != nullptr);   // calck (solve.jl:147-148)
        }
        // ---- loopfooter! (:597-677)
        do_error_check = true;
        reeval_fsal = false;                                            // loopfooter_reset!
        R ttmp = t + dt;
        R q = (R)1;
        if (o.adaptive) {
            // stepsize_controller!(integrator, ::PIControllerCache, alg) (controllers.jl:805-821)
            R qmax_cur = (success_iter == 0) ? qmax_first_step : qmax;      // get_current_qmax (:288-293)
            if (EEst == (R)0) q = (R)1 / qmax_cur;
            else {
                R q11_new = fastpower(EEst, beta1);
                q = q11_new / fastpower(errold, beta2);
                q11 = q11_new;
                q = q / gamma;
                R lo = (R)1 / qmax_cur, hi = (R)1 / qmin;
                q = q < lo ? lo : (q > hi ? hi : q);
            }
            // isout = opts.isoutofdomain(u, p, ttmp); accept_step = !isout && accept_step_controller (:612-618, :245-250)
            isout = o.isout ? (((R (*)(const R*, const R*, R))o.isout)(u, p, ttmp) != (R)0) : false;
            accept_step = !isout && (EEst <= (R)1);
        } else {
            accept_step = true;                                            // not adaptive (:650-659)
        }
        if (accept_step) {
            naccept += 1;
            tprev = t;
            if (o.adaptive && next_step_tstop) dt = dtpropose;         // (:629-633)
            if (next_step_tstop) { next_step_tstop = false; t = tstop_target; } else t = ttmp;   // fixed_t_for_tstop_error!
            if (o.adaptive) {
                // step_accept_controller! (controllers.jl:823-836)
                if (qsteady_min <= q && q <= qsteady_max) q = (R)1;
                errold = jl_max(EEst, qoldinit);
                R dtnew = dt / q;
                // calc_dt_propose! (:1199-1210)
                dtpropose = tdir * jl_min(std::fabs(dtmax), std::fabs(dtnew));
                dtpropose = tdir * jl_max(std::fabs(dtpropose), std::fabs(jl_max(jl_eps(t), opts_dtmin)));
            } else {
                dtpropose = dt;
            }
            // handle_callbacks! (integrator_utils.jl:1081-1132) -> savevalues! (:340-414)
            if (o.ncb > 0) handle_callbacks(); else savevalues(false);
            if (terminated) break;                                      // terminate!: the tstops heap was emptied
            // handle_tstop! (integrator_utils.jl:1290-1314): pop every copy of a stop that was reached
            while (o.ntstops > 0 && t == cur_tstop && tstop_idx + 1 < o.ntstops) cur_tstop = o.tstops[++tstop_idx];
        } else {
            nreject += 1;
        }
    }
    // postamble! -> solution_endpoint_match_cur_integrator! (:540-587)
    if (o.save_end &&
        (nsaved == 0 || (last_saved_t != t && (o.save_end_user || t == tf || o.nsaveat == 0))))
        emit(t, u);
    if (retcode == RC_DEFAULT) retcode = RC_SUCCESS;
    for (int i = 0; i < n; ++i) out.u_final[(size_t)idx * n + i] = u[i];
    if (out.t_final) out.t_final[idx] = t;
    if (out.nsaved) out.nsaved[idx] = nsaved;
    if (out.naccept) out.naccept[idx] = naccept;
    if (out.nreject) out.nreject[idx] = nreject;
    if (out.nf) out.nf[idx] = stats.nf;
    if (out.njacs) out.njacs[idx] = stats.njacs;
    if (out.nw) out.nw[idx] = stats.nw;
    if (out.nsolve) out.nsolve[idx] = stats.nsolve;
    if (out.retcode) out.retcode[idx] = retcode;
#undef q11
#undef errold
}

template <typename R, typename Alg>
static void dense_one(const ProblemFns<R>& P, const R* u0, const R* p, R t0, R tf, const Opts<R>& o, long long idx,
                      const Out<R>& out_in, const R* tq, int M, R* dense_out);

// per-trajectory time spans (a prob_func that remakes the problem with its own tspan): (t0_i, tf_i) pairs or NULL;
// dtmax_default: opts.dtmax was not given, so each trajectory's dtmax is its own tf_i - t0_i (solve.jl:152)
static const double* g_tspans = nullptr;
static bool g_dtmax_default = false;

template <typename R, typename Alg>
static void solve_batch(const ProblemFns<R>& P, long long N, const R* u0, int u0_shared, const R* p, int p_shared, R t0,
                        R tf, const Opts<R>& o, const Out<R>& out, int nthreads, const R* tq = nullptr, int M = 0,
                        R* dense_out = nullptr) {
    const double* tspans = g_tspans;
    const bool dtmax_default = g_dtmax_default;
    // the structural analogue of EnsembleThreads' Threads.@threads over trajectories
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (long long i = 0; i < N; ++i) {
        const R* ui = u0_shared ? u0 : u0 + (size_t)i * P.n;
        const R* pi = p_shared ? p : p + (size_t)i * P.np;
        if (tspans != nullptr) {
            // every trajectory is its own solve(prob_i, alg; kwargs...): its span, its default dtmax, and of a saveat
            // list the entries inside (t0_i, tf_i] (solve.jl:1103-1124 filters the grid against the problem's own tspan)
            const R t0i = (R)tspans[2 * i], tfi = (R)tspans[2 * i + 1];
            Opts<R> oi = o;
            if (dtmax_default) oi.dtmax = tfi - t0i;
            int a = 0;
            while (a < o.nsaveat && !(o.saveat[a] > t0i)) ++a;
            int b = a;
            while (b < o.nsaveat && o.saveat[b] <= tfi) ++b;
            oi.saveat = o.saveat + a; oi.nsaveat = b - a;
            solve_one<R, Alg>(P, ui, pi, t0i, tfi, oi, i, out);
            continue;
        }
        if (dense_out) dense_one<R, Alg>(P, ui, pi, t0, tf, o, i, out, tq, M, dense_out);
        else solve_one<R, Alg>(P, ui, pi, t0, tf, o, i, out);
    }
}

// sol(tq) post hoc — ode_interpolation (dense/generic_dense.jl:833-867) over the stored (ts, us, ks):
// i+ = min(lastindex, max(previous i+, searchsortedfirst(ts, t))) starting from i+ = 2, i- = i+ - 1,
// dt = ts[i+] - ts[i-], Θ = (t - ts[i-]) / dt, then _ode_addsteps!(ks[i+], ts[i-], u[i-], u[i+], dt, ...)
// (lazy stages only: the stored k already holds the step's own stages) and ode_interpolant.
template <typename R, typename Alg>
static void dense_one(const ProblemFns<R>& P, const R* u0, const R* p, R t0, R tf, const Opts<R>& o, long long idx,
                      const Out<R>& out_in, const R* tq, int M, R* dense_out) {
    DenseSink<R, Alg> sink;
    Out<R> out = out_in;
    out.dense_sink = &sink; out.us = nullptr; out.row_offsets = nullptr;
    solve_one<R, Alg>(P, u0, p, t0, tf, o, idx, out);
    const int n = P.n;
    const int nrows = (int)sink.ts.size();
    const R tdir = tf < t0 ? (R)-1 : (R)1;
    R* dst = dense_out + (size_t)idx * M * n;
    int hi = 1;
    for (int j = 0; j < M; ++j) {
        const R t = tq[j];
        R* v = dst + (size_t)j * n;
        if (nrows < 2) { for (int i = 0; i < n; ++i) v[i] = nrows == 1 ? sink.us[i] : (R)0; continue; }
        while (hi < nrows - 1 && tdir * sink.ts[hi] < tdir * t) hi += 1;     // searchsortedfirst by tdir * t (generic_dense.jl:838-849)
        const int ip = hi, im = hi - 1;
        const R dt = sink.ts[ip] - sink.ts[im];
        const R Theta = (dt == (R)0) ? (R)1 : (t - sink.ts[im]) / dt;
        Alg k = sink.ks[ip];
        const R* um = &sink.us[(size_t)im * n];
        const R* up = &sink.us[(size_t)ip * n];
        k.addsteps(um, up, p, sink.ts[im], dt);
        k.interpolant(Theta, dt, um, up, v);
    }
}

struct OracleArgs {
    int alg, dtype, n, np;
    void *rhs, *jac, *tgrad;
    long long N;
    const void* u0; int u0_shared;
    const void* p; int p_shared;
    double t0, tf;
    double reltol, abstol, dt, dtmin, dtmax;
    long long maxiters;
    const double* saveat; int nsaveat;
    int save_start, save_end;     // -1 default
    int linsolve;
    int nthreads;
    // outputs
    void* u_final; void* t_final; void* us; int nslots;
    int *nsaved, *naccept, *nreject, *nf, *njacs, *nw, *nsolve, *retcode;
    // save_everystep: row_offsets == NULL is the counting pass
    int save_everystep; const long long* row_offsets; void* ts_rag;
    const int* save_idxs; int nsave_idxs;
    const double* tstops; int ntstops;      // the tstops keyword, unfiltered
    int fixed_dt;                           // 1: adaptive = false
    const OracleCallback* cbs; int ncb;     // the CallbackSet (Tsit5 only)
    const double* abstol_v; const double* reltol_v;    // per-component tolerances (n entries each) or NULL
    const double* disc; int ndisc;          // the d_discontinuities keyword, unfiltered
    const double* tspans;                   // per-trajectory (t0_i, tf_i) pairs or NULL
    int mirror_of_reverse;                  // test switch, see Opts::mirror_of_reverse
};

template <typename R> static int run(const OracleArgs& a, const double* tq64 = nullptr, int M = 0, void* dense_out = nullptr) {
    ProblemFns<R> P;
    P.f = (typename Fn<R>::rhs_t)a.rhs; P.jac = (typename Fn<R>::rhs_t)a.jac; P.tgrad = (typename Fn<R>::rhs_t)a.tgrad;
    P.n = a.n; P.np = a.np;
    std::vector<R> grid(a.nsaveat > 0 ? a.nsaveat : 0);
    for (int i = 0; i < a.nsaveat; ++i) grid[i] = (R)a.saveat[i];
    Opts<R> o;
    o.reltol = (R)(a.reltol > 0 ? a.reltol : 1e-3);
    o.abstol = (R)(a.abstol > 0 ? a.abstol : 1e-6);
    o.dt = (R)a.dt; o.dtmin = (R)a.dtmin;
    // dtmax: the default is tspan[2] - tspan[1] (signed, solve.jl:152); for a reversed span either sign may be given (solve.jl:401)
    o.dtmax = (R)((a.dtmax > 0 || (a.dtmax < 0 && a.tf < a.t0)) ? a.dtmax : (a.tf - a.t0));
    o.maxiters = a.maxiters > 0 ? a.maxiters : 1000000;
    std::vector<R> atv, rtv;
    if (a.abstol_v) { atv.assign(a.abstol_v, a.abstol_v + a.n); o.abstol_v = atv.data(); }
    if (a.reltol_v) { rtv.assign(a.reltol_v, a.reltol_v + a.n); o.reltol_v = rtv.data(); }
    o.saveat = grid.data(); o.nsaveat = a.nsaveat;
    o.save_start = a.save_start != 0;
    o.save_end = a.save_end != 0;
    o.save_end_user = a.save_end > 0;
    o.linsolve = a.linsolve;
    o.save_everystep = a.save_everystep == 1;      // 2: ragged rows without the per-step rows (callbacks + saveat)
    o.adaptive = a.fixed_dt == 0;
    o.mirror_of_reverse = a.mirror_of_reverse != 0;
    std::vector<OracleCallback> real_cbs;
    if (a.ncb > 0) {
        if (!a.cbs) return -5;
        for (int i = 0; i < a.ncb; ++i) {
            if (a.cbs[i].kind == 2) o.isout = a.cbs[i].condition; else real_cbs.push_back(a.cbs[i]);
        }
        if (!real_cbs.empty()) {
            // ContinuousCallback needs the stepper's _ode_addsteps!(always_calc_begin = true): Tsit5 (first slice of SURVEY
            // §8(f) row 4); DiscreteCallback works with every stepper
            for (auto& c : real_cbs) if (c.kind == 1 && a.alg != ALG_TSIT5) return -5;
            o.cbs = real_cbs.data(); o.ncb = (int)real_cbs.size();
        }
    }
    if (!o.adaptive && a.dt == 0.0 && !(a.tstops && a.ntstops > 0)) return -4;     // solve.jl:277-280
    std::vector<R> stops, discs;
    if ((a.tstops && a.ntstops > 0) || (a.disc && a.ndisc > 0)) {
        // tdir_t0 < tdir * t < tdir_tf (solve.jl:1025-1036); the heaps order by tdir * t = the order the times are met
        const R td = a.tf < a.t0 ? (R)-1 : (R)1;
        for (int i = 0; a.tstops && i < a.ntstops; ++i) {
            R v = (R)a.tstops[i];
            if (td * v > td * (R)a.t0 && td * v < td * (R)a.tf) stops.push_back(v);
        }
        // d_discontinuities: the ones inside (t0, tf) are stops too (initialize_tstops, solve.jl:1033-1036); the heap of
        // discontinuities keeps every entry >= t0 (reinit_d_discontinuities!, solve.jl:1185-1197)
        for (int i = 0; a.disc && i < a.ndisc; ++i) {
            R v = (R)a.disc[i];
            if (td * v > td * (R)a.t0 && td * v < td * (R)a.tf) stops.push_back(v);
            if (td * v >= td * (R)a.t0) discs.push_back(v);
        }
        auto met_first = [td](R x, R y) { return td * x < td * y; };
        std::sort(stops.begin(), stops.end(), met_first);
        std::sort(discs.begin(), discs.end(), met_first);
        stops.push_back((R)a.tf);
        o.tstops = stops.data(); o.ntstops = (int)stops.size();
        o.disc = discs.data(); o.ndisc = (int)discs.size();
    }
    g_tspans = a.tspans; g_dtmax_default = !(a.dtmax > 0 || (a.dtmax < 0 && a.tf < a.t0));
    // reverse time (tf < t0): the integrator core, the callbacks (callbacks.jl:201,478-491,565-567) and the dense evaluation
    // above are direction-aware; per-trajectory spans are restated forward only
    if (a.tf < a.t0 && a.tspans) return -7;
    if (a.tspans) for (long long i = 0; i < a.N; ++i) if (!(a.tspans[2 * i + 1] > a.tspans[2 * i])) return -7;
    if (a.tspans && ((a.tstops && a.ntstops > 0) || (a.disc && a.ndisc > 0) || M > 0)) return -6;
    Out<R> out;
    out.row_offsets = a.row_offsets; out.ts_rag = (R*)a.ts_rag;
    if (a.save_idxs && a.nsave_idxs > 0) {
        for (int i = 0; i < a.nsave_idxs; ++i) if (a.save_idxs[i] < 0 || a.save_idxs[i] >= a.n) return -3;
        out.save_idxs = a.save_idxs; out.nsave = a.nsave_idxs;
    }
    out.u_final = (R*)a.u_final; out.t_final = (R*)a.t_final; out.us = (R*)a.us; out.nslots = a.nslots;
    out.nsaved = a.nsaved; out.naccept = a.naccept; out.nreject = a.nreject; out.nf = a.nf;
    out.njacs = a.njacs; out.nw = a.nw; out.nsolve = a.nsolve; out.retcode = a.retcode;
    const R* u0 = (const R*)a.u0; const R* p = (const R*)a.p;
    std::vector<R> tq(M > 0 ? M : 1);
    for (int j = 0; j < M; ++j) tq[j] = (R)tq64[j];
    switch (a.alg) {
        case ALG_TSIT5: solve_batch<R, Tsit5<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
#ifdef ORACLE_HAVE_VERN7
        case ALG_VERN7: solve_batch<R, Vern7<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
#endif
#ifdef ORACLE_HAVE_ROSENBROCK
        case ALG_ROS23: solve_batch<R, Rosenbrock23<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_ROS32: solve_batch<R, Rosenbrock32<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS5P: solve_batch<R, Rodas5P<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS5PE: solve_batch<R, Rodas5Pe<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS5: solve_batch<R, Rodas5<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS4: solve_batch<R, Rodas4<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS42: solve_batch<R, Rodas42<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS4P: solve_batch<R, Rodas4P<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS23W: solve_batch<R, Rodas23W<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS3P: solve_batch<R, Rodas3P<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_RODAS4P2: solve_batch<R, Rodas4P2<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
#endif
#ifdef ORACLE_HAVE_COMPOSITE
        case ALG_AUTOTSIT5_ROS23: solve_batch<R, AutoTsit5Ros23<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
#endif
#ifdef ORACLE_HAVE_VERNER_GEN
        case ALG_VERN6: solve_batch<R, Vern6<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_VERN8: solve_batch<R, Vern8<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_VERN9: solve_batch<R, Vern9<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_VERN7_GENERATED: solve_batch<R, Vern7Gen<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
#endif
#ifdef ORACLE_HAVE_LOWRK
        case ALG_DP5: solve_batch<R, DP5<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
        case ALG_BS3: solve_batch<R, BS3<R>>(P, a.N, u0, a.u0_shared, p, a.p_shared, (R)a.t0, (R)a.tf, o, out, a.nthreads, tq.data(), M, (R*)dense_out); break;
#endif
        default: return -2;
    }
    return 0;
}

extern "C" {
int oracle_solve(const OracleArgs* a) {
    if (!a || a->n < 1 || a->n > ORACLE_MAXN || !a->rhs) return -1;
    return a->dtype == 1 ? run<float>(*a) : run<double>(*a);
}
// dense = true: integrate with save_everystep, keep (ts, us, ks) per trajectory, evaluate sol(tq[j]);
// out is real[N][M][n]
int oracle_dense_solve(const OracleArgs* a, const double* tq, int M, void* out) {
    if (!a || a->n < 1 || a->n > ORACLE_MAXN || !a->rhs || !tq || M < 1 || !out) return -1;
    return a->dtype == 1 ? run<float>(*a, tq, M, out) : run<double>(*a, tq, M, out);
}
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
// scalar entry points for the known-answer tests
double oracle_fastpower(double x, double y) { return fastpower(x, y); }
float oracle_fastpower_f32(float x, float y) { return fastpower(x, y); }
double oracle_norm(const double* v, int n) { return rms(v, n); }
double oracle_eps(double x) { return jl_eps(x); }
double oracle_log10(double x) { return cr_log10(x); }
double oracle_exp10(double x) { return cr_exp10(x); }
double oracle_initdt(void* rhs, int n, int np, const double* u0, const double* p, double t0, double tf, double abstol,
                     double reltol, int order) {
    ProblemFns<double> P; P.f = (Fn<double>::rhs_t)rhs; P.n = n; P.np = np;
    return ode_initdt<double>(P, u0, p, t0, tf - t0, abstol, reltol, 0.0, order, nullptr, nullptr, tf < t0 ? -1.0 : 1.0);
}
}
