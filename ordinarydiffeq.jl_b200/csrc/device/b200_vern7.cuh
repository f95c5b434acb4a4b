// b200_vern7.cuh — Vern7 (Verner's "most efficient" 7/6 pair), 10 stages, not FSAL,
// with the lazy 7th-order interpolant (6 extra stages computed only for steps that
// contain a saveat point).
//
// Reference behaviour reproduced:
//   perform_step!(…, ::Vern7ConstantCache)   lib/OrdinaryDiffEqVerner/src/verner_rk_perform_step.jl:256-383
//   _ode_addsteps!(…, ::Vern7ConstantCache)   lib/OrdinaryDiffEqVerner/src/verner_addsteps.jl:654-800 (extra stages)
//   _ode_interpolant(…, ::Vern7ConstantCache) lib/OrdinaryDiffEqVerner/src/interpolants.jl:183-235
//   coefficients                              b200_tableaus_gen.cuh (generated from verner_tableaus.jl)
//   isfsal(::Vern7) = false, alg_order = 7     lib/OrdinaryDiffEqVerner/src/alg_utils.jl
#pragma once
#include "b200_base.cuh"
#include "b200_tableaus_gen.cuh"

// B200_VLEN: components held by one thread.  One-trajectory-per-thread kernels hold the whole
// state (B200_N); the sliced kernel (b200_sliced.cuh) holds ceil(n/G) components per thread.
#ifndef B200_VLEN
#define B200_VLEN B200_N
#endif
#ifndef B200_NORM
#define B200_NORM(res, u) b200_norm_local(res)
B200_D real b200_norm_local(const real* res) {
    real acc = res[0] * res[0];
#pragma unroll
    for (int i = 1; i < B200_VLEN; ++i) acc = acc + res[i] * res[i];
    bool bad = false;
    real e = b200_sqrt_fast(b200_div_const_fast(acc, (real)B200_N, (real)1 / (real)B200_N, bad), bad);
    if (bad) e = b200_sqrt(acc / (real)B200_N);
    return e;
}
#endif

struct B200Vern7Coeffs {
#define B200_X(name, val) real name;
    B200_VERN7_TABLEAU(B200_X)
    B200_VERN7_EXTRA(B200_X)
    B200_VERN7_INTERP(B200_X)
#undef B200_X
};
__constant__ B200Vern7Coeffs B200_VERN7_C = {
#define B200_X(name, val) (real)val,
    B200_VERN7_TABLEAU(B200_X)
    B200_VERN7_EXTRA(B200_X)
    B200_VERN7_INTERP(B200_X)
#undef B200_X
};

// Σ_j a_j k_j with MuladdMacro nesting: the last product is the outermost fma
#define B200_V7_2(a1, x1, a2, x2) b200_fma(C.a2, x2[i], C.a1 * x1[i])
#define B200_V7_3(a1, x1, a2, x2, a3, x3) b200_fma(C.a3, x3[i], B200_V7_2(a1, x1, a2, x2))
#define B200_V7_4(a1, x1, a2, x2, a3, x3, a4, x4) b200_fma(C.a4, x4[i], B200_V7_3(a1, x1, a2, x2, a3, x3))
#define B200_V7_5(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5) \
    b200_fma(C.a5, x5[i], B200_V7_4(a1, x1, a2, x2, a3, x3, a4, x4))
#define B200_V7_6(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6) \
    b200_fma(C.a6, x6[i], B200_V7_5(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5))
#define B200_V7_7(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6, a7, x7) \
    b200_fma(C.a7, x7[i], B200_V7_6(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6))
#define B200_V7_8(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6, a7, x7, a8, x8) \
    b200_fma(C.a8, x8[i], B200_V7_7(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6, a7, x7))
#define B200_V7_9(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6, a7, x7, a8, x8, a9, x9) \
    b200_fma(C.a9, x9[i], B200_V7_8(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6, a7, x7, a8, x8))
#define B200_V7_10(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6, a7, x7, a8, x8, a9, x9, a10, x10) \
    b200_fma(C.a10, x10[i], B200_V7_9(a1, x1, a2, x2, a3, x3, a4, x4, a5, x5, a6, x6, a7, x7, a8, x8, a9, x9))
// Component loops are fully unrolled by default; -DB200_STAGE_UNROLL=k rolls them (smaller code,
// but stage vectors are then indexed dynamically and live in local memory).
#ifndef B200_STAGE_UNROLL
#define B200_STAGE_UNROLL B200_VLEN   // full unroll: measured best also for n = 28 (see b200ode_shim.cu)
#endif
#define B200_PRAGMA_(x) _Pragma(#x)
#define B200_PRAGMA(x) B200_PRAGMA_(x)
#define B200_UNROLL_STAGE B200_PRAGMA(unroll B200_STAGE_UNROLL)
#define B200_V7_STAGE(dst, expr)                                                  \
    B200_UNROLL_STAGE for (int i = 0; i < B200_VLEN; ++i) dst[i] = b200_fma(dt, (expr), uprev[i]);

struct B200Vern7 {
    real k1[B200_VLEN], k2[B200_VLEN], k3[B200_VLEN], k4[B200_VLEN], k5[B200_VLEN], k6[B200_VLEN], k7[B200_VLEN], k8[B200_VLEN],
        k9[B200_VLEN], k10[B200_VLEN];
    real k11[B200_VLEN], k12[B200_VLEN], k13[B200_VLEN], k14[B200_VLEN], k15[B200_VLEN], k16[B200_VLEN];   // lazy extra stages

#ifdef B200_STEPPER_EXTRA_MEMBERS
    B200_STEPPER_EXTRA_MEMBERS
#endif
    static B200_D int order() { return 7; }
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1; }

    // initialize!: nothing is evaluated (get_fsalfirstlast is (nothing, nothing))
    B200_D void init(const real*, const real*, real, int&) {}

    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt, real reltol, real abstol,
                        int& nf) {
        const B200Vern7Coeffs& C = B200_VERN7_C;
        real tmp[B200_VLEN];
        B200_RHS(k1, uprev, p, t);
        const real a = dt * C.a021;
        B200_UNROLL_STAGE
        for (int i = 0; i < B200_VLEN; ++i) tmp[i] = b200_fma(a, k1[i], uprev[i]);
        B200_RHS(k2, tmp, p, b200_fma(C.c2, dt, t));
        B200_V7_STAGE(tmp, B200_V7_2(a031, k1, a032, k2))
        B200_RHS(k3, tmp, p, b200_fma(C.c3, dt, t));
        B200_V7_STAGE(tmp, B200_V7_2(a041, k1, a043, k3))
        B200_RHS(k4, tmp, p, b200_fma(C.c4, dt, t));
        B200_V7_STAGE(tmp, B200_V7_3(a051, k1, a053, k3, a054, k4))
        B200_RHS(k5, tmp, p, b200_fma(C.c5, dt, t));
        B200_V7_STAGE(tmp, B200_V7_4(a061, k1, a063, k3, a064, k4, a065, k5))
        B200_RHS(k6, tmp, p, b200_fma(C.c6, dt, t));
        B200_V7_STAGE(tmp, B200_V7_5(a071, k1, a073, k3, a074, k4, a075, k5, a076, k6))
        B200_RHS(k7, tmp, p, b200_fma(C.c7, dt, t));
        B200_V7_STAGE(tmp, B200_V7_6(a081, k1, a083, k3, a084, k4, a085, k5, a086, k6, a087, k7))
        B200_RHS(k8, tmp, p, b200_fma(C.c8, dt, t));
        B200_V7_STAGE(tmp, B200_V7_7(a091, k1, a093, k3, a094, k4, a095, k5, a096, k6, a097, k7, a098, k8))
        B200_RHS(k9, tmp, p, t + dt);
        B200_V7_STAGE(tmp, B200_V7_6(a101, k1, a103, k3, a104, k4, a105, k5, a106, k6, a107, k7))
        B200_RHS(k10, tmp, p, t + dt);
        nf += 10;
        B200_V7_STAGE(u, B200_V7_7(b1, k1, b4, k4, b5, k5, b6, k6, b7, k7, b8, k8, b9, k9))
        // calculate_residuals: flagged fast divisions (they interleave), plain operator in the flagged case
        real res[B200_VLEN], ut[B200_VLEN], den[B200_VLEN];
        bool bad = false;
        B200_UNROLL_STAGE
        for (int i = 0; i < B200_VLEN; ++i) {
            ut[i] = dt * B200_V7_8(btilde1, k1, btilde4, k4, btilde5, k5, btilde6, k6, btilde7, k7, btilde8, k8,
                                   btilde9, k9, btilde10, k10);
            den[i] = b200_fma(b200_max_fast(b200_abs(uprev[i]), b200_abs(u[i])), B200_RTOL_AT(i, reltol), B200_ATOL_AT(i, abstol));
            res[i] = b200_div_fast(ut[i], den[i], bad);
        }
        if (bad) {
            B200_UNROLL_STAGE
            for (int i = 0; i < B200_VLEN; ++i) res[i] = ut[i] / den[i];
        }
        return B200_NORM(res, u);
    }

    B200_D void accept() {}

    // _ode_addsteps! extra stages k11..k16 at (tprev, uprev) with the step's dt; not counted in nf
    B200_D void dense_prepare(const real* uprev, const real* /*u*/, const real* p, real t, real dt) {
        const B200Vern7Coeffs& C = B200_VERN7_C;
        real tmp[B200_VLEN];
        B200_V7_STAGE(tmp, B200_V7_7(a1101, k1, a1104, k4, a1105, k5, a1106, k6, a1107, k7, a1108, k8, a1109, k9))
        B200_RHS(k11, tmp, p, b200_fma(C.c11, dt, t));
        B200_V7_STAGE(tmp, B200_V7_8(a1201, k1, a1204, k4, a1205, k5, a1206, k6, a1207, k7, a1208, k8, a1209, k9,
                                     a1211, k11))
        B200_RHS(k12, tmp, p, b200_fma(C.c12, dt, t));
        B200_V7_STAGE(tmp, B200_V7_9(a1301, k1, a1304, k4, a1305, k5, a1306, k6, a1307, k7, a1308, k8, a1309, k9,
                                     a1311, k11, a1312, k12))
        B200_RHS(k13, tmp, p, b200_fma(C.c13, dt, t));
        B200_V7_STAGE(tmp, B200_V7_10(a1401, k1, a1404, k4, a1405, k5, a1406, k6, a1407, k7, a1408, k8, a1409, k9,
                                      a1411, k11, a1412, k12, a1413, k13))
        B200_RHS(k14, tmp, p, b200_fma(C.c14, dt, t));
        B200_V7_STAGE(tmp, B200_V7_10(a1501, k1, a1504, k4, a1505, k5, a1506, k6, a1507, k7, a1508, k8, a1509, k9,
                                      a1511, k11, a1512, k12, a1513, k13))
        B200_RHS(k15, tmp, p, b200_fma(C.c15, dt, t));
        B200_V7_STAGE(tmp, B200_V7_10(a1601, k1, a1604, k4, a1605, k5, a1606, k6, a1607, k7, a1608, k8, a1609, k9,
                                      a1611, k11, a1612, k12, a1613, k13))
        B200_RHS(k16, tmp, p, b200_fma(C.c16, dt, t));
    }

    B200_D void interp(real th, real dt, const real* y0, const real* /*y1*/, real* out) const {
        const B200Vern7Coeffs& C = B200_VERN7_C;
        const real th2 = th * th;
#define B200_P6(a, b, c, d, e, f) \
    b200_fma(th, b200_fma(th, b200_fma(th, b200_fma(th, b200_fma(th, C.f, C.e), C.d), C.c), C.b), C.a)
        const real b1 = th * b200_fma(th, B200_P6(r012, r013, r014, r015, r016, r017), C.r011);
        const real b4 = th2 * B200_P6(r042, r043, r044, r045, r046, r047);
        const real b5 = th2 * B200_P6(r052, r053, r054, r055, r056, r057);
        const real b6 = th2 * B200_P6(r062, r063, r064, r065, r066, r067);
        const real b7 = th2 * B200_P6(r072, r073, r074, r075, r076, r077);
        const real b8 = th2 * B200_P6(r082, r083, r084, r085, r086, r087);
        const real b9 = th2 * B200_P6(r092, r093, r094, r095, r096, r097);
        const real b11 = th2 * B200_P6(r112, r113, r114, r115, r116, r117);
        const real b12 = th2 * B200_P6(r122, r123, r124, r125, r126, r127);
        const real b13 = th2 * B200_P6(r132, r133, r134, r135, r136, r137);
        const real b14 = th2 * B200_P6(r142, r143, r144, r145, r146, r147);
        const real b15 = th2 * B200_P6(r152, r153, r154, r155, r156, r157);
        const real b16 = th2 * B200_P6(r162, r163, r164, r165, r166, r167);
#undef B200_P6
        B200_UNROLL_STAGE
        for (int i = 0; i < B200_VLEN; ++i) {
            real s = k1[i] * b1;
            s = b200_fma(k4[i], b4, s);
            s = b200_fma(k5[i], b5, s);
            s = b200_fma(k6[i], b6, s);
            s = b200_fma(k7[i], b7, s);
            s = b200_fma(k8[i], b8, s);
            s = b200_fma(k9[i], b9, s);
            s = b200_fma(k11[i], b11, s);
            s = b200_fma(k12[i], b12, s);
            s = b200_fma(k13[i], b13, s);
            s = b200_fma(k14[i], b14, s);
            s = b200_fma(k15[i], b15, s);
            s = b200_fma(k16[i], b16, s);
            out[i] = b200_fma(dt, s, y0[i]);
        }
    }
};
