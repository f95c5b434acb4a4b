// b200_lowrk.cuh — DP5 (Dormand-Prince 5(4)) and BS3 (Bogacki-Shampine 3(2)), the low-order
// explicit pairs of lib/OrdinaryDiffEqLowOrderRK, on the same per-thread skeleton as Tsit5
// (SURVEY §8(f) row 3).  One trajectory per thread, stage vectors in registers.
//
// Reference behaviour reproduced (arithmetic order and fusion included):
//   DP5  perform_step!(…, ::DP5ConstantCache)   low_order_rk_perform_step.jl:667-710
//        tableau (Float64 literals, convert(T,·)) low_order_rk_tableaus.jl:1096-1151
//        dense vectors k[1..4] = update, bspl, update-k7-bspl, Σ d_i k_i  (:704-708)
//        interpolant  y0 + dt*(k1 b10 + k2 b20 + k3 b30 + k4 b40)        interpolants.jl:29-42
//        PI exponents beta2 = 4//100, beta1 = 1//5 - 3 beta2/4 = 17//100  alg_utils.jl:37-39
//   BS3  perform_step!(…, ::BS3ConstantCache)   low_order_rk_perform_step.jl:13-36
//        tableau                                  low_order_rk_tableaus.jl:27-44
//        k = [fsalfirst, fsallast] -> cubic Hermite interpolant (OrdinaryDiffEqCore
//        dense/generic_dense.jl:1527-1537, differential_vars === nothing)
// Fusion follows MuladdMacro as everywhere else; an n-ary product in a sum
// splits off its LAST factor:  a*b*c + d -> muladd(a*b, c, d).
#pragma once
#include "b200_base.cuh"

struct B200DP5Coeffs {
    real c1, c2, c3, c4;
    real a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65, a71, a73, a74, a75, a76;
    real bt1, bt3, bt4, bt5, bt6, bt7;
    real d1, d3, d4, d5, d6, d7;
};
__constant__ B200DP5Coeffs B200_DP5_C = {
    (real)0.2, (real)0.3, (real)0.8, (real)0.8888888888888888,
    (real)0.2, (real)0.075, (real)0.225,
    (real)0.9777777777777777, (real)-3.7333333333333334, (real)3.5555555555555554,
    (real)2.9525986892242035, (real)-11.595793324188385, (real)9.822892851699436, (real)-0.2908093278463649,
    (real)2.8462752525252526, (real)-10.757575757575758, (real)8.906422717743473, (real)0.2784090909090909,
    (real)-0.2735313036020583,
    (real)0.09114583333333333, (real)0.44923629829290207, (real)0.6510416666666666, (real)-0.322376179245283,
    (real)0.13095238095238096,
    (real)-0.0012326388888888888, (real)0.0042527702905061394, (real)-0.03697916666666667,
    (real)0.05086379716981132, (real)-0.0419047619047619, (real)0.025,
    (real)-1.1270175653862835, (real)2.675424484351598, (real)-5.685526961588504, (real)3.5219323679207912,
    (real)-1.7672812570757455, (real)2.382468931778144,
};

struct B200DP5 {
    real k1[B200_N], k3[B200_N], k4[B200_N], k5[B200_N], k6[B200_N], k7[B200_N];
    real upd[B200_N];                       // `update`
    real dk2[B200_N], dk3[B200_N], dk4[B200_N];   // dense vectors k[2..4] (k[1] = update)

    static B200_D int order() { return 5; }
    static B200_D bool fsal() { return true; }
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1; }
    static B200_D real beta2() { return (real)(4.0 / 100.0); }
    static B200_D real beta1() { return (real)(17.0 / 100.0); }

    B200_D void init(const real* u, const real* p, real t, int& nf) {
        B200_RHS(k1, u, p, t);
        nf += 1;
    }

    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt,
                        real reltol, real abstol, int& nf) {
#define B200_DP(name) const real name = B200_DP5_C.name
        B200_DP(c1); B200_DP(c2); B200_DP(c3); B200_DP(c4);
        B200_DP(a21); B200_DP(a31); B200_DP(a32); B200_DP(a41); B200_DP(a42); B200_DP(a43);
        B200_DP(a51); B200_DP(a52); B200_DP(a53); B200_DP(a54);
        B200_DP(a61); B200_DP(a62); B200_DP(a63); B200_DP(a64); B200_DP(a65);
        B200_DP(a71); B200_DP(a73); B200_DP(a74); B200_DP(a75); B200_DP(a76);
        B200_DP(bt1); B200_DP(bt3); B200_DP(bt4); B200_DP(bt5); B200_DP(bt6); B200_DP(bt7);
#undef B200_DP
        real tmp[B200_N], k2[B200_N];
        const real a = dt * a21;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) tmp[i] = b200_fma(a, k1[i], uprev[i]);
        B200_RHS(k2, tmp, p, b200_fma(c1, dt, t));
#pragma unroll
        for (int i = 0; i < B200_N; ++i) tmp[i] = b200_fma(dt, b200_fma(a32, k2[i], a31 * k1[i]), uprev[i]);
        B200_RHS(k3, tmp, p, b200_fma(c2, dt, t));
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            tmp[i] = b200_fma(dt, b200_fma(a43, k3[i], b200_fma(a42, k2[i], a41 * k1[i])), uprev[i]);
        B200_RHS(k4, tmp, p, b200_fma(c3, dt, t));
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            tmp[i] = b200_fma(dt, b200_fma(a54, k4[i], b200_fma(a53, k3[i], b200_fma(a52, k2[i], a51 * k1[i]))),
                              uprev[i]);
        B200_RHS(k5, tmp, p, b200_fma(c4, dt, t));
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            tmp[i] = b200_fma(dt,
                              b200_fma(a65, k5[i],
                                       b200_fma(a64, k4[i], b200_fma(a63, k3[i], b200_fma(a62, k2[i], a61 * k1[i])))),
                              uprev[i]);
        B200_RHS(k6, tmp, p, t + dt);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) {
            upd[i] = b200_fma(a76, k6[i], b200_fma(a75, k5[i], b200_fma(a74, k4[i], b200_fma(a73, k3[i], a71 * k1[i]))));
            u[i] = b200_fma(dt, upd[i], uprev[i]);
        }
        B200_RHS(k7, u, p, t + dt);
        nf += 6;
        real ut[B200_N];
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            ut[i] = dt * b200_fma(bt7, k7[i],
                                    b200_fma(bt6, k6[i],
                                             b200_fma(bt5, k5[i], b200_fma(bt4, k4[i], b200_fma(bt3, k3[i], bt1 * k1[i])))));
        return b200_residual_norm(ut, uprev, u, reltol, abstol);
    }

    B200_D void accept() {
#pragma unroll
        for (int i = 0; i < B200_N; ++i) k1[i] = k7[i];
    }

    // integrator.k[2..4] of the step just taken (the reference fills them in every perform_step!;
    // they are pure functions of the stages, so they are formed only when a row needs them).
    // Must run before accept() overwrites k1.
    B200_D void dense_prepare(const real*, const real*, const real*, real, real) {
        const real d1 = B200_DP5_C.d1, d3 = B200_DP5_C.d3, d4 = B200_DP5_C.d4, d5 = B200_DP5_C.d5,
                   d6 = B200_DP5_C.d6, d7 = B200_DP5_C.d7;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) {
            const real bspl = k1[i] - upd[i];
            dk2[i] = bspl;
            dk3[i] = (upd[i] - k7[i]) - bspl;
            dk4[i] = b200_fma(d7, k7[i], b200_fma(d6, k6[i], b200_fma(d5, k5[i], b200_fma(d4, k4[i], b200_fma(d3, k3[i], d1 * k1[i])))));
        }
    }

    B200_D void interp(real th, real dt, const real* y0, const real* /*y1*/, real* out) const {
        const real b10 = th;
        const real b20 = th * ((real)1 - th);
        const real b30 = th * b20;
        const real b40 = b20 * b20;
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            out[i] = b200_fma(dt, b200_fma(dk4[i], b40, b200_fma(dk3[i], b30, b200_fma(dk2[i], b20, upd[i] * b10))), y0[i]);
    }
};

// ---------------------------------------------------------------------------
struct B200BS3 {
    real k1[B200_N], k4[B200_N];      // integrator.k = [fsalfirst, fsallast]; rows are written before accept()

    static B200_D int order() { return 3; }
    static B200_D bool fsal() { return true; }
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1; }
    static B200_D real beta2() { return (real)(2.0 / (5.0 * 3)); }
    static B200_D real beta1() { return (real)(7.0 / (10.0 * 3)); }

    B200_D void init(const real* u, const real* p, real t, int& nf) {
        B200_RHS(k1, u, p, t);
        nf += 1;
    }

    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt,
                        real reltol, real abstol, int& nf) {
        const real a21 = (real)0.5, a32 = (real)0.75, a41 = (real)0.2222222222222222, a42 = (real)0.3333333333333333,
                   a43 = (real)0.4444444444444444, c1 = (real)0.5, c2 = (real)0.75;
        const real bt1 = (real)0.06944444444444445, bt2 = (real)-0.08333333333333333, bt3 = (real)-0.1111111111111111,
                   bt4 = (real)0.125;
        real tmp[B200_N], k2[B200_N], k3[B200_N];
        const real a1 = dt * a21;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) tmp[i] = b200_fma(a1, k1[i], uprev[i]);
        B200_RHS(k2, tmp, p, b200_fma(c1, dt, t));
        const real a2 = dt * a32;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) tmp[i] = b200_fma(a2, k2[i], uprev[i]);
        B200_RHS(k3, tmp, p, b200_fma(c2, dt, t));
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            u[i] = b200_fma(dt, b200_fma(a43, k3[i], b200_fma(a42, k2[i], a41 * k1[i])), uprev[i]);
        B200_RHS(k4, u, p, t + dt);
        nf += 3;
        real ut[B200_N];
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            ut[i] = dt * b200_fma(bt4, k4[i], b200_fma(bt3, k3[i], b200_fma(bt2, k2[i], bt1 * k1[i])));
        return b200_residual_norm(ut, uprev, u, reltol, abstol);
    }

    B200_D void accept() {
#pragma unroll
        for (int i = 0; i < B200_N; ++i) k1[i] = k4[i];
    }

    B200_D void dense_prepare(const real*, const real*, const real*, real, real) {}

    // hermite_interpolant:
    //   (1-Θ) y0 + Θ y1 + Θ(Θ-1) ((1-2Θ)(y1-y0) + (Θ-1) dt k[1] + Θ dt k[2])
    B200_D void interp(real th, real dt, const real* y0, const real* y1, real* out) const {
        const real omt = (real)1 - th;
        const real tm1 = th - (real)1;
        const real ttm1 = th * tm1;
        const real om2t = b200_fma((real)-2, th, (real)1);
        const real c1 = tm1 * dt, c2 = th * dt;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) {
            const real inner = b200_fma(c2, k4[i], b200_fma(c1, k1[i], om2t * (y1[i] - y0[i])));
            out[i] = b200_fma(ttm1, inner, b200_fma(th, y1[i], omt * y0[i]));
        }
    }
};
