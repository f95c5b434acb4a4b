// b200_tsit5.cuh — Tsit5 stage loop, embedded error estimate and free 4th-order
// interpolant, one trajectory per thread, all stage vectors in registers.
//
// Reference behaviour reproduced (arithmetic order and fusion included):
//   perform_step!(…, ::Tsit5ConstantCache)  lib/OrdinaryDiffEqTsit5/src/tsit_perform_step.jl:140-186
//   initialize!                              …/tsit_perform_step.jl:125-138
//   tableau (Float64 literals, convert(T,·)) …/tsit_tableaus.jl:52-92
//   interpolant b_i(Θ), y0 + dt*Σ k_i b_i    …/interpolants.jl:32-57, coefficients tsit_tableaus.jl:244-276
// Fusion follows MuladdMacro: in a sum the last product is the outermost muladd
// (SURVEY §8 T2), `dt * (Σ)` of the error estimate is a plain multiply.
#pragma once
#include "b200_base.cuh"

#define B200_TSIT5_ORDER 5

// Tableau in constant memory: DFMA/FFMA take a constant-bank operand directly, so a
// coefficient costs no instruction (immediates would cost two UMOVs per double).
struct B200Tsit5Coeffs {
    real c1, c2, c3, c4;
    real a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65, a71, a72, a73, a74, a75, a76;
    real bt1, bt2, bt3, bt4, bt5, bt6, bt7;
    real r11, r12, r13, r14, r22, r23, r24, r32, r33, r34, r42, r43, r44, r52, r53, r54, r62, r63, r64, r72, r73, r74;
};
__constant__ B200Tsit5Coeffs B200_TSIT5_C = {
    (real)0.161, (real)0.327, (real)0.9, (real)0.9800255409045097,
    (real)0.161, (real)-0.008480655492356989, (real)0.335480655492357,
    (real)2.8971530571054935, (real)-6.359448489975075, (real)4.3622954328695815,
    (real)5.325864828439257, (real)-11.748883564062828, (real)7.4955393428898365, (real)-0.09249506636175525,
    (real)5.86145544294642, (real)-12.92096931784711, (real)8.159367898576159, (real)-0.071584973281401,
    (real)-0.028269050394068383,
    (real)0.09646076681806523, (real)0.01, (real)0.4798896504144996, (real)1.379008574103742,
    (real)-3.290069515436081, (real)2.324710524099774,
    (real)-0.00178001105222577714, (real)-0.0008164344596567469, (real)0.007880878010261995,
    (real)-0.1447110071732629, (real)0.5823571654525552, (real)-0.45808210592918697, (real)0.015151515151515152,
    (real)1.0, (real)-2.763706197274826, (real)2.9132554618219126, (real)-1.0530884977290216,
    (real)0.13169999999999998, (real)-0.2234, (real)0.1017,
    (real)3.9302962368947516, (real)-5.941033872131505, (real)2.490627285651253,
    (real)-12.411077166933676, (real)30.33818863028232, (real)-16.548102889244902,
    (real)37.50931341651104, (real)-88.1789048947664, (real)47.37952196281928,
    (real)-27.896526289197286, (real)65.09189467479366, (real)-34.87065786149661,
    (real)1.5, (real)-4.0, (real)2.5,
};


// the seven interpolation weights b_j(Theta) of _ode_interpolant(..., ::Tsit5ConstantCache, ..., Val{0}) — the same
// expressions as B200Tsit5::interp below (used by the staged saveat queue, which reads the stages from shared memory)
B200_D void b200_tsit5_interp_weights(real th, real* b) {
#define B200_T5(name) const real name = B200_TSIT5_C.name
    B200_T5(r11); B200_T5(r12); B200_T5(r13); B200_T5(r14); B200_T5(r22); B200_T5(r23); B200_T5(r24);
    B200_T5(r32); B200_T5(r33); B200_T5(r34); B200_T5(r42); B200_T5(r43); B200_T5(r44);
    B200_T5(r52); B200_T5(r53); B200_T5(r54); B200_T5(r62); B200_T5(r63); B200_T5(r64);
    B200_T5(r72); B200_T5(r73); B200_T5(r74);
#undef B200_T5
    const real th2 = th * th;
    b[0] = th * b200_fma(th, b200_fma(th, b200_fma(th, r14, r13), r12), r11);
    b[1] = th2 * b200_fma(th, b200_fma(th, r24, r23), r22);
    b[2] = th2 * b200_fma(th, b200_fma(th, r34, r33), r32);
    b[3] = th2 * b200_fma(th, b200_fma(th, r44, r43), r42);
    b[4] = th2 * b200_fma(th, b200_fma(th, r54, r53), r52);
    b[5] = th2 * b200_fma(th, b200_fma(th, r64, r63), r62);
    b[6] = th2 * b200_fma(th, b200_fma(th, r74, r73), r72);
}

struct B200Tsit5 {
    real k1[B200_N], k2[B200_N], k3[B200_N], k4[B200_N], k5[B200_N], k6[B200_N], k7[B200_N];

    static B200_D int order() { return 5; }
    static B200_D bool fsal() { return true; }
    // qsteady_min/max defaults for explicit methods (alg_utils.jl:833,853)
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1; }

    // initialize!: fsalfirst = f(uprev, p, t); nf += 1
    B200_D void init(const real* u, const real* p, real t, int& nf) {
        B200_RHS(k1, u, p, t);
        nf += 1;
    }

    // one attempted step; returns EEst
    // g6out (optional): receives the stage state of k6, which the composite algorithm's stiffness estimate reads
    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt,
                        real reltol, real abstol, int& nf, real* g6out = nullptr) {
        const real c1 = B200_TSIT5_C.c1, c2 = B200_TSIT5_C.c2, c3 = B200_TSIT5_C.c3, c4 = B200_TSIT5_C.c4;
#define B200_T5(name) const real name = B200_TSIT5_C.name
        B200_T5(a21); B200_T5(a31); B200_T5(a32); B200_T5(a41); B200_T5(a42); B200_T5(a43);
        B200_T5(a51); B200_T5(a52); B200_T5(a53); B200_T5(a54);
        B200_T5(a61); B200_T5(a62); B200_T5(a63); B200_T5(a64); B200_T5(a65);
        B200_T5(a71); B200_T5(a72); B200_T5(a73); B200_T5(a74); B200_T5(a75); B200_T5(a76);
        B200_T5(bt1); B200_T5(bt2); B200_T5(bt3); B200_T5(bt4); B200_T5(bt5); B200_T5(bt6); B200_T5(bt7);
#undef B200_T5
        real tmp[B200_N];
        const real a = dt * a21;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) tmp[i] = b200_fma(a, k1[i], uprev[i]);
        B200_RHS(k2, tmp, p, b200_fma(c1, dt, t));
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            tmp[i] = b200_fma(dt, b200_fma(a32, k2[i], a31 * k1[i]), uprev[i]);
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
        if (g6out != nullptr) {
#pragma unroll
            for (int i = 0; i < B200_N; ++i) g6out[i] = tmp[i];
        }
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            u[i] = b200_fma(dt,
                            b200_fma(a76, k6[i],
                                     b200_fma(a75, k5[i],
                                              b200_fma(a74, k4[i],
                                                       b200_fma(a73, k3[i], b200_fma(a72, k2[i], a71 * k1[i]))))),
                            uprev[i]);
        B200_RHS(k7, u, p, t + dt);
        nf += 6;
        // utilde = dt*(Σ btilde_i k_i); atmp = calculate_residuals; EEst = internalnorm(atmp)
        real ut[B200_N];
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            ut[i] = dt * b200_fma(bt7, k7[i],
                                  b200_fma(bt6, k6[i],
                                           b200_fma(bt5, k5[i],
                                                    b200_fma(bt4, k4[i],
                                                             b200_fma(bt3, k3[i],
                                                                      b200_fma(bt2, k2[i], bt1 * k1[i]))))));
        return b200_residual_norm(ut, uprev, u, reltol, abstol);
    }

    // apply_step!/update_fsal!: fsalfirst = fsallast
    B200_D void accept() {
#pragma unroll
        for (int i = 0; i < B200_N; ++i) k1[i] = k7[i];
    }

    // reset_fsal! (integrator_utils.jl:1325-1343): fsalfirst = f(u, p, t); nf += 1 (after a callback modified u)
    B200_D void reset_fsal(const real* u, const real* p, real t, int& nf) {
        B200_RHS(k1, u, p, t);
        nf += 1;
    }

    // _ode_addsteps!(k, t, uprev, u, dt, f, p, ::Tsit5ConstantCache, always_calc_begin = true) (tsit_perform_step.jl:40-82):
    // all seven stages again from uprev with the (shortened) dt, after change_t_via_interpolation! moved t to an event.
    // Note `uprev + dt*(a21*k1)` here against perform_step!'s `a = dt*a21; uprev + a*k1`; stats are not touched.
    B200_D void addsteps_always(const real* uprev, const real* p, real t, real dt) {
        const real c1 = B200_TSIT5_C.c1, c2 = B200_TSIT5_C.c2, c3 = B200_TSIT5_C.c3, c4 = B200_TSIT5_C.c4;
#define B200_T5(name) const real name = B200_TSIT5_C.name
        B200_T5(a21); B200_T5(a31); B200_T5(a32); B200_T5(a41); B200_T5(a42); B200_T5(a43);
        B200_T5(a51); B200_T5(a52); B200_T5(a53); B200_T5(a54);
        B200_T5(a61); B200_T5(a62); B200_T5(a63); B200_T5(a64); B200_T5(a65);
        B200_T5(a71); B200_T5(a72); B200_T5(a73); B200_T5(a74); B200_T5(a75); B200_T5(a76);
#undef B200_T5
        real tmp[B200_N];
        B200_RHS(k1, uprev, p, t);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) tmp[i] = b200_fma(dt, a21 * k1[i], uprev[i]);
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
            tmp[i] = b200_fma(dt, b200_fma(a54, k4[i], b200_fma(a53, k3[i], b200_fma(a52, k2[i], a51 * k1[i]))), uprev[i]);
        B200_RHS(k5, tmp, p, b200_fma(c4, dt, t));
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            tmp[i] = b200_fma(dt, b200_fma(a65, k5[i], b200_fma(a64, k4[i], b200_fma(a63, k3[i], b200_fma(a62, k2[i], a61 * k1[i])))), uprev[i]);
        B200_RHS(k6, tmp, p, t + dt);
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            tmp[i] = b200_fma(dt, b200_fma(a76, k6[i], b200_fma(a75, k5[i], b200_fma(a74, k4[i], b200_fma(a73, k3[i], b200_fma(a72, k2[i], a71 * k1[i]))))), uprev[i]);
        B200_RHS(k7, tmp, p, t + dt);
    }

    // nothing to prepare: all 7 stages are kept (ode_addsteps! is a no-op once length(k) >= 7)
    B200_D void dense_prepare(const real*, const real*, const real*, real, real) {}

    // _ode_interpolant(Θ, dt, y0, y1, k, ::Tsit5ConstantCache, nothing, Val{0})
    B200_D void interp(real th, real dt, const real* y0, const real* /*y1*/, real* out) const {
#define B200_T5(name) const real name = B200_TSIT5_C.name
        B200_T5(r11); B200_T5(r12); B200_T5(r13); B200_T5(r14); B200_T5(r22); B200_T5(r23); B200_T5(r24);
        B200_T5(r32); B200_T5(r33); B200_T5(r34); B200_T5(r42); B200_T5(r43); B200_T5(r44);
        B200_T5(r52); B200_T5(r53); B200_T5(r54); B200_T5(r62); B200_T5(r63); B200_T5(r64);
        B200_T5(r72); B200_T5(r73); B200_T5(r74);
#undef B200_T5
        const real th2 = th * th;
        const real b1 = th * b200_fma(th, b200_fma(th, b200_fma(th, r14, r13), r12), r11);
        const real b2 = th2 * b200_fma(th, b200_fma(th, r24, r23), r22);
        const real b3 = th2 * b200_fma(th, b200_fma(th, r34, r33), r32);
        const real b4 = th2 * b200_fma(th, b200_fma(th, r44, r43), r42);
        const real b5 = th2 * b200_fma(th, b200_fma(th, r54, r53), r52);
        const real b6 = th2 * b200_fma(th, b200_fma(th, r64, r63), r62);
        const real b7 = th2 * b200_fma(th, b200_fma(th, r74, r73), r72);
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            out[i] = b200_fma(dt,
                              b200_fma(k7[i], b7,
                                       b200_fma(k6[i], b6,
                                                b200_fma(k5[i], b5,
                                                         b200_fma(k4[i], b4,
                                                                  b200_fma(k3[i], b3,
                                                                           b200_fma(k2[i], b2, k1[i] * b1)))))),
                              y0[i]);
    }
};
