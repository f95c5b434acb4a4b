// b200_rosenbrock.cuh — Rosenbrock23 and Rodas5P, one trajectory per thread: analytic
// Jacobian (user source), W = J - I/(dt*gamma), small dense solve in registers.
//
// Reference behaviour reproduced (out-of-place / SVector forms):
//   Rosenbrock23  perform_step!(…, ::Rosenbrock23ConstantCache)  lib/OrdinaryDiffEqRosenbrock/src/rosenbrock_perform_step.jl:249-332
//                 tableau d = 1/(2+√2), c32 = 6+√2                …/rosenbrock_tableaus.jl:6-10
//                 interpolant                                     …/rosenbrock_interpolants.jl:46-61
//   Rodas5P       perform_step!(…, ::RosenbrockCombinedConstantCache) …/rosenbrock_perform_step.jl:431-559
//                 tableau                                         …/rosenbrock_tableaus.jl:23-71
//                 interpolant (interp_order 3)                    …/rosenbrock_interpolants.jl:175-205
//   J, dT, W      calc_rosenbrock_differentiation / calc_W / calc_J / calc_tderivative
//                 lib/OrdinaryDiffEqDifferentiation/src/derivative_utils.jl:1050-1110,937-1002,304-355,235-265
//                 (has_jac / has_tgrad branches; every attempt gets a fresh J: the W-method
//                 reuse logic :46-171 returns (true,true) for max_jac_age = 1 and for non-W methods;
//                 stats: nw += 1 and njacs += 2 per attempt — calc_W's calc_J plus the
//                 `jac_reuse.cached_J = calc_J(...)` re-evaluation at :1078)
//   linear solve  W is a StaticWOperator (EXT SciMLOperators): for n <= 7 it stores inv(W) and
//                 `W \ v` is a mat-vec.  Restated for n = 3 with StaticArrays' 3x3 `inv`
//                 (cross-product form); other n use partial-pivot LU (documented deviation,
//                 DESIGN.md).  issuccess_W(::StaticWOperator) is always true (T11).
#pragma once
#include "b200_base.cuh"
#include "b200_tableaus_gen.cuh"

#ifndef B200_LINSOLVE_LU
#if B200_N == 3 || B200_N == 1
#define B200_LINSOLVE_LU 0
#else
#define B200_LINSOLVE_LU 1
#endif
#endif

// ---- W factor object: either the explicit inverse (n = 1, 3) or LU with partial pivoting
struct B200WFact {
#if !B200_LINSOLVE_LU
    real inv[B200_N * B200_N];     // row-major inverse
#else
    real lu[B200_N * B200_N];      // row-major, unit-lower L below the diagonal, U on and above
    int piv[B200_N];
#endif
    bool ok;

    // W (column-major n×n, as f.jac returns it) -> factor
    B200_D void factor(const real* W) {
        ok = true;
#if !B200_LINSOLVE_LU
#if B200_N == 1
        inv[0] = (real)1 / W[0];
#else
        // StaticArrays _inv(::Size{(3,3)}, A): x0,x1,x2 = columns; y0 = x1 × x2; d = x0·y0;
        // x0 /= d; y0 /= d; y1 = x2 × x0; y2 = x0 × x1; rows of A^-1 are y0, y1, y2.
        real x0[3] = {W[0], W[1], W[2]}, x1[3] = {W[3], W[4], W[5]}, x2[3] = {W[6], W[7], W[8]};
        real y0[3], y1[3], y2[3];
        y0[0] = x1[1] * x2[2] - x1[2] * x2[1];
        y0[1] = x1[2] * x2[0] - x1[0] * x2[2];
        y0[2] = x1[0] * x2[1] - x1[1] * x2[0];
        const real d = (x0[0] * y0[0] + x0[1] * y0[1]) + x0[2] * y0[2];
        // six quotients by the same d: one correctly rounded reciprocal + the exact residual correction each
        // (b200_div_const_fast0: zero dividends — structural zeros of J — are exact too; bit-identical to IEEE x / d).  The reciprocal is the flagged branch-free sequence, the
        // exponent tests of the six dividends and the window test of d OR into the same flag, and ONE cold block redoes
        // all seven with the plain operator (singular or badly scaled W) — seven branches of r1's form become one.
        {
            const real ad = b200_abs(d);
            bool bad = !(ad >= (real)1e-30 && ad <= (real)1e30);
            const real rd = b200_div_fast((real)1, d, bad);
            real qx[3], qy[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { qx[i] = b200_div_const_fast0(x0[i], d, rd, bad); qy[i] = b200_div_const_fast0(y0[i], d, rd, bad); }
            if (bad) {      // (unrolled: a rolled loop would index the register arrays dynamically)
#pragma unroll
                for (int i = 0; i < 3; ++i) { qx[i] = b200_div_cold(x0[i], d); qy[i] = b200_div_cold(y0[i], d); }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) { x0[i] = qx[i]; y0[i] = qy[i]; }
        }
        y1[0] = x2[1] * x0[2] - x2[2] * x0[1];
        y1[1] = x2[2] * x0[0] - x2[0] * x0[2];
        y1[2] = x2[0] * x0[1] - x2[1] * x0[0];
        y2[0] = x0[1] * x1[2] - x0[2] * x1[1];
        y2[1] = x0[2] * x1[0] - x0[0] * x1[2];
        y2[2] = x0[0] * x1[1] - x0[1] * x1[0];
#pragma unroll
        for (int j = 0; j < 3; ++j) { inv[0 * 3 + j] = y0[j]; inv[1 * 3 + j] = y1[j]; inv[2 * 3 + j] = y2[j]; }
#endif
#else
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
#pragma unroll
            for (int j = 0; j < B200_N; ++j) lu[i * B200_N + j] = W[i + B200_N * j];
        for (int k = 0; k < B200_N; ++k) {
            int pr = k; real best = b200_abs(lu[k * B200_N + k]);
            for (int i = k + 1; i < B200_N; ++i) {
                real v = b200_abs(lu[i * B200_N + k]);
                if (v > best) { best = v; pr = i; }
            }
            piv[k] = pr;
            if (pr != k)
                for (int j = 0; j < B200_N; ++j) {
                    real tswap = lu[k * B200_N + j]; lu[k * B200_N + j] = lu[pr * B200_N + j]; lu[pr * B200_N + j] = tswap;
                }
            const real pivot = lu[k * B200_N + k];
            if (pivot == (real)0) { ok = false; continue; }
            for (int i = k + 1; i < B200_N; ++i) {
                const real l = lu[i * B200_N + k] / pivot;
                lu[i * B200_N + k] = l;
                for (int j = k + 1; j < B200_N; ++j) lu[i * B200_N + j] = lu[i * B200_N + j] - l * lu[k * B200_N + j];
            }
        }
#endif
    }

    // x = W \ b
    B200_D void solve(const real* b, real* x) const {
#if !B200_LINSOLVE_LU
#pragma unroll
        for (int i = 0; i < B200_N; ++i) {
            real s = inv[i * B200_N + 0] * b[0];
#pragma unroll
            for (int j = 1; j < B200_N; ++j) s = s + inv[i * B200_N + j] * b[j];
            x[i] = s;
        }
#else
        real y[B200_N];
        for (int i = 0; i < B200_N; ++i) y[i] = b[i];
        for (int k = 0; k < B200_N; ++k) {
            const int pr = piv[k];
            if (pr != k) { real tswap = y[k]; y[k] = y[pr]; y[pr] = tswap; }
            for (int i = k + 1; i < B200_N; ++i) y[i] = y[i] - lu[i * B200_N + k] * y[k];
        }
        for (int i = B200_N - 1; i >= 0; --i) {
            real s = y[i];
            for (int j = i + 1; j < B200_N; ++j) s = s - lu[i * B200_N + j] * y[j];
            y[i] = s / lu[i * B200_N + i];
        }
        for (int i = 0; i < B200_N; ++i) x[i] = y[i];
#endif
    }
};

// J, dT at (uprev, t); W = J - I * inv(dtgamma)
// opnorm_out (composite algorithms only): receives opnorm(J, Inf) = the largest row sum of |J| (rows summed left to right),
// which calc_W stores in integrator.eigen_est when the algorithm is a CompositeAlgorithm (derivative_utils.jl:996-999)
// lam_in: inv(dtgamma) when the caller has it already (Rosenbrock23 needs the same quotient for its stages)
B200_D void b200_build_W(const real* uprev, const real* p, real t, real dtgamma, real* dT, B200WFact& F,
                         int& njacs, int& nw, real* opnorm_out = nullptr, const real* lam_in = nullptr) {
    real J[B200_N * B200_N];
    B200_JAC(J, uprev, p, t);
    if (opnorm_out != nullptr) {
        real nrm = (real)0;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) {
            real row = b200_abs(J[i]);
#pragma unroll
            for (int j = 1; j < B200_N; ++j) row = row + b200_abs(J[i + B200_N * j]);
            nrm = (i == 0) ? row : b200_max(nrm, row);
        }
        *opnorm_out = nrm;
    }
#ifdef B200_TGRAD
    B200_TGRAD(dT, uprev, p, t);
#else
#pragma unroll
    for (int i = 0; i < B200_N; ++i) dT[i] = (real)0;     // autonomous system
#endif
    njacs += 2;
    nw += 1;
    real lam;
    if (lam_in != nullptr) lam = *lam_in;
    else {   // inv(dtgamma): flagged branch-free IEEE quotient, the plain operator in the (cold) flagged case
        bool bad = false;
        lam = b200_div_fast((real)1, dtgamma, bad);
        if (bad) lam = b200_div_cold((real)1, dtgamma);
    }
#pragma unroll
    for (int i = 0; i < B200_N; ++i) J[i + B200_N * i] = J[i + B200_N * i] - lam;
    F.factor(J);
}

B200_D real b200_err_norm(const real* ut, const real* uprev, const real* u, real reltol, real abstol) {
    return b200_residual_norm(ut, uprev, u, reltol, abstol);
}

// ---------------------------------------------------------------------------
// Rosenbrock23 and Rosenbrock32 share the three stages (rosenbrock_perform_step.jl:249-332 / :333-417).
// Rosenbrock32 (B200_ROS_THIRD) advances with the third-order combination u = uprev + dt/6 (k1 + 4 k2 + k3), keeps
// f(uprev + dt k2) as fsallast, and forms k1 as (W \ -(fsalfirst + dtγ dT)) / dtγ.
#define B200_ROS_THIRD (B200_ALG == B200_ALG_ROS32)
struct B200Ros23 {
    real k1[B200_N], k2[B200_N];       // dense output rows (integrator.k[1], k[2])
    real f0[B200_N], f2[B200_N];       // fsalfirst, fsallast

    static B200_D int order() { return B200_ROS_THIRD ? 3 : 2; }
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1.2; }     // 6//5 for implicit methods (alg_utils.jl:857)

    B200_D void init(const real* u, const real* p, real t, int& nf) {
        B200_RHS(f0, u, p, t);
        nf += 1;
    }

    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt, real reltol, real abstol,
                        int& nf, int& njacs, int& nw, int& nsolve, bool /*calck*/, real* opnorm_out = nullptr) {
        const real d = (real)0.2928932188134525;       // convert(T, 1/(2+sqrt(2)))
        const real c32 = (real)7.414213562373095;      // convert(T, 6+sqrt(2))
        const real dtg = dt * d;
        // -inv(dtγ), dt/2, dt/6: the two true quotients as one flagged group (dt/6 through the correctly rounded 1/6)
        real lam, dto6;
        {
            bool bad = false;
            lam = b200_div_fast((real)1, dtg, bad);
            dto6 = b200_div_const_fast(dt, (real)6, (real)1 / (real)6, bad);
            if (bad) { lam = b200_div_cold((real)1, dtg); dto6 = b200_div_cold(dt, (real)6); }
        }
        const real ninv = -lam;
        const real dto2 = dt / (real)2;
        real dT[B200_N], rhs[B200_N], tmp[B200_N], f1[B200_N], k3[B200_N];
        B200WFact F;
        b200_build_W(uprev, p, t, dtg, dT, F, njacs, nw, opnorm_out, &lam);
        if (!F.ok) return (real)2;
#if B200_ROS_THIRD
#pragma unroll
        for (int i = 0; i < B200_N; ++i) rhs[i] = -b200_fma(dtg, dT[i], f0[i]);
        F.solve(rhs, k1);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) k1[i] = k1[i] / dtg;
#else
#pragma unroll
        for (int i = 0; i < B200_N; ++i) rhs[i] = b200_fma(dtg, dT[i], f0[i]);
        F.solve(rhs, k1);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) k1[i] = k1[i] * ninv;
#endif
        nsolve += 1;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) tmp[i] = b200_fma(dto2, k1[i], uprev[i]);
        B200_RHS(f1, tmp, p, t + dto2);
        nf += 1;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) rhs[i] = f1[i] - k1[i];
        F.solve(rhs, k2);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) k2[i] = b200_fma(k2[i], ninv, k1[i]);
        nsolve += 1;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) u[i] = b200_fma(dt, k2[i], uprev[i]);
        B200_RHS(f2, u, p, t + dt);
        nf += 1;
#pragma unroll
        for (int i = 0; i < B200_N; ++i)
            rhs[i] = b200_fma(dt, dT[i],
                              b200_fma((real)-2, k1[i] - f0[i], b200_fma(-c32, k2[i] - f1[i], f2[i])));
        F.solve(rhs, k3);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) k3[i] = k3[i] * ninv;
        nsolve += 1;
#if B200_ROS_THIRD
        // (u held uprev + dt k2 for fsallast) u = uprev + dto6*(k1 + 4k2 + k3) = muladd(dto6, muladd(4, k2, k1 + k3), uprev)
#pragma unroll
        for (int i = 0; i < B200_N; ++i) u[i] = b200_fma(dto6, b200_fma((real)4, k2[i], k1[i] + k3[i]), uprev[i]);
#endif
#pragma unroll
        for (int i = 0; i < B200_N; ++i) tmp[i] = dto6 * (b200_fma((real)-2, k2[i], k1[i]) + k3[i]);
        return b200_err_norm(tmp, uprev, u, reltol, abstol);
    }

    B200_D void accept() {
#pragma unroll
        for (int i = 0; i < B200_N; ++i) f0[i] = f2[i];
    }
    B200_D void dense_prepare(const real*, const real*, const real*, real, real) {}

    B200_D void interp(real th, real dt, const real* y0, const real* /*y1*/, real* out) const {
        const real d = (real)0.2928932188134525;
        const real den = b200_fma((real)-2, d, (real)1);          // 1 - 2d
        const real c1 = th * ((real)1 - th) / den;
        const real c2 = th * b200_fma((real)-2, d, th) / den;     // Θ(Θ - 2d)/(1 - 2d)
#pragma unroll
        for (int i = 0; i < B200_N; ++i) out[i] = b200_fma(dt, b200_fma(c2, k2[i], c1 * k1[i]), y0[i]);
    }
};

// ---------------------------------------------------------------------------
// Tableau selection for the generic Rodas-type stepper (RodasTableau family; b = [A[S,1:S-1]; 1],
// btilde = e_S; rosenbrock_tableaus.jl of OrdinaryDiffEqRosenbrock / OrdinaryDiffEqRosenbrockTableaus).
// Rodas3P (lib/OrdinaryDiffEqRosenbrockTableaus/src/rosenbrock_tableaus.jl:386-445), written as the reference writes it
// (4.0 / 3.0 etc. are Float64 quotients, then convert(T, .)): 5 stages, explicit b / btilde, stage 5 repeats stage 4's
// (c, A row) so its f evaluation is skipped (rosenbrock_perform_step.jl:474-481), three rows of H of which the
// interpolant uses two (interp_order = 2, rosenbrock_caches.jl:487-488).
#define B200_RODAS3P_S 5
#define B200_RODAS3P_HR 3
#define B200_RODAS3P_GAMMA (1.0 / 3.0)
#define B200_RODAS3P_A { {0, 0, 0, 0, 0}, {4.0 / 3.0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {2.90625, 3.375, 0.40625, 0, 0}, {2.90625, 3.375, 0.40625, 0, 0} }
#define B200_RODAS3P_C { {0, 0, 0, 0}, {-4.0, 0, 0, 0}, {8.25, 6.75, 0, 0}, {1.21875, -5.0625, -1.96875, 0}, {4.03125, -15.1875, -4.03125, 6.0} }
#define B200_RODAS3P_c {0, 4.0 / 9.0, 0, 1, 1}
#define B200_RODAS3P_d {1.0 / 3.0, -(1.0 / 9.0), 1.0, 0, 0}
#define B200_RODAS3P_H { {1.78125, 6.75, 0.15625, -6.0, -1.0}, {4.21875, -15.1875, -3.09375, 9.0, 0}, {4.21875, -2.025, -1.63125, -1.7, -0.1} }
#define B200_RODAS3P_B {2.90625, 3.375, 0.40625, 0, 1}
#define B200_RODAS3P_BTILDE {0, 0, 0, -1, 1}

// Rodas23W (lib/OrdinaryDiffEqRosenbrock/src/rosenbrock_tableaus.jl:179-236): Rodas3P's stages, the weights of the
// second-order solution, H rows (h2, 0, h2).  A W-method whose Jacobian reuse never engages for SVector states
// (`cache.W isa AbstractSciMLOperator` holds for a StaticWOperator: _rosenbrock_jac_reuse_decision returns (true, true)).
#define B200_RODAS23W_S 5
#define B200_RODAS23W_HR 3
#define B200_RODAS23W_GAMMA B200_RODAS3P_GAMMA
#define B200_RODAS23W_A B200_RODAS3P_A
#define B200_RODAS23W_C B200_RODAS3P_C
#define B200_RODAS23W_c B200_RODAS3P_c
#define B200_RODAS23W_d B200_RODAS3P_d
#define B200_RODAS23W_H { {4.21875, -2.025, -1.63125, -1.7, -0.1}, {0, 0, 0, 0, 0}, {4.21875, -2.025, -1.63125, -1.7, -0.1} }
#define B200_RODAS23W_B {2.90625, 3.375, 0.40625, 1, 0}
#define B200_RODAS23W_BTILDE {0, 0, 0, 1, -1}

#if B200_ALG == B200_ALG_RODAS23W
#define B200_RODAS_NAME(x) B200_RODAS23W_##x
#define B200_RODAS_ORDER 3
#define B200_RODAS_EXPLICIT_B 1
#define B200_RODAS_INTERP 2
#define B200_RODAS_FSKIP 4
#elif B200_ALG == B200_ALG_RODAS3P
#define B200_RODAS_NAME(x) B200_RODAS3P_##x
#define B200_RODAS_ORDER 3
#define B200_RODAS_EXPLICIT_B 1         // b and btilde are stored vectors (zero entries skipped like the reference)
#define B200_RODAS_INTERP 2             // interp_order
#define B200_RODAS_FSKIP 4              // 0-based stage that reuses the previous stage's f (asserted on the host side too)
#elif B200_ALG == B200_ALG_RODAS5
#define B200_RODAS_NAME(x) B200_RODAS5_##x
#define B200_RODAS_ORDER 5
#elif B200_ALG == B200_ALG_RODAS4
#define B200_RODAS_NAME(x) B200_RODAS4_##x
#define B200_RODAS_ORDER 4
#elif B200_ALG == B200_ALG_RODAS42
#define B200_RODAS_NAME(x) B200_RODAS42_##x
#define B200_RODAS_ORDER 4
#elif B200_ALG == B200_ALG_RODAS4P
#define B200_RODAS_NAME(x) B200_RODAS4P_##x
#define B200_RODAS_ORDER 4
#elif B200_ALG == B200_ALG_RODAS4P2
#define B200_RODAS_NAME(x) B200_RODAS4P2_##x
#define B200_RODAS_ORDER 4
#else      // Rodas5P and Rodas5Pe (same tableau; Rodas5Pe brings a full vector of embedded error weights)
#define B200_RODAS_NAME(x) B200_RODAS5P_##x
#define B200_RODAS_ORDER 5
#endif
#define B200_RODAS_S B200_RODAS_NAME(S)
#define B200_RODAS_HR B200_RODAS_NAME(HR)
#ifndef B200_RODAS_EXPLICIT_B
#define B200_RODAS_EXPLICIT_B 0
#endif
#ifndef B200_RODAS_INTERP
#define B200_RODAS_INTERP B200_RODAS_HR
#endif
#ifndef B200_RODAS_FSKIP
#define B200_RODAS_FSKIP (-1)
#endif

struct B200Rodas5PCoeffs {
    real A[B200_RODAS_S][B200_RODAS_S];
    real C[B200_RODAS_S][B200_RODAS_S - 1];
    real c[B200_RODAS_S];
    real d[B200_RODAS_S];
    real H[B200_RODAS_HR][B200_RODAS_S];
    real gamma;
#if B200_ALG == B200_ALG_RODAS5PE
    real btilde[B200_RODAS_S];
#endif
#if B200_RODAS_EXPLICIT_B
    real b[B200_RODAS_S];
    real btilde[B200_RODAS_S];
#endif
};
__constant__ B200Rodas5PCoeffs B200_RODAS5P_TAB = {B200_RODAS_NAME(A), B200_RODAS_NAME(C), B200_RODAS_NAME(c), B200_RODAS_NAME(d),
                                                 B200_RODAS_NAME(H), B200_RODAS_NAME(GAMMA)
#if B200_ALG == B200_ALG_RODAS5PE
                                                 , B200_RODAS5PE_BTILDE
#endif
#if B200_RODAS_EXPLICIT_B
                                                 , B200_RODAS_NAME(B), B200_RODAS_NAME(BTILDE)
#endif
};

struct B200Rodas5P {
    real dense[B200_RODAS_HR][B200_N];   // integrator.k[1..size(H,1)] (only filled when rows are interpolated)

    static B200_D int order() { return B200_RODAS_ORDER; }
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1.2; }

    B200_D void init(const real*, const real*, real, int&) {}    // not FSAL (alg_utils.jl:60)

    // the S stages of one attempt (rosenbrock_perform_step.jl:431-559, RodasTableau form).  EXACT: dtC entries by the
    // plain division (dt outside the window in which the reciprocal form is the IEEE quotient)
    template <bool EXACT>
    B200_D void stages(const real* uprev, const real* p, real t, real dt, real rdt, const real* dT, const B200WFact& F,
                       real (*ks)[B200_N], int& nf, int& nsolve) {
        const B200Rodas5PCoeffs& T = B200_RODAS5P_TAB;
        real du[B200_N], lt[B200_N], us[B200_N];
        B200_RHS(du, uprev, p, t);
        nf += 1;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) lt[i] = -b200_fma(dt * T.d[0], dT[i], du[i]);
        F.solve(lt, ks[0]);
#pragma unroll
        for (int s = 1; s < B200_RODAS_S; ++s) {
#pragma unroll
            for (int i = 0; i < B200_N; ++i) us[i] = uprev[i];
#pragma unroll
            for (int j = 0; j < s; ++j)
#pragma unroll
                for (int i = 0; i < B200_N; ++i) us[i] = b200_fma(T.A[s][j], ks[j][i], us[i]);
            if (s != B200_RODAS_FSKIP) {        // a stage that repeats the previous (c, A row) keeps du (:474-481)
                B200_RHS(du, us, p, b200_fma(T.c[s], dt, t));
                nf += 1;
            }
#pragma unroll
            for (int i = 0; i < B200_N; ++i) lt[i] = (real)0;
#pragma unroll
            for (int j = 0; j < s; ++j) {
                real q;
                if (EXACT) q = b200_div_cold(T.C[s][j], dt);
                else {
                    const real q0 = T.C[s][j] * rdt;
                    q = b200_fma(b200_fma(-dt, q0, T.C[s][j]), rdt, q0);
                }
#pragma unroll
                for (int i = 0; i < B200_N; ++i) lt[i] = b200_fma(q, ks[j][i], lt[i]);
            }
            const real dtd = dt * T.d[s];
#pragma unroll
            for (int i = 0; i < B200_N; ++i) lt[i] = -b200_fma(dtd, dT[i], du[i] + lt[i]);
            F.solve(lt, ks[s]);
            nsolve += 1;
        }
    }

    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt, real reltol, real abstol,
                        int& nf, int& njacs, int& nw, int& nsolve, bool calck) {
        const B200Rodas5PCoeffs& T = B200_RODAS5P_TAB;
        const real dtgamma = dt * T.gamma;
        real dT[B200_N], du[B200_N], lt[B200_N], us[B200_N];
        real ks[B200_RODAS_S][B200_N];
        B200WFact F;
        b200_build_W(uprev, p, t, dtgamma, dT, F, njacs, nw);
        if (!F.ok) return (real)2;
        // dtC = C ./ dt : all quotients share the divisor, so one correctly rounded reciprocal + the exact residual
        // correction gives each correctly rounded quotient (same bits as IEEE division; see b200_div_const).  The
        // reciprocal is the flagged branch-free sequence; its flag and the window test of dt are taken ONCE, and the
        // (cold) flagged case runs a second copy of the stage loop that divides with the plain operator — r1 tested the
        // window inside every unrolled (s, j) term, 28 branches per attempt.
        bool bad = !(b200_abs(dt) >= (real)1e-30 && b200_abs(dt) <= (real)1e30);
        const real rdt = b200_div_fast((real)1, dt, bad);
        if (bad) stages<true>(uprev, p, t, dt, rdt, dT, F, ks, nf, nsolve);
        else stages<false>(uprev, p, t, dt, rdt, dT, F, ks, nf, nsolve);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) u[i] = uprev[i];
#pragma unroll
        for (int j = 0; j < B200_RODAS_S; ++j) {
#if B200_RODAS_EXPLICIT_B
            const real b = T.b[j];
            if (b == (real)0) continue;
#else
            const real b = (j < B200_RODAS_S - 1) ? T.A[B200_RODAS_S - 1][j] : (real)1;      // b = [A[S,1:S-1]; 1]
            if (j < B200_RODAS_S - 1 && T.A[B200_RODAS_S - 1][j] == (real)0) continue;
#endif
#pragma unroll
            for (int i = 0; i < B200_N; ++i) u[i] = b200_fma(b, ks[j][i], u[i]);
        }
#if B200_RODAS_EXPLICIT_B
        // du = zero(uprev); for i: if !iszero(btilde[i]) du = du + btilde[i]*ks[i]
#pragma unroll
        for (int i = 0; i < B200_N; ++i) du[i] = (real)0;
#pragma unroll
        for (int j = 0; j < B200_RODAS_S; ++j) {
            if (T.btilde[j] == (real)0) continue;
#pragma unroll
            for (int i = 0; i < B200_N; ++i) du[i] = b200_fma(T.btilde[j], ks[j][i], du[i]);
        }
#elif B200_ALG == B200_ALG_RODAS5PE
        // du = zero(uprev); for i: du = du + btilde[i]*ks[i] (no entry of Rodas5Pe's btilde is zero)
#pragma unroll
        for (int i = 0; i < B200_N; ++i) du[i] = (real)0;
#pragma unroll
        for (int j = 0; j < B200_RODAS_S; ++j)
#pragma unroll
            for (int i = 0; i < B200_N; ++i) du[i] = b200_fma(T.btilde[j], ks[j][i], du[i]);
#else
        // btilde = e_S: du = 0 + 1*ks[S]
#pragma unroll
        for (int i = 0; i < B200_N; ++i) du[i] = b200_fma((real)1, ks[B200_RODAS_S - 1][i], (real)0);
#endif
        const real EEst = b200_err_norm(du, uprev, u, reltol, abstol);
        if (calck) {
#pragma unroll
            for (int r = 0; r < B200_RODAS_HR; ++r)
#pragma unroll
                for (int i = 0; i < B200_N; ++i) dense[r][i] = (real)0;
#pragma unroll
            for (int j = 0; j < B200_RODAS_S; ++j)
#pragma unroll
                for (int r = 0; r < B200_RODAS_HR; ++r)
#pragma unroll
                    for (int i = 0; i < B200_N; ++i) dense[r][i] = b200_fma(T.H[r][j], ks[j][i], dense[r][i]);
        }
        return EEst;
    }

    B200_D void accept() {}
    B200_D void dense_prepare(const real*, const real*, const real*, real, real) {}

    // interp_order = size(H,1) (rosenbrock_interpolants.jl:172-206):
    //   3: Θ1*y0 + Θ*(y1 + Θ1*(k1 + Θ*(k2 + Θ*k3)))     2: Θ1*y0 + Θ*(y1 + Θ1*(k1 + Θ*k2))
    B200_D void interp(real th, real /*dt*/, const real* y0, const real* y1, real* out) const {
        const real th1 = (real)1 - th;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) {
#if B200_RODAS_INTERP == 3
            real in = b200_fma(th, dense[2][i], dense[1][i]);
            in = b200_fma(th, in, dense[0][i]);
#else
            real in = b200_fma(th, dense[1][i], dense[0][i]);
#endif
            in = b200_fma(th1, in, y1[i]);
            out[i] = b200_fma(th, in, th1 * y0[i]);
        }
    }
};
