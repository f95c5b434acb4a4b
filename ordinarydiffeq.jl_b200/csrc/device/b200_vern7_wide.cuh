// b200_vern7_wide.cuh — Vern7 for WIDE states (the register-pressure path, e.g. Pleiades n = 28):
// one trajectory per thread, the stage derivatives k1..k10 of every thread in SHARED memory.
//
// Why: 10 stage vectors x 28 doubles do not fit a thread's 255 registers.  The plain one-thread kernel lets
// ptxas put them in local memory (r1: 13 GB of DRAM write-back per 64 Ki trajectories, FP64 pipe 22 % busy); the
// lane-group kernel (b200_coop.cuh) spreads them over 16 lanes but then evaluates every pair force of the RHS
// ~2.4 times.  Here the stage vectors live in the SM's shared memory — 9 slots x n reals per thread (k2 and k10 share
// a slot: k2 is dead after stage 3) laid out [slot][component][thread], so a warp's access to one component is
// one conflict-free row — and only the stage STATE (uprev, the stage argument, the RHS accumulators) is in
// registers.  A 227 KB SM holds 9 x 28 x 8 B = 2016 B per thread for 112 threads: warps 0..2 full, warp 3 half
// (the kernel is launched with 128 threads, the last 16 never take a trajectory).  That is one warp per scheduler,
// so the instruction stream has to carry its own parallelism:
//   * the stages run in ONE loop with the a-matrix as a __constant__ table, so the user's RHS is inlined exactly
//     once, as straight-line code on registers (21 x 4 independent pair forces for Pleiades) — the 16 inlined
//     copies of the plain kernel (470 KB of SASS) thrashed the instruction cache;
//   * the RHS is compiled with sqrt / B200_DIV(a, b) mapped to the flagged branch-free IEEE sequences of
//     b200_base.cuh (one basic block: independent roots and quotients overlap, and quotients with one divisor
//     share the reciprocal refinement through ordinary common-subexpression elimination); the plain operators
//     re-evaluate the stage only if the flag was raised (cold, out of line).
// Arithmetic, fma nesting and operation order are those of b200_vern7.cuh (perform_step!(…, ::Vern7ConstantCache),
// lib/OrdinaryDiffEqVerner/src/verner_rk_perform_step.jl:256-383): results are bit-identical to the plain kernel.
// Limitations of this variant: no interior saveat rows / dense output (the lazy stages k11..k16 would need six more
// slots); save_start / save_end rows, final states and all statistics are served.  The host rejects the rest.
#pragma once
#include "b200_base.cuh"
#include "b200_tableaus_gen.cuh"

#ifndef B200_WIDE_NT
#error "B200_WIDE_NT (threads per CTA that own a trajectory) must be defined for the shared-memory stage kernel"
#endif
#define B200_WIDE_SLOTS 9

extern __shared__ __align__(16) unsigned char b200_wide_smem_raw[];

struct B200V7Named {
#define B200_X(name, val) real name;
    B200_VERN7_TABLEAU(B200_X)
#undef B200_X
};
constexpr B200V7Named b200_v7n = {
#define B200_X(name, val) (real)val,
    B200_VERN7_TABLEAU(B200_X)
#undef B200_X
};
__constant__ B200V7Named B200_V7W_C = {
#define B200_X(name, val) (real)val,
    B200_VERN7_TABLEAU(B200_X)
#undef B200_X
};

// stage s (0-based) = uprev + dt * sum_j coef[s][j] * K[slot[s][j]], terms in the reference's order; slots:
// k1 0, k2 1, k3 2, k4 3, k5 4, k6 5, k7 6, k8 7, k9 8, k10 2 (and uprev: see attempt)
struct B200V7WTab {
    real coef[10][8];
    int slot[10][8];
    int nterms[10];
    int out[10];
    real c[10];
};
#define B200_N7 b200_v7n
__constant__ B200V7WTab B200_V7W = {
    {{(real)0},
     {B200_N7.a021},
     {B200_N7.a031, B200_N7.a032},
     {B200_N7.a041, B200_N7.a043},
     {B200_N7.a051, B200_N7.a053, B200_N7.a054},
     {B200_N7.a061, B200_N7.a063, B200_N7.a064, B200_N7.a065},
     {B200_N7.a071, B200_N7.a073, B200_N7.a074, B200_N7.a075, B200_N7.a076},
     {B200_N7.a081, B200_N7.a083, B200_N7.a084, B200_N7.a085, B200_N7.a086, B200_N7.a087},
     {B200_N7.a091, B200_N7.a093, B200_N7.a094, B200_N7.a095, B200_N7.a096, B200_N7.a097, B200_N7.a098},
     {B200_N7.a101, B200_N7.a103, B200_N7.a104, B200_N7.a105, B200_N7.a106, B200_N7.a107}},
    {{0}, {0}, {0, 1}, {0, 2}, {0, 2, 3}, {0, 2, 3, 4}, {0, 2, 3, 4, 5}, {0, 2, 3, 4, 5, 6}, {0, 2, 3, 4, 5, 6, 7},
     {0, 2, 3, 4, 5, 6}},
    {0, 1, 2, 2, 3, 4, 5, 6, 7, 6},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 2},
    {(real)0, B200_N7.c2, B200_N7.c3, B200_N7.c4, B200_N7.c5, B200_N7.c6, B200_N7.c7, B200_N7.c8, (real)1, (real)1}};
#undef B200_N7

// the plain-operator evaluation of a flagged stage: out of line, on copies (keeps the hot arrays in registers)
__device__ __noinline__ void b200_wide_rhs_exact(real* du, const real* x, const real* p, real t) {
    B200UserExact e;
    e.B200_USER_FULL_NAME(du, x, p, t);
}
B200_D void b200_wide_rhs(real* du, const real* x, const real* p, real t) {
    B200UserFast f;
    f.b200_bad = false;
#pragma unroll
    for (int k = 0; k <= B200_WIDE_WINDOW; ++k) f.b200_win[k] = false;
    f.B200_USER_FULL_NAME(du, x, p, t);
    if (f.b200_bad) {
        real xc[B200_N], dc[B200_N];
#pragma unroll
        for (int i = 0; i < B200_N; ++i) xc[i] = x[i];
        b200_wide_rhs_exact(dc, xc, p, t);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) du[i] = dc[i];
    }
}

#define B200_WK(slot, i) Kt[((slot) * B200_N + (i)) * B200_WIDE_NT]

struct B200Vern7Wide {
    real* Kt;       // this thread's column of K[slot][component][thread]
    static B200_D int order() { return 7; }
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1; }

    B200_D void bind() { Kt = reinterpret_cast<real*>(b200_wide_smem_raw) + (threadIdx.x < B200_WIDE_NT ? threadIdx.x : 0); }
    B200_D void init(const real*, const real*, real, int&) {}
    B200_D void accept() {}
    // never reached: programs of this variant are not launched with interior saveat points (the shim rejects them)
    B200_D void dense_prepare(const real*, const real*, const real*, real, real) {}
    B200_D void interp(real, real, const real* y0, const real*, real* out) const {
#pragma unroll
        for (int i = 0; i < B200_N; ++i) out[i] = y0[i];
    }

    // sum of squared residuals (left fold, ODE_DEFAULT_NORM's numerator); FAST: flagged division.  uprev is read from
    // its shared-memory slot (UP), u goes to registers
    template <bool FAST>
    B200_D real finish(real* u, real dt, real reltol, real abstol, bool& bad) const {
        const B200V7Named& C = B200_V7W_C;
        real acc = (real)0;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) {
            const real k1 = B200_WK(0, i), k4 = B200_WK(3, i), k5 = B200_WK(4, i), k6 = B200_WK(5, i), k7 = B200_WK(6, i),
                       k8 = B200_WK(7, i), k9 = B200_WK(8, i), k10 = B200_WK(2, i), up = B200_WK(1, i);
            real sb = C.b1 * k1;
            sb = b200_fma(C.b4, k4, sb); sb = b200_fma(C.b5, k5, sb); sb = b200_fma(C.b6, k6, sb);
            sb = b200_fma(C.b7, k7, sb); sb = b200_fma(C.b8, k8, sb); sb = b200_fma(C.b9, k9, sb);
            const real ui = b200_fma(dt, sb, up);
            u[i] = ui;
            real se = C.btilde1 * k1;
            se = b200_fma(C.btilde4, k4, se); se = b200_fma(C.btilde5, k5, se); se = b200_fma(C.btilde6, k6, se);
            se = b200_fma(C.btilde7, k7, se); se = b200_fma(C.btilde8, k8, se); se = b200_fma(C.btilde9, k9, se);
            se = b200_fma(C.btilde10, k10, se);
            const real ut = dt * se;
            const real den = b200_fma(b200_max_fast(b200_abs(up), b200_abs(ui)), B200_RTOL_AT(i, reltol), B200_ATOL_AT(i, abstol));
            const real r = FAST ? b200_div_fast(ut, den, bad) : b200_div_cold(ut, den);
            acc = (i == 0) ? r * r : acc + r * r;
        }
        return acc;
    }

    // Shared-memory slots over one attempt (9 slots of n reals per thread):
    //   k1 0 | k2 1 | k3 2 | k4 3 | k5 4 | k6 5 | k7 6 | k8 7 | k9 8 | k10 2 (k3 is dead once stage 10's argument exists)
    //   uprev: slot 8 until stage 9's argument exists, then slot 1 (k2 is dead after stage 3) — so the 56 registers of
    //   uprev are free while the RHS runs; the caller's register copy is refreshed from slot 1 at the end.
    B200_D real attempt(real* uprev, real* u, const real* p, real t, real dt, real reltol, real abstol, int& nf) {
        const B200V7WTab& W = B200_V7W;
#pragma unroll
        for (int i = 0; i < B200_N; ++i) B200_WK(8, i) = uprev[i];
#pragma unroll 1
        for (int it = 0; it < 10; ++it) {
            int s = it;
            asm volatile("" : "+r"(s));       // opaque stage index: ONE loop body (no peeled first stage => no second RHS copy)
            const real* Kup = Kt + (s == 9 ? 1 : 8) * (B200_N * B200_WIDE_NT);
            real x[B200_N];
            if (s == 0) {
#pragma unroll
                for (int i = 0; i < B200_N; ++i) x[i] = Kup[i * B200_WIDE_NT];
            } else if (s == 1) {
                const real a = dt * W.coef[1][0];
#pragma unroll
                for (int i = 0; i < B200_N; ++i) x[i] = b200_fma(a, B200_WK(0, i), Kup[i * B200_WIDE_NT]);
            } else {
                // sum_j a_sj k_j in the reference's nesting (first product, then fmas in ascending j).  The loop over the
                // terms is software pipelined by hand: the 28 loads of term j + 1 are issued before the 28 fmas of term j,
                // so the shared-memory latency of a term hides behind the previous term's arithmetic (one warp per
                // scheduler: nothing else would hide it).  Every stage has at least two terms.
                const int nt = W.nterms[s];
                real cur[B200_N], nxt[B200_N];
                {
                    const real* K0 = Kt + W.slot[s][0] * (B200_N * B200_WIDE_NT);
                    const real* K1 = Kt + W.slot[s][1] * (B200_N * B200_WIDE_NT);
#pragma unroll
                    for (int i = 0; i < B200_N; ++i) cur[i] = K0[i * B200_WIDE_NT];
#pragma unroll
                    for (int i = 0; i < B200_N; ++i) nxt[i] = K1[i * B200_WIDE_NT];
                    const real a = W.coef[s][0];
#pragma unroll
                    for (int i = 0; i < B200_N; ++i) x[i] = a * cur[i];
                }
#pragma unroll 1
                for (int j = 1; j < nt; ++j) {
#pragma unroll
                    for (int i = 0; i < B200_N; ++i) cur[i] = nxt[i];
                    // (the last pass re-reads its own term: a harmless load that keeps the loop body branch-free)
                    const real* Kn = Kt + W.slot[s][j + 1 < nt ? j + 1 : j] * (B200_N * B200_WIDE_NT);
#pragma unroll
                    for (int i = 0; i < B200_N; ++i) nxt[i] = Kn[i * B200_WIDE_NT];
                    const real a = W.coef[s][j];
#pragma unroll
                    for (int i = 0; i < B200_N; ++i) x[i] = b200_fma(a, cur[i], x[i]);
                }
#pragma unroll
                for (int i = 0; i < B200_N; ++i) x[i] = b200_fma(dt, x[i], Kup[i * B200_WIDE_NT]);
                if (s == 8) {       // k9 is about to take slot 8: uprev moves to slot 1
#pragma unroll
                    for (int i = 0; i < B200_N; ++i) B200_WK(1, i) = B200_WK(8, i);
                }
            }
            const real ts = (s == 0) ? t : (s >= 8 ? t + dt : b200_fma(W.c[s], dt, t));
            real du[B200_N];
            b200_wide_rhs(du, x, p, ts);
            real* Ko = Kt + W.out[s] * (B200_N * B200_WIDE_NT);
#pragma unroll
            for (int i = 0; i < B200_N; ++i) Ko[i * B200_WIDE_NT] = du[i];
        }
        nf += 10;
        // u = uprev + dt (b . k), utilde = dt (btilde . k), calculate_residuals, norm — every k read once, nothing but u
        // kept: the flagged case recomputes the residuals from shared memory with the plain division
        bool bad = false;
        real acc = finish<true>(u, dt, reltol, abstol, bad);
        if (bad) acc = finish<false>(u, dt, reltol, abstol, bad);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) uprev[i] = B200_WK(1, i);
        bool bad2 = false;
        real e = b200_sqrt_fast(b200_div_const_fast(acc, (real)B200_N, (real)1 / (real)B200_N, bad2), bad2);
        if (bad2) e = b200_sqrt(acc / (real)B200_N);
        return e;
    }
};
