// b200_coop.cuh — lane-group ("warp-cooperative") ensemble kernel for systems too large for one thread.
//
// One trajectory per thread needs every stage vector of a trajectory in one thread: for Pleiades/Vern7
// (n = 28, 16 stage vectors) that is 3.6 KB of state per thread, which ptxas can only keep in local memory
// (r1: 1.375 M traj/s, 13 GB of DRAM write-back for 15 MB of results, FP64 pipe 22 % busy).  Here a group of
// B200_L lanes of one warp owns a trajectory:
//     lane g of the group holds components c = g + LA*l, l = 0 .. VLEN-1   (VLEN = ceil(n / L), LA = ceil(n / VLEN))
// of every stage vector in registers (16 x VLEN doubles), so nothing spills.  For Pleiades L = 16, VLEN = 2,
// LA = 14: lane g holds position component g and its velocity component g + 14, two trajectories per warp.
//   * RHS: the lanes publish their components of the stage state to shared memory (double buffered, one
//     __syncwarp per evaluation) and each lane evaluates ITS components from the full state.  This needs the
//     right-hand side in component form   real f(int i, const real* u, const real* p, real t)  = du_i
//     (program option B200ODE_OPT_COMPONENT_RHS) — one compact out-of-line function shared by all lanes and all
//     stages, which is what keeps the instruction stream inside the instruction cache (the r1 "sliced" kernel
//     inlined 7 differently dead-code-eliminated copies of a 43 KB straight-line RHS and was fetch bound).
//   * stage sums, error residuals, interpolation: lane-local on the VLEN components (same fma nesting as the
//     one-thread kernels — the stepper headers are shared, see B200_VLEN / B200_RHS / B200_NORM in b200_vern7.cuh).
//   * error norm: the reference's left fold over components 0..n-1 (common_defaults.jl:102-107) — every lane of the
//     group reads the n squared residuals back from shared memory and adds them in that order, so all lanes hold
//     the same EEst bits and the scalar controller state is replicated without any further exchange.
//   * scalar control (loopheader!, check_error, PI controller, saveat bookkeeping): identical in all lanes of a
//     group by construction; the code is the one-thread kernel's (b200_controller_t, b200_modify_dt_for_tstops).
//   * the whole warp stays converged: every lane executes every RHS evaluation (groups that are not stepping
//     compute on stale state and discard the result), finished groups pull the next trajectory from the global
//     counter.
// Same reference semantics as b200_ensemble.cuh (citations there).  Limitations of this variant: adaptive stepping,
// tstops = {tf}, rectangular saveat output (no save_everystep / save_idxs / dense); steppers: Vern7 (B200_VLEN hooks of
// b200_vern7.cuh) and Rosenbrock23 with a warp-shuffle LU (b200_ros23_coop.cuh).
#pragma once

#ifndef B200_L
#error "B200_L (lanes per trajectory: 2, 4, 8, 16 or 32) must be defined for the lane-group kernel"
#endif
#if (B200_L & (B200_L - 1)) != 0 || B200_L < 2 || B200_L > 32
#error "B200_L must be a power of two in 2..32"
#endif
#define B200_VLEN ((B200_N + B200_L - 1) / B200_L)
#define B200_LA ((B200_N + B200_VLEN - 1) / B200_VLEN)      // lanes of a group that own components
#define B200_GPW (32 / B200_L)                               // groups (trajectories) per warp

// shared memory per group: stage state U[2][n] (double buffered) + squared residuals R[n]
#define B200_COOP_WORDS (3 * B200_N)
__shared__ real b200_coop_smem[(B200_BLOCK / 32) * B200_GPW * B200_COOP_WORDS];

// All VLEN components of lane g, evaluated by ONE out-of-line function (one copy in the kernel for all stages).
// The components are evaluated with the flagged fast division / square root first (B200UserFast, see the shim): all of
// them in one basic block, so common subexpressions of neighbouring components (Pleiades: r^3 of the x and y
// acceleration of one body) are computed once and independent terms overlap; the plain operators (B200UserExact) redo
// the lane only if a flag was raised.
struct B200VRet { real v[B200_VLEN]; };
#ifndef B200_COOP_INLINE
#define B200_COOP_INLINE 0      // 1: inline the lane evaluator at every stage (measured: see DESIGN.md §4)
#endif
#if B200_COOP_INLINE
__device__ __forceinline__ B200VRet b200_rhs_lane(int g, const real* Ub, const real* p, real t);
#else
__device__ __noinline__ B200VRet b200_rhs_lane(int g, const real* Ub, const real* p, real t);
#endif
B200VRet b200_rhs_lane(int g, const real* Ub, const real* p, real t) {
    B200VRet r;
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) r.v[l] = (real)0;
    if (g >= B200_LA) return r;                 // lanes that own nothing (the function contains no warp sync)
    __builtin_assume(g >= 0 && g < B200_LA);    // lets the compiler resolve index tests of the form i < k per slot
    B200UserFast f;
    f.b200_bad = false;
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = g + B200_LA * l;
        if (B200_LA * l + B200_LA <= B200_N || c < B200_N) r.v[l] = f.B200_USER_COMP_NAME(c, Ub, p, t);
    }
    if (f.b200_bad) {
        B200UserExact e;
#pragma unroll
        for (int l = 0; l < B200_VLEN; ++l) {
            const int c = g + B200_LA * l;
            if (c < B200_N) r.v[l] = e.B200_USER_COMP_NAME(c, Ub, p, t);
        }
    }
    return r;
}

// publish my components of the stage state, sync the warp, evaluate my components
B200_D void b200_rhs_coop(real* kout, const real* xin, const real* p, real t, int& sbuf, real* sm, int g) {
    real* Ub = sm + sbuf * B200_N;
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = g + B200_LA * l;
        if (g < B200_LA && c < B200_N) Ub[c] = xin[l];
    }
    __syncwarp();
    const B200VRet r = b200_rhs_lane(g, Ub, p, t);
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = g + B200_LA * l;
        kout[l] = (g < B200_LA && c < B200_N) ? r.v[l] : (real)0;
    }
    sbuf ^= 1;      // the next stage state goes to the other buffer: no lane can still be reading it (one sync behind)
}

// error norm (left fold over components 0..n-1) + "the new state is finite", exchanged inside the group
B200_D real b200_norm_coop(const real* res, const real* u, bool& all_finite, real* sm, int g, unsigned gmask) {
    real* R = sm + 2 * B200_N;
    bool fin = true;
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = g + B200_LA * l;
        if (g < B200_LA && c < B200_N) { R[c] = res[l] * res[l]; fin = fin && b200_isfinite(u[l]); }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, fin);      // also orders the R stores before the loads below
    all_finite = ((bal & gmask) == gmask);
    __syncwarp();
    real acc = R[0];
#pragma unroll
    for (int c = 1; c < B200_N; ++c) acc = acc + R[c];
    bool bad = false;
    real e = b200_sqrt_fast(b200_div_const_fast(acc, (real)B200_N, (real)1 / (real)B200_N, bad), bad);
    if (bad) e = b200_sqrt(acc / (real)B200_N);
    return e;
}

#define B200_STEPPER_EXTRA_MEMBERS int sbuf; bool all_finite; real* sm; int g; unsigned gmask;
#define B200_RHS(du, u, p, t) b200_rhs_coop((du), (u), (p), (t), sbuf, sm, g)
#define B200_NORM(res, u) b200_norm_coop((res), (u), all_finite, sm, g, gmask)
#if B200_ALG == B200_ALG_ROS23
#include "b200_ros23_coop.cuh"
typedef B200Ros23Coop B200CoopStepper;
#else
#include "b200_vern7.cuh"
typedef B200Vern7 B200CoopStepper;
#endif

struct B200CTraj {
    real u[B200_VLEN], uprev[B200_VLEN];
    real p[B200_NP > 0 ? B200_NP : 1];
    B200CoopStepper st;
    real t, tprev, dt, dtpropose;
    real q11, EEst, fpe, rfpe, next_save;
    int naccept, nreject, nf;
    int save_idx, nsaved;
    int retcode;
    bool accept, tstop_flag;
#if B200_IS_ROSENBROCK
    int njacs, nw, nsolve;
#endif
};

B200_D void b200c_emit(const B200Params& P, long long idx, B200CTraj& T, const real* v, int g) {
    if (P.nslots > 0 && T.nsaved < P.nslots) {
        real* dst = P.us + ((size_t)idx * (size_t)P.nslots + (size_t)T.nsaved) * B200_N;
#pragma unroll
        for (int l = 0; l < B200_VLEN; ++l) {
            const int c = g + B200_LA * l;
            if (g < B200_LA && c < B200_N) dst[c] = v[l];
        }
    }
    T.nsaved += 1;
}

B200_D void b200c_modify_dt_for_tstops(B200CTraj& T, real dist, real tol100) {
    const real orig = b200_abs(T.dt);
    T.dtpropose = orig;
    T.tstop_flag = !(orig + tol100 < dist);
    T.dt = b200_min_c(dist, orig);
}

B200_D void b200c_begin(const B200Params& P, long long idx, B200CTraj& T, int g) {
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = g + B200_LA * l;
        const real v = (g < B200_LA && c < B200_N) ? P.u0[idx * P.u0_ts + c * P.u0_cs] : (real)0;
        T.u[l] = v; T.uprev[l] = v;
    }
#pragma unroll
    for (int c = 0; c < B200_NP; ++c) T.p[c] = P.p[idx * P.p_ts + c * P.p_cs];
    T.t = P.t0; T.tprev = P.t0;
    T.nf = 0; T.nsaved = 0; T.save_idx = 0;
    if (P.save_start) b200c_emit(P, idx, T, T.u, g);
    if (P.dt_user == (real)0) { T.dt = P.dt0[idx]; T.nf += 2; } else T.dt = P.dt_user;
    T.dtpropose = T.dt;
    T.q11 = (real)1; T.EEst = (real)1;
    T.fpe = P.fpe0; T.rfpe = P.rfpe0;
    T.next_save = (P.nsaveat > 0) ? P.saveat[0] : b200_inf();
    T.naccept = 0; T.nreject = 0;
    T.accept = false; T.tstop_flag = false;
    T.retcode = B200_RC_DEFAULT;
    T.st.all_finite = true;
#if B200_IS_ROSENBROCK
    T.njacs = 0; T.nw = 0; T.nsolve = 0;
#endif
    T.st.init(T.u, T.p, T.t, T.nf);         // (a stepper with a first-same-as-last stage defers the evaluation to its first attempt)
}

B200_D void b200c_end(const B200Params& P, long long idx, B200CTraj& T, int g) {
    if (T.retcode == B200_RC_DEFAULT) T.retcode = B200_RC_SUCCESS;
    if (P.save_end) {       // solution_endpoint_match_cur_integrator! (see b200_traj_end)
        bool emit;
        if (T.nsaved == 0) emit = true;
        else {
            const real last_t = (T.save_idx > 0) ? P.saveat[T.save_idx - 1] : P.t0;
            emit = (last_t != T.t) && (P.save_end == 2 || T.t == P.tf || P.nsaveat == 0);
        }
        if (emit) b200c_emit(P, idx, T, T.u, g);
    }
    if (P.nslots > 0 && T.nsaved < P.nslots) {          // a failed trajectory leaves its remaining rows zero
        for (int s = T.nsaved; s < P.nslots; ++s) {
            real* dst = P.us + ((size_t)idx * (size_t)P.nslots + (size_t)s) * B200_N;
#pragma unroll
            for (int l = 0; l < B200_VLEN; ++l) { const int c = g + B200_LA * l; if (g < B200_LA && c < B200_N) dst[c] = (real)0; }
        }
    }
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = g + B200_LA * l;
        if (g < B200_LA && c < B200_N) P.u_final[idx * P.uf_ts + c * P.uf_cs] = T.u[l];
    }
    if (g == 0) {
        P.t_final[idx] = T.t;
        P.naccept[idx] = T.naccept; P.nreject[idx] = T.nreject; P.nf[idx] = T.nf;
        P.retcode[idx] = T.retcode; P.nsaved[idx] = T.nsaved;
#if B200_IS_ROSENBROCK
        P.njacs[idx] = T.njacs; P.nw[idx] = T.nw; P.nsolve[idx] = T.nsolve;
#endif
    }
}

// One pass of the solve! loop body for the trajectories of this warp.  Every lane executes every __syncwarp;
// `live` groups are the ones whose trajectory is still running.  Returns true when this group's trajectory has
// finished (all lanes of a group agree: their scalar state is identical).
B200_D bool b200c_iterate(const B200Params& P, long long idx, B200CTraj& T, bool live, int g) {
    const int iter0 = T.naccept + T.nreject;
    const real dist = b200_abs(P.tf - T.t);
    const real at = b200_abs(T.t), atf = b200_abs(P.tf);
    const real tol100 = P.tol_const ? P.tol100_tf : (real)100 * b200_eps_finite(at > atf ? at : atf);
    const real eps_t = b200_eps_finite(T.t);
    const real dtmin_t = eps_t > P.dtmin ? eps_t : P.dtmin;
    bool ok = true, skip = false;
    if (live) {
        // ---- loopheader! ---- (a rejected step already carries its reduced dt, see b200_controller_t)
        if (iter0 > 0 && T.accept) {
#pragma unroll
            for (int l = 0; l < B200_VLEN; ++l) T.uprev[l] = T.u[l];
            T.dt = T.dtpropose;
            T.st.accept();                      // update_fsal!
            b200c_modify_dt_for_tstops(T, dist, tol100);
        }
        T.dt = b200_min_c(P.dtmax, T.dt);
        T.dt = b200_max_c(dtmin_t, T.dt);
        b200c_modify_dt_for_tstops(T, dist, tol100);
        // ---- check_error ----
        const bool c_nan = b200_isnan(T.dt);
        const bool c_max = ((long long)iter0 + 1 > P.maxiters);
        const bool c_min = (b200_abs(T.dt) <= b200_abs(P.dtmin)) & (!T.accept | (T.t + T.dt < P.tf));
        const bool c_uns = (!T.accept) & (b200_abs(T.dt) <= eps_t);
        const bool c_inf = T.accept & !T.st.all_finite;
        ok = !(c_nan | c_max | c_min | c_uns | c_inf);
        if (!ok)
            T.retcode = c_nan ? B200_RC_DTNAN : (c_max ? B200_RC_MAXITERS : (c_min ? B200_RC_DTLESSTHANMIN : B200_RC_UNSTABLE));
        skip = T.tstop_flag && b200_abs(T.dt) < eps_t;
    }
    // ---- perform_step!: executed by every lane (it contains warp syncs); groups that are not stepping compute on
    // their stale state and discard the result
    const bool do_step = live && ok && !skip;
    real unew[B200_VLEN];
    int nf_dummy = 0;
    const bool fin_before = T.st.all_finite;
#if B200_IS_ROSENBROCK
    int nj_dummy = 0, nw_dummy = 0, ns_dummy = 0;
    const real e = T.st.attempt(T.uprev, unew, T.p, T.t, T.dt, P.reltol, P.abstol, do_step ? T.nf : nf_dummy,
                                do_step ? T.njacs : nj_dummy, do_step ? T.nw : nw_dummy, do_step ? T.nsolve : ns_dummy, do_step);
#else
    const real e = T.st.attempt(T.uprev, unew, T.p, T.t, T.dt, P.reltol, P.abstol, do_step ? T.nf : nf_dummy);
#endif
    if (do_step) {
        T.EEst = e;
#pragma unroll
        for (int l = 0; l < B200_VLEN; ++l) T.u[l] = unew[l];
    } else {
        T.st.all_finite = fin_before;
    }
    bool finished = !live || !ok;
    bool want_dense = false;
    if (live && ok) {
        // ---- loopfooter! ----
        const real ttmp = T.t + T.dt;
        B200Ctl ctl;
        {
            bool bad = false;
            ctl = b200_controller_t<true>(T.EEst, T.q11, T.fpe, T.rfpe, T.dt, T.dtpropose, T.tstop_flag, T.naccept == 0, bad, b200_ctl_cfg_static());
            if (bad) {
                bool unused = false;
                ctl = b200_controller_t<false>(T.EEst, T.q11, T.fpe, T.rfpe, T.dt, T.dtpropose, T.tstop_flag, T.naccept == 0, unused, b200_ctl_cfg_static());
            }
        }
        T.q11 = ctl.q11;
        T.accept = ctl.accept;
        if (T.accept) {
            T.naccept += 1;
            T.tprev = T.t;
            T.dt = ctl.num;
            T.t = T.tstop_flag ? P.tf : ttmp;
            T.tstop_flag = false;
            T.fpe = ctl.fpe;
            T.rfpe = ctl.rfpe;
            const real eps_n = b200_eps_finite(T.t);
            T.dtpropose = b200_max_c(eps_n > P.dtmin ? eps_n : P.dtmin, b200_min_c(b200_abs(P.dtmax), b200_abs(ctl.dtdiv)));
            want_dense = (T.next_save < T.t);       // an interior saveat point in (tprev, t)
        } else {
            T.nreject += 1;
            T.dt = ctl.dtdiv;
        }
    }
    // ---- savevalues!: the lazy extra stages are evaluated by the whole warp if any group needs them
    if (__any_sync(0xffffffffu, want_dense)) T.st.dense_prepare(T.uprev, T.u, T.p, T.tprev, T.dt);
    if (live && ok && T.accept) {
        while (T.next_save <= T.t) {
            const real curt = T.next_save;
            T.save_idx += 1;
            T.next_save = (T.save_idx < P.nsaveat) ? P.saveat[T.save_idx] : b200_inf();
            if (curt != T.t) {
                const real th = (curt - T.tprev) / T.dt;
                real out[B200_VLEN];
                T.st.interp(th, T.dt, T.uprev, T.u, out);
                b200c_emit(P, idx, T, out, g);
            } else {
                if (curt == P.tf && !P.save_end) continue;
                b200c_emit(P, idx, T, T.u, g);
            }
        }
        finished = !(T.t < P.tf);
    }
    return finished;
}

extern "C" __global__ void __launch_bounds__(B200_BLOCK, B200_MINBLOCKS) b200_integrate(B200Params P) {
    B200CTraj T;
    const int lane = threadIdx.x & 31;
    const int g = lane % B200_L;                       // my lane inside the group
    const int grp = lane / B200_L;                     // my group inside the warp
    const unsigned gmask = (B200_L == 32 ? 0xffffffffu : ((1u << B200_L) - 1u)) << (grp * B200_L);
    T.st.sbuf = 0;
    T.st.g = g;
    T.st.gmask = gmask;
    T.st.sm = b200_coop_smem + ((size_t)(threadIdx.x >> 5) * B200_GPW + grp) * B200_COOP_WORDS;
    T.st.all_finite = true;
#if B200_ALG == B200_ALG_ROS23
    T.st.need_f0 = false;
    T.st.f0[0] = (real)0; T.st.f2[0] = (real)0; T.st.k1[0] = (real)0; T.st.k2[0] = (real)0;
#endif
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) { T.u[l] = (real)0; T.uprev[l] = (real)0; }
#pragma unroll
    for (int c = 0; c < (B200_NP > 0 ? B200_NP : 1); ++c) T.p[c] = (real)0;
    T.t = P.t0; T.tprev = P.t0; T.dt = (real)0; T.dtpropose = (real)0;
    long long idx = 0;
    bool live = false, exhausted = false;
    // a group that has no trajectory shadows trajectory 0 (valid memory, nothing is written)
    for (;;) {
        if (!live && !exhausted) {
            long long next = 0;
            if (g == 0) next = (long long)atomicAdd(P.work_counter, 1ull);
            next = __shfl_sync(gmask, next, grp * B200_L);
            if (next < P.N) {
                idx = next; live = true;
                b200c_begin(P, idx, T, g);
                if (!(T.t < P.tf)) { b200c_end(P, idx, T, g); live = false; }
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, live)) {
            if (__all_sync(0xffffffffu, exhausted)) break;
            continue;
        }
        const bool fin = b200c_iterate(P, live ? idx : 0, T, live, g);
        if (live && fin) { b200c_end(P, idx, T, g); live = false; }
    }
}
