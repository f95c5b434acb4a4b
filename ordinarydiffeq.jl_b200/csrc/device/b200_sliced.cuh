// b200_sliced.cuh — "component-sliced" ensemble kernel for systems too large for one thread.
//
// One-trajectory-per-thread needs every stage vector of a trajectory in one thread: for
// Pleiades/Vern7 (n = 28, 16 stage vectors) that is 3.6 KB of state per thread, which lives in
// local memory and makes the kernel L1/L2-bandwidth bound (ncu r1: 13 GB of DRAM write-back,
// stall_no_instruction 2.1, 7 warps/SM).  Here a CTA of G warps owns 32 trajectories:
//   lane  = trajectory slot,  warp w = slice: components c with c % G == w  (VLEN = ceil(n/G) each).
// Stage vectors are exchanged through shared memory (double buffered, one __syncthreads per
// RHS evaluation); every warp evaluates only its own components of the user's RHS — the RHS is
// inlined once per slice inside one out-of-line function and the compiler's dead-code
// elimination strips the other components (the C source must be straight-line, as Symbolics
// emits it).  Each thread then holds 16 x VLEN stage values in registers.
// The scalar controller state is replicated in all G warps (bit-identical by construction);
// the error norm is the same left fold over components 0..n-1, read back from shared memory.
//
// Same reference semantics as b200_ensemble.cuh (citations there).  Limitations of this variant:
// explicit non-FSAL steppers (Vern7); one batch of 32 trajectories per CTA at a time.
#pragma once

#ifndef B200_G
#error "B200_G (slices per trajectory) must be defined for the sliced kernel"
#endif
#define B200_VLEN ((B200_N + B200_G - 1) / B200_G)
// B200_K groups of G slice-warps per CTA; group k owns trajectories [32k, 32k+32) of the CTA's batch.
// All warps of the CTA run in lockstep (one barrier per RHS evaluation), so warps that execute the
// same slice fetch the same instructions together.  G = 1 degenerates to one trajectory per thread
// with CTA-lockstepped stages ("lockstep" mode).
#ifndef B200_K
#define B200_K 1
#endif
#define B200_SLICE() ((int)(threadIdx.x >> 5) % B200_G)
#define B200_GROUP() ((int)(threadIdx.x >> 5) / B200_G)

extern __shared__ double b200_smem_raw[];

struct B200SlicedSmem {
    // U[2][N][32], R[N][32], F[G][32] (as real), I[32] (long long refill indices)
    B200_D static real* base() { return (real*)b200_smem_raw + (size_t)B200_GROUP() * (3 * B200_N * 32 + B200_G * 32); }
    B200_D static real* U(int buf) { return base() + (size_t)buf * B200_N * 32; }
    B200_D static real* R() { return base() + (size_t)2 * B200_N * 32; }
    B200_D static real* F() { return base() + (size_t)3 * B200_N * 32; }
};

// ---- one slice of the user's RHS -------------------------------------------------------------
#if B200_G > 1
struct B200VRet { real v[B200_VLEN]; };

template <int W>
B200_D B200VRet b200_eval_one(const real* __restrict__ Ub, int lane, const real* p, real t) {
    real uf[B200_N], df[B200_N];
#pragma unroll
    for (int c = 0; c < B200_N; ++c) uf[c] = Ub[c * 32 + lane];
    B200_USER_RHS(df, uf, p, t);
    B200VRet r;
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) r.v[l] = (W + B200_G * l < B200_N) ? df[(W + B200_G * l < B200_N) ? (W + B200_G * l) : 0] : (real)0;
    return r;
}

template <int W>
struct B200EvalDispatch {
    B200_D static B200VRet run(int w, const real* Ub, int lane, const real* p, real t) {
        if (w == W) return b200_eval_one<W>(Ub, lane, p, t);
        return B200EvalDispatch<W + 1>::run(w, Ub, lane, p, t);
    }
};
template <>
struct B200EvalDispatch<B200_G> {
    B200_D static B200VRet run(int, const real*, int, const real*, real) { B200VRet r;
#pragma unroll
        for (int l = 0; l < B200_VLEN; ++l) r.v[l] = (real)0;
        return r; }
};

// out of line: ONE copy of the (sliced) RHS in the kernel instead of one per stage
__device__ __noinline__ B200VRet b200_eval_slice(int w, const real* Ub, int lane, const real* p, real t) {
    return B200EvalDispatch<0>::run(w, Ub, lane, p, t);
}

#endif  // B200_G > 1

// publish my components of the stage vector, barrier, evaluate my slice
B200_D void b200_rhs_sliced(real* kout, const real* xin, const real* p, real t, int& sbuf) {
#if B200_G == 1
    __syncthreads();                       // lockstep only: keeps the CTA's warps on the same instructions
    B200_USER_RHS(kout, xin, p, t);
    (void)sbuf;
    return;
#endif
#if B200_G > 1
    const int lane = threadIdx.x & 31, w = B200_SLICE();
    real* Ub = B200SlicedSmem::U(sbuf);
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = w + B200_G * l;
        if (c < B200_N) Ub[c * 32 + lane] = xin[l];
    }
    __syncthreads();
    B200VRet r = b200_eval_slice(w, Ub, lane, p, t);
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) kout[l] = r.v[l];
    sbuf ^= 1;
#endif
}

// error norm: left fold of the squared residuals of components 0..n-1 (any warp order) + the
// "u is finite" flag of the new state, both exchanged through shared memory
B200_D real b200_norm_sliced(const real* res, const real* u, bool& all_finite) {
#if B200_G == 1
    {
        bool fin1 = true;
        real acc1 = res[0] * res[0];
#pragma unroll
        for (int i = 1; i < B200_N; ++i) acc1 = acc1 + res[i] * res[i];
#pragma unroll
        for (int i = 0; i < B200_N; ++i) fin1 = fin1 && b200_isfinite(u[i]);
        all_finite = fin1;
        return b200_sqrt(b200_div_const(acc1, (real)B200_N, (real)1 / (real)B200_N));
    }
#endif
    const int lane = threadIdx.x & 31, w = B200_SLICE();
    real* R = B200SlicedSmem::R();
    real* F = B200SlicedSmem::F();
    bool fin = true;
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = w + B200_G * l;
        if (c < B200_N) { R[c * 32 + lane] = res[l] * res[l]; fin = fin && b200_isfinite(u[l]); }
    }
    F[w * 32 + lane] = fin ? (real)1 : (real)0;
    __syncthreads();
    real acc = R[lane];
#pragma unroll
    for (int c = 1; c < B200_N; ++c) acc = acc + R[c * 32 + lane];
    bool af = true;
#pragma unroll
    for (int g = 0; g < B200_G; ++g) af = af && (F[g * 32 + lane] != (real)0);
    all_finite = af;
    __syncthreads();           // R/F are rewritten by the next attempt
    return b200_sqrt(b200_div_const(acc, (real)B200_N, (real)1 / (real)B200_N));
}

#define B200_STEPPER_EXTRA_MEMBERS int sbuf; bool all_finite;
#define B200_RHS(du, u, p, t) b200_rhs_sliced((du), (u), (p), (t), sbuf)
#define B200_NORM(res, u) b200_norm_sliced((res), (u), all_finite)
#include "b200_vern7.cuh"
typedef B200Vern7 B200SlicedStepper;

struct B200STraj {
    real u[B200_VLEN], uprev[B200_VLEN];
    real p[B200_NP > 0 ? B200_NP : 1];
    B200SlicedStepper st;
    real t, tprev, dt, dtpropose;
    real q11, EEst, fpe, rfpe, next_save;
    int naccept, nreject, nf;
    int save_idx, nsaved;
    int retcode;
    bool accept, tstop_flag;
};

B200_D void b200s_emit(const B200Params& P, long long idx, B200STraj& T, const real* v) {
    const int w = B200_SLICE();
    if (P.nslots > 0 && T.nsaved < P.nslots) {
        real* dst = P.us + ((size_t)idx * (size_t)P.nslots + (size_t)T.nsaved) * B200_N;
#pragma unroll
        for (int l = 0; l < B200_VLEN; ++l) {
            const int c = w + B200_G * l;
            if (c < B200_N) dst[c] = v[l];
        }
    }
    T.nsaved += 1;
}

B200_D void b200s_modify_dt_for_tstops(B200STraj& T, real dist, real tol100) {
    real orig = b200_abs(T.dt);
    T.dtpropose = orig;
    T.tstop_flag = !(orig + tol100 < dist);
    T.dt = b200_min_c(dist, orig);
}

B200_D void b200s_begin(const B200Params& P, long long idx, B200STraj& T, bool live) {
    const int w = B200_SLICE();
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = w + B200_G * l;
        real v = (c < B200_N) ? P.u0[idx * P.u0_ts + c * P.u0_cs] : (real)0;
        T.u[l] = v; T.uprev[l] = v;
    }
#pragma unroll
    for (int c = 0; c < B200_NP; ++c) T.p[c] = P.p[idx * P.p_ts + c * P.p_cs];
    T.t = P.t0; T.tprev = P.t0;
    T.nf = 0; T.nsaved = 0; T.save_idx = 0;
    if (P.save_start && live) b200s_emit(P, idx, T, T.u);
    if (P.dt_user == (real)0) { T.dt = P.dt0[idx]; T.nf += 2; } else T.dt = P.dt_user;
    T.dtpropose = T.dt;
    T.q11 = (real)1; T.EEst = (real)1;
    {
        const real beta2 = (real)(2.0 / (5.0 * B200SlicedStepper::order()));
        T.fpe = b200_fastpower((real)1e-4, beta2);
        T.rfpe = (real)1 / T.fpe;
    }
    T.next_save = (P.nsaveat > 0) ? P.saveat[0] : b200_inf();
    T.naccept = 0; T.nreject = 0;
    T.accept = false; T.tstop_flag = false;
    T.retcode = B200_RC_DEFAULT;
    T.st.all_finite = true;
}

B200_D void b200s_end(const B200Params& P, long long idx, B200STraj& T) {
    const int w = B200_SLICE();
    if (T.retcode == B200_RC_DEFAULT) T.retcode = B200_RC_SUCCESS;
    if (P.save_end) {
        bool emit;
        if (T.nsaved == 0) emit = true;
        else {
            const real last_t = (T.save_idx > 0) ? P.saveat[T.save_idx - 1] : P.t0;
            emit = (last_t != T.t) && (P.save_end == 2 || T.t == P.tf || P.nsaveat == 0);
        }
        if (emit) b200s_emit(P, idx, T, T.u);
    }
    if (P.nslots > 0 && T.nsaved < P.nslots) {
        for (int s = T.nsaved; s < P.nslots; ++s) {
            real* dst = P.us + ((size_t)idx * (size_t)P.nslots + (size_t)s) * B200_N;
#pragma unroll
            for (int l = 0; l < B200_VLEN; ++l) { const int c = w + B200_G * l; if (c < B200_N) dst[c] = (real)0; }
        }
    }
#pragma unroll
    for (int l = 0; l < B200_VLEN; ++l) {
        const int c = w + B200_G * l;
        if (c < B200_N) P.u_final[idx * P.uf_ts + c * P.uf_cs] = T.u[l];
    }
    if (w == 0) {
        P.t_final[idx] = T.t;
        P.naccept[idx] = T.naccept; P.nreject[idx] = T.nreject; P.nf[idx] = T.nf;
        P.retcode[idx] = T.retcode; P.nsaved[idx] = T.nsaved;
    }
}

// One pass of the solve! loop body for the 32 trajectories of this CTA.  Every thread executes
// every barrier; `live` lanes are the ones whose trajectory is still running.  Returns true when
// this lane's trajectory has finished (all G warps of a lane agree: their scalar state is identical).
B200_D bool b200s_iterate(const B200Params& P, long long idx, B200STraj& T, bool live) {
    const real qmin = (real)0.2, qmax = (real)10, gamma = (real)0.9;
    const real beta1 = (real)(7.0 / (10.0 * B200SlicedStepper::order()));
    const real beta2 = (real)(2.0 / (5.0 * B200SlicedStepper::order()));
    const int iter0 = T.naccept + T.nreject;
    const real dist = b200_abs(P.tf - T.t);
    const real at = b200_abs(T.t), atf = b200_abs(P.tf);
    const real tol100 = P.tol_const ? P.tol100_tf : (real)100 * b200_eps_finite(at > atf ? at : atf);
    const real eps_t = b200_eps_finite(T.t);
    const real dtmin_t = eps_t > P.dtmin ? eps_t : P.dtmin;
    bool ok = true;
    bool skip = false;
    if (live) {
        // ---- loopheader! ----
        if (iter0 > 0) {
            if (T.accept) {
#pragma unroll
                for (int l = 0; l < B200_VLEN; ++l) T.uprev[l] = T.u[l];
                T.dt = T.dtpropose;
                b200s_modify_dt_for_tstops(T, dist, tol100);
            } else {
                T.dt = T.dt / b200_min_c((real)1 / qmin, b200_div_const(T.q11, gamma, (real)1 / gamma));
            }
        }
        T.dt = b200_min_c(P.dtmax, T.dt);
        T.dt = b200_max_c(dtmin_t, T.dt);
        b200s_modify_dt_for_tstops(T, dist, tol100);
        // ---- check_error ----
        int rc = B200_RC_SUCCESS;
        if (b200_isnan(T.dt)) rc = B200_RC_DTNAN;
        else if ((long long)iter0 + 1 > P.maxiters) rc = B200_RC_MAXITERS;
        else if (b200_abs(T.dt) <= b200_abs(P.dtmin) && (!T.accept || T.t + T.dt < P.tf)) rc = B200_RC_DTLESSTHANMIN;
        else if (!T.accept && b200_abs(T.dt) <= eps_t) rc = B200_RC_UNSTABLE;
        else if (T.accept && !T.st.all_finite) rc = B200_RC_UNSTABLE;
        ok = (rc == B200_RC_SUCCESS);
        if (!ok) T.retcode = rc;
        skip = T.tstop_flag && b200_abs(T.dt) < eps_t;
    }
    // ---- perform_step!: executed by every thread (it contains CTA barriers); lanes that are not
    // stepping compute on their stale state and discard the result
    const bool do_step = live && ok && !skip;
    real unew[B200_VLEN];
    int nf_dummy = 0;
    const real e = T.st.attempt(T.uprev, unew, T.p, T.t, T.dt, P.reltol, P.abstol, do_step ? T.nf : nf_dummy);
    if (do_step) {
        T.EEst = e;
#pragma unroll
        for (int l = 0; l < B200_VLEN; ++l) T.u[l] = unew[l];
    }
    bool finished = !live || !ok;
    bool want_dense = false;
    real ttmp = T.t + T.dt;
    real q = (real)1;
    if (live && ok) {
        // ---- loopfooter! ----
        const real qmax_eff = (T.naccept == 0) ? (real)10000 : qmax;
        if (T.EEst == (real)0) {
            q = (real)1 / qmax_eff;
        } else {
            real q11 = b200_fastpower(T.EEst, beta1);
            q = b200_div_const(q11, T.fpe, T.rfpe);
            T.q11 = q11;
            q = b200_div_const(q, gamma, (real)1 / gamma);
            const real lo = (real)1 / qmax_eff, hi = (real)1 / qmin;
            q = q < lo ? lo : (q > hi ? hi : q);
        }
        T.accept = (T.EEst <= (real)1);
        if (T.accept) {
            T.naccept += 1;
            T.tprev = T.t;
            if (T.tstop_flag) T.dt = T.dtpropose;
            T.t = T.tstop_flag ? P.tf : ttmp;
            T.tstop_flag = false;
            if (B200SlicedStepper::qsteady_min() <= q && q <= B200SlicedStepper::qsteady_max()) q = (real)1;
            {
                const real errold = b200_max_c((real)1e-4, T.EEst);
                T.fpe = b200_fastpower(errold, beta2);
                T.rfpe = (real)1 / T.fpe;
            }
            const real dtnew = T.dt / q;
            const real eps_n = b200_eps_finite(T.t);
            T.dtpropose = b200_max_c(eps_n > P.dtmin ? eps_n : P.dtmin, b200_min_c(b200_abs(P.dtmax), b200_abs(dtnew)));
            want_dense = (T.next_save < T.t);       // an interior saveat point in (tprev, t)
        } else {
            T.nreject += 1;
        }
    }
    // ---- savevalues!: the lazy extra stages are evaluated by the whole CTA if any lane needs them
    if (__syncthreads_or(want_dense ? 1 : 0)) T.st.dense_prepare(T.uprev, T.u, T.p, T.tprev, T.dt);
    if (live && ok && T.accept) {
        while (T.next_save <= T.t) {
            const real curt = T.next_save;
            T.save_idx += 1;
            T.next_save = (T.save_idx < P.nsaveat) ? P.saveat[T.save_idx] : b200_inf();
            if (curt != T.t) {
                const real th = (curt - T.tprev) / T.dt;
                real out[B200_VLEN];
                T.st.interp(th, T.dt, T.uprev, T.u, out);
                b200s_emit(P, idx, T, out);
            } else {
                if (curt == P.tf && !P.save_end) continue;
                b200s_emit(P, idx, T, T.u);
            }
        }
        finished = !(T.t < P.tf);
    }
    return finished;
}

extern "C" __global__ void __launch_bounds__(32 * B200_G * B200_K, B200_MINBLOCKS) b200_integrate(B200Params P) {
    B200STraj T;
    T.st.sbuf = 0;
    const int lane = threadIdx.x & 31;
    // batches of 32*K trajectories, CTA-strided
    const long long nbatch = (P.N + 32 * B200_K - 1) / (32 * B200_K);
    for (long long b = blockIdx.x; b < nbatch; b += gridDim.x) {
        const long long idx = b * (32 * B200_K) + B200_GROUP() * 32 + lane;
        bool live = idx < P.N;
        const long long idx_c = live ? idx : (P.N - 1);      // inactive lanes shadow a valid trajectory, write nothing
        b200s_begin(P, idx_c, T, live);
        if (live && !(T.t < P.tf)) { b200s_end(P, idx, T); live = false; }
        while (__syncthreads_or(live ? 1 : 0)) {
            const bool fin = b200s_iterate(P, idx_c, T, live);
            if (live && fin) { b200s_end(P, idx, T); live = false; }
        }
    }
}
