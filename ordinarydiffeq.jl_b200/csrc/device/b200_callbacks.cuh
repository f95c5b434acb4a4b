// b200_callbacks.cuh — events: ContinuousCallback (root finding on the step's interpolant) and DiscreteCallback,
// per trajectory, inside the accepting step of b200_integrate.  First slice of SURVEY §8(f) row 4: continuous callbacks
// with Tsit5 (they need the stepper's _ode_addsteps!(always_calc_begin = true)), discrete callbacks with every stepper.
//
// Reference behaviour reproduced (file:line under /root/reference):
//   handle_callbacks!                       lib/OrdinaryDiffEqCore/src/integrators/integrator_utils.jl:1081-1132
//   find_first_continuous_callback          lib/DiffEqBase/src/callbacks.jl:140-226 (earliest event wins, ties keep the first)
//   find_callback_time(::ContinuousCallback) callbacks.jl:361-403; nudge_tprev :413-422; check_event_occurrence :427-454
//   get_condition                           callbacks.jl:91-137 (u at t, uprev at tprev, the interpolant in between)
//   is_event_occurrence                     callbacks.jl:523-528
//   find_root                               callbacks.jl:478-491 — IntervalNonlinearProblem + ModAB(), abstol = reltol = 0 (EXT,
//                                           BracketingNonlinearSolve): every bracketing method ends on the two adjacent floats
//                                           around the sign change; bisection here, an exact zero counts as the far side
//   apply_callback!                         callbacks.jl:557-637 (set_proposed_dt!, change_t_via_interpolation!, save_positions)
//   apply_discrete_callback!                callbacks.jl:649-690
//   change_t_via_interpolation!             lib/OrdinaryDiffEqCore/src/integrators/integrator_interface.jl:5-39
//   reeval_internals_due_to_modification!   integrator_interface.jl:54-80 (k recomputed for the shortened step, reeval_fsal)
//   update_fsal! / reset_fsal!              integrator_utils.jl:215-239,1325-1343 (fsalfirst = f(u, p, t), nf += 1)
//   terminate!                              integrator_interface.jl:443-446
//   range(tprev, stop = t, length = interp_points)   Julia Base twiceprecision.jl (_linspace + getindex; EXT)
//
// The shim generates, from the user's C sources, the table B200_CB[] and the two dispatchers
//   real b200_cb_condition(int k, const real* u, const real* p, real t)
//   void b200_cb_affect(int k, bool neg, real* u, real* p, real t, int* terminate)
// (continuous callbacks first, then discrete ones — the CallbackSet order).
#pragma once

struct B200CbInfo {
    int kind;                 // 0 discrete, 1 continuous
    int has_affect, has_neg;  // affect! / affect_neg! !== nothing
    int rootfind;             // 0 none, 1 left, 2 right
    int interp_points;
    int save_before, save_after;
    double abstol;            // compared in Float64 (10eps() is a Float64 whatever the state type is)
    real nudge;
};

__constant__ B200CbInfo B200_CB[] = B200_CB_TABLE;

B200_D real b200_sign(real x) { return x > (real)0 ? (real)1 : (x < (real)0 ? (real)-1 : x); }   // sign(±0) = ±0, sign(NaN) = NaN

// ---- Julia's range(start, stop = stop, length = len)[i] (general _linspace path, see the oracle's JlLinspace) -----------
struct B200Linspace {
    real ref_hi, ref_lo, step_hi, step_lo; int offset;
    static B200_D void add12(real x, real y, real& hi, real& lo) {
        if (b200_abs(y) > b200_abs(x)) { const real s = x; x = y; y = s; }
        hi = x + y; lo = (x - hi) + y;
    }
    static B200_D real truncbits(real x, int nb) {
#if B200_F32
        return b200_u2f(b200_f2u(x) & (0xFFFFFFFFu << nb));
#else
        return b200_u2d(b200_d2u(x) & (0xFFFFFFFFFFFFFFFFull << nb));
#endif
    }
    B200_D void build(real start, real stop, int len) {
        const real delta = stop - start;
        const real tmin = -(start / delta);
        const real timin = rint(tmin * (real)(len - 1) + (real)1);
        int imin = timin <= (real)1 ? 1 : (timin >= (real)len ? len : (int)timin);
        real ref, step;
        if (1 < imin && imin < len) {
            const double t = (double)(imin - 1) / (double)(len - 1);
            ref = (real)((1.0 - t) * (double)start + t * (double)stop);
            step = (imin - 1 < len - imin) ? (ref - start) / (real)(imin - 1) : (stop - ref) / (real)(len - imin);
        } else if (imin <= 1) { imin = 1; ref = start; step = delta / (real)(len - 1); }
        else { imin = len; ref = stop; step = delta / (real)(len - 1); }
#if B200_F32
        const real m = b200_u2f(0x7F7FFFFEu);               // prevfloat(floatmax(Float32))
        const int prec_half = 12;
#else
        const real m = b200_u2d(0x7FEFFFFFFFFFFFFEull);     // prevfloat(floatmax(Float64))
        const int prec_half = 27;
#endif
        const int mx = (imin - 1 > len - imin) ? imin - 1 : len - imin;
        const real k = (real)mx;
        const real lo1 = -(m + ref) / k, lo2 = (-m + ref) / k, hi1 = (m - ref) / k, hi2 = (m + ref) / k;
        const real lo = lo1 > lo2 ? lo1 : lo2, hi = hi1 < hi2 ? hi1 : hi2;
        const real step_pre = step < lo ? lo : (step > hi ? hi : step);
        int nbl = 0;                                        // len < 2 ? 0 : ceil(Int, log2(mx)) + 1 (integer arithmetic)
        if (len >= 2) { int c = 0; while ((1 << c) < mx) ++c; nbl = c + 1; }
        const int nb = prec_half < nbl ? prec_half : nbl;
        step_hi = truncbits(step_pre, nb);
        real x1h, x1l, x2h, x2l;
        add12((real)(1 - imin) * step_hi, ref, x1h, x1l);
        add12((real)(len - imin) * step_hi, ref, x2h, x2l);
        const real a = (start - x1h) - x1l, b = (stop - x2h) - x2l;
        step_lo = (b - a) / (real)(len - 1);
        ref_hi = ref; ref_lo = a - (real)(1 - imin) * step_lo;
        offset = imin;
    }
    B200_D real at(int i) const {
        const real u = (real)(i - offset);
        const real sh = u * step_hi, sl = u * step_lo;
        real xh, xl; add12(ref_hi, sh, xh, xl);
        return xh + (xl + (sl + ref_lo));
    }
};

// ---- savevalues! (integrator_utils.jl:336-414) as a function; returns savedexactly ------------------------------------
B200_D bool b200_savevalues(const B200Params& P, long long idx, B200Traj& T, bool force_save) {
    bool savedexactly = false, dense_ready = false;
    while (T.next_save <= T.t) {
        const real curt = T.next_save;
        T.save_idx += 1;
        T.next_save = (T.save_idx < P.nsaveat) ? P.saveat[T.save_idx] : b200_inf();
        if (curt != T.t) {
            if (!dense_ready) { T.st.dense_prepare(T.uprev, T.u, T.p, T.tprev, T.dt); dense_ready = true; }
            const real th = (curt - T.tprev) / T.dt;
            real out[B200_N];
            T.st.interp(th, T.dt, T.uprev, T.u, out);
            b200_emit(P, idx, T, curt, out);
        } else {
            if (curt == P.tf && !P.save_end) continue;      // skip_saveat_at_tspan_end
            savedexactly = true;
            b200_emit(P, idx, T, T.t, T.u);
        }
    }
#if B200_EVERYSTEP
    const bool everystep = !(P.flags & B200_FLAG_NO_STEP_ROWS);
    if (force_save || (everystep && (T.nsaved == 0 || ((T.t != T.last_t || T.dt == (real)0) && (P.save_end || T.t != P.tf))))) {
        savedexactly = true;
        b200_emit(P, idx, T, T.t, T.u, (real)0);
    }
#else
    (void)force_save;       // rectangular programs are compiled only with save_positions = (false, false)
#endif
    return savedexactly;
}

B200_D void b200_run_affect(int k, bool neg, B200Traj& T) {
    int term = 0;
    b200_cb_affect(k, neg, T.u, T.p, T.t, &term);
    if (term) { T.terminated = true; T.retcode = B200_RC_TERMINATED; }
}

#if B200_NCC > 0
// get_condition
B200_D real b200_get_condition(int k, B200Traj& T, real abst) {
    if (abst == T.t) return b200_cb_condition(k, T.u, T.p, abst);
    if (abst == T.tprev) return b200_cb_condition(k, T.uprev, T.p, abst);
    real val[B200_N];
    const real th = (abst - T.tprev) / T.dt;                // current_interpolant
    T.st.interp(th, T.dt, T.uprev, T.u, val);
    return b200_cb_condition(k, val, T.p, abst);
}

B200_D bool b200_is_event(const B200CbInfo& cb, real prev_sign, real next_sign) {
    return ((prev_sign < (real)0 && cb.has_affect) || (prev_sign > (real)0 && cb.has_neg)) && prev_sign * next_sign <= (real)0;
}

// find_callback_time(integrator, callback::ContinuousCallback, callback_idx)
B200_D bool b200_find_callback_time(int k, int callback_idx, B200Traj& T, real& callback_t, real& bottom_sign, real& residual) {
    const B200CbInfo cb = B200_CB[k];
    real bottom_t = T.tprev;
    real bottom_condition = b200_get_condition(k, T, bottom_t);
    if (T.event_last == callback_idx) {
        if ((double)b200_abs(bottom_condition - T.last_event_error) <= cb.abstol) bottom_t = T.tprev + T.dt * cb.nudge;
        else bottom_t = T.tprev;
        bottom_condition = b200_get_condition(k, T, bottom_t);
    }
    bottom_sign = b200_sign(bottom_condition);
    real top_t = T.t;
    real top_sign = b200_sign(b200_get_condition(k, T, top_t));
    bool occurred = b200_is_event(cb, bottom_sign, top_sign);
    if (cb.interp_points >= 2 && !occurred) {
        B200Linspace ts;
        ts.build(T.tprev, T.t, cb.interp_points);
        for (int i = 2; i <= cb.interp_points; ++i) {
            top_t = (i == cb.interp_points) ? T.t : ts.at(i);
            top_sign = b200_sign(b200_get_condition(k, T, top_t));
            occurred = b200_is_event(cb, bottom_sign, top_sign);
            if (occurred) break;
        }
    }
    if (!occurred) { callback_t = T.t; residual = (real)0; }
    else if (cb.rootfind == 0 || top_sign == (real)0) { callback_t = top_t; residual = (real)0; }
    else {
        real left = bottom_t, right = top_t;
        for (;;) {
            const real mid = left + (right - left) / (real)2;
            if (!(left < mid && mid < right)) break;
            const real sm = b200_sign(b200_get_condition(k, T, mid));
            if (sm == bottom_sign) left = mid; else right = mid;
        }
        callback_t = (cb.rootfind == 1) ? left : right;
        residual = b200_get_condition(k, T, callback_t);
    }
    return occurred;
}


// apply_callback!
B200_D void b200_apply_callback(const B200Params& P, long long idx, int k, B200Traj& T, real cb_time, real prev_sign,
                                bool& saved_in_cb) {
    const B200CbInfo cb = B200_CB[k];
#if B200_ADAPTIVE
    T.dtpropose = b200_max(b200_nextfloat(P.dtmin), T.dt);          // set_proposed_dt!(max(nextfloat(dtmin), dtrelax*dt)), dtrelax = 1
#endif
    if (cb_time != T.t) {                                           // change_t_via_interpolation!
        real val[B200_N];
        const real th = (cb_time - T.tprev) / T.dt;
        T.st.interp(th, T.dt, T.uprev, T.u, val);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) T.u[i] = val[i];
        T.t = cb_time;
        T.dt = T.t - T.tprev;
        T.st.addsteps_always(T.uprev, T.p, T.tprev, T.dt);          // reeval_internals_due_to_modification!(continuous)
        T.reeval_fsal = true;
    }
    const bool savedexactly = b200_savevalues(P, idx, T, false);
    saved_in_cb = true;
    if (cb.save_before && !savedexactly) b200_savevalues(P, idx, T, true);
    const bool up = prev_sign < (real)0, down = prev_sign > (real)0;
    if ((up && cb.has_affect) || (down && cb.has_neg)) {
        b200_run_affect(k, down, T);
        T.st.addsteps_always(T.uprev, T.p, T.tprev, T.dt);          // reeval_internals_due_to_modification! once more
        T.reeval_fsal = true;
        if (cb.save_after) { b200_savevalues(P, idx, T, true); saved_in_cb = true; }
    }
}

#endif  // B200_NCC > 0

// handle_callbacks!
B200_D void b200_handle_callbacks(const B200Params& P, long long idx, B200Traj& T) {
    bool saved_in_cb = false;
#if B200_NCC > 0
    {
        bool event_occurred = false; real tmin = T.t, upcrossing = (real)0, residual = (real)0; int identified = 0;
#pragma unroll
        for (int k = 0; k < B200_NCC; ++k) {
            real t2, s2, r2;
            const bool occ2 = b200_find_callback_time(k, k + 1, T, t2, s2, r2);
            if (k == 0) { tmin = t2; upcrossing = s2; residual = r2; event_occurred = occ2; identified = 0; }
            else if (occ2 && (!event_occurred || t2 < tmin)) { tmin = t2; upcrossing = s2; residual = r2; event_occurred = true; identified = k; }
        }
        if (event_occurred) {
            T.last_event_error = residual;
            T.event_last = identified + 1;
#pragma unroll
            for (int k = 0; k < B200_NCC; ++k)
                if (k == identified) b200_apply_callback(P, idx, k, T, tmin, upcrossing, saved_in_cb);
        } else T.event_last = 0;
    }
#endif
#pragma unroll
    for (int k = B200_NCC; k < B200_NCB; ++k) {                     // apply_discrete_callback!, in order
        const B200CbInfo cb = B200_CB[k];
        if (b200_cb_condition(k, T.u, T.p, T.t) != (real)0) {
            const bool savedexactly = b200_savevalues(P, idx, T, false);
            saved_in_cb = true;
            if (cb.save_before && !savedexactly) b200_savevalues(P, idx, T, true);
            if (cb.has_affect) { b200_run_affect(k, false, T); T.reeval_fsal = true; }
            if (cb.save_after) { b200_savevalues(P, idx, T, true); saved_in_cb = true; }
        }
    }
    if (!saved_in_cb) b200_savevalues(P, idx, T, false);
}
