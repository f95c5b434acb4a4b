// b200_composite.cuh — AutoTsit5(Rosenbrock23()): the reference's CompositeAlgorithm((Tsit5(), Rosenbrock23()),
// AutoSwitch(...)) as ONE stepper whose lanes switch between the two methods independently.
//
// Reference behaviour reproduced:
//   AutoTsit5(stiff_alg) = AutoAlgSwitch(Tsit5(), stiff_alg)   lib/OrdinaryDiffEqTsit5/src/algorithms.jl:27-33
//   AutoSwitch defaults, is_stiff, the choice function         lib/OrdinaryDiffEqCore/src/composite_algs.jl:4-97
//   initialize! / perform_step! / choose_algorithm!            lib/OrdinaryDiffEqCore/src/perform_step/composite_perform_step.jl:107-129,168-221
//   eigen_est of a Tsit5 step                                  lib/OrdinaryDiffEqTsit5/src/tsit_perform_step.jl:157-165
//   alg_stability_size(Tsit5) = 3.5068                         lib/OrdinaryDiffEqTsit5/src/alg_utils.jl:3
//   one PI controller cache per branch (CompositeController)   lib/OrdinaryDiffEqCore/src/integrators/controllers.jl:1254-1338
//   "CompositeAlgorithm always recomputes" J and W             lib/OrdinaryDiffEqDifferentiation/src/derivative_utils.jl:107-111
//   eigen_est of a Rosenbrock23 step = opnorm(J, Inf)           lib/OrdinaryDiffEqDifferentiation/src/derivative_utils.jl:996-999 (calc_W)
//
// Lanes of a warp may sit in different branches; the warp then runs both step bodies one after the other (the same
// divergence the reference's per-trajectory `if cache.current == 1` has, paid per warp instead of per thread).
#pragma once
#include "b200_tsit5.cuh"
#include "b200_rosenbrock.cuh"

struct B200AutoTsit5Ros23 {
    B200Tsit5 ns;
    B200Ros23 stf;
    real g6[B200_N];            // stage state of k6 (tsit_perform_step.jl:151), for the stiffness estimate
    real eigen_est;             // integrator.eigen_est, inv(one(tType)) at __init (solve.jl:697)
    int current;                // cache.current: 1 Tsit5, 2 Rosenbrock23
    int count, successive;      // AutoSwitchCache.count, AutoSwitch.successive_switches
    bool do_error_check;        // integrator.do_error_check (composite_algs.jl:37-42, reset in loopfooter!)

    // get_current_alg_order at __init (current = 1) feeds the initial dt
    static B200_D int order() { return 5; }
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1; }
    // controller parameters of the active branch: beta2 = 2//(5 order), beta1 = 7//(10 order) with alg_order 5 / 2
    B200_D real beta1() const { return current == 2 ? (real)(7.0 / 20.0) : (real)(7.0 / 50.0); }
    B200_D real beta2() const { return current == 2 ? (real)(2.0 / 10.0) : (real)(2.0 / 25.0); }
    B200_D real qsteady_max_cur() const { return current == 2 ? (real)1.2 : (real)1; }

    // initialize!(integrator, cache::CompositeCache): current = choice_function(integrator): AS.current == 0 -> 1
    B200_D void init(const real* u, const real* p, real t, int& nf) {
        current = 1; count = 0; successive = 0; do_error_check = true;
        eigen_est = (real)1;
        ns.init(u, p, t, nf);
    }

    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt, real reltol, real abstol,
                        int& nf, int& njacs, int& nw, int& nsolve, bool calck) {
        real EEst = (real)0;
        if (current == 1) {
            EEst = ns.attempt(uprev, u, p, t, dt, reltol, abstol, nf, g6);
            // Hairer II p. 22 with the Inf norm: norm(x, Inf) = mapreduce(abs, max, x), Base.max keeps NaN
            // (the n quotients as one flagged group; u == g6 in a component gives Inf / NaN through the plain operator)
            real q[B200_N];
            bool bad = false;
#pragma unroll
            for (int i = 0; i < B200_N; ++i) q[i] = b200_div_fast(ns.k7[i] - ns.k6[i], u[i] - g6[i], bad);
            if (bad) {
#pragma unroll
                for (int i = 0; i < B200_N; ++i) q[i] = b200_div_cold(ns.k7[i] - ns.k6[i], u[i] - g6[i]);
            }
            real m = b200_abs(q[0]);
#pragma unroll
            for (int i = 1; i < B200_N; ++i) m = b200_max(m, b200_abs(q[i]));
            eigen_est = b200_abs(m);
        } else {
            // calc_W of a CompositeAlgorithm: integrator.eigen_est = opnorm(J, Inf) (derivative_utils.jl:996-999)
            EEst = stf.attempt(uprev, u, p, t, dt, reltol, abstol, nf, njacs, nw, nsolve, calck, &eigen_est);
        }
        return EEst;
    }

    B200_D void accept() { if (current == 1) ns.accept(); else stf.accept(); }
    // reset_fsal! after a callback modified u: fsalfirst = f(u, p, t) in the running branch, nf += 1
    B200_D void reset_fsal(const real* u, const real* p, real t, int& nf) { if (current == 1) ns.init(u, p, t, nf); else stf.init(u, p, t, nf); }
    B200_D void dense_prepare(const real*, const real*, const real*, real, real) {}
    B200_D void interp(real th, real dt, const real* y0, const real* y1, real* out) const {
        if (current == 1) ns.interp(th, dt, y0, y1, out); else stf.interp(th, dt, y0, y1, out);
    }

    // choose_algorithm! with the AutoSwitch choice function; returns true when the branch changed (the caller swaps
    // the controller caches).  dt may be doubled / halved (dtfac = 2).
    B200_D bool choose(real& dt, const real* uprev, const real* p, real t, int& nf) {
        // is_stiff: abs(eigen_est * dt / alg_stability_size(nonstiffalg)); the Float64 constant promotes the quotient
        double stiffness;
        {
            bool bad = false;
            stiffness = fabs(b200_div_fast((double)(eigen_est * dt), 3.5068, bad));
            if (bad) stiffness = fabs(b200_div_cold((double)(eigen_est * dt), 3.5068));
        }
        const bool stiff = !(stiffness <= 0.9);                 // os * tol = 1.0 * 9//10 (both tolerances)
        const bool in_stiff = (current == 2);
        successive = stiff ? 0 : successive + 1;
        do_error_check = (successive > 5) | !stiff | in_stiff;
        count = stiff ? (count < 0 ? 1 : count + 1) : (count > 0 ? -1 : count - 1);
        int next = current;
        if (!in_stiff && count > 10) { dt = dt * (real)2; next = 2; }          // maxstiffstep = 10
        else if (in_stiff && count < -3) { dt = dt / (real)2; next = 1; }      // maxnonstiffstep = 3
        if (next == current) return false;
        current = next;
        // initialize!(integrator, caches[new]): fsalfirst = f(uprev, p, t); nf += 1 (both constant caches)
        if (current == 1) ns.init(uprev, p, t, nf); else stf.init(uprev, p, t, nf);
        return true;
    }
};
