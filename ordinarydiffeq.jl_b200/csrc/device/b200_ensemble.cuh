// b200_ensemble.cuh — the per-trajectory integrator loop and the two kernels of
// the ensemble hot path.
//
//   b200_initdt     one thread per trajectory, no divergence: Hairer initial step
//                   (lib/OrdinaryDiffEqCore/src/initdt.jl:346-459, OOP form).
//   b200_integrate  persistent warps; every lane owns one trajectory at a time and
//                   pulls the next index from a warp-local pool that is refilled
//                   from one global counter (guided chunking), so lanes stay busy
//                   although trajectories take different numbers of adaptive steps.
//
// The loop body follows the reference's solve! / loopheader! / check_error /
// perform_step! / loopfooter! / savevalues! sequence
//   lib/OrdinaryDiffEqCore/src/solve.jl:904-946
//   lib/OrdinaryDiffEqCore/src/integrators/integrator_utils.jl:84-127,130-150,175-203,
//       268-333,340-414,597-677,1027-1034,1199-1256
//   lib/OrdinaryDiffEqCore/src/integrators/controllers.jl:245-250,288-293,805-843
//   lib/DiffEqBase/src/check_error.jl:70-118
// written for forward time (reverse-time programs run the mirrored problem through the same code: B200_REVERSE below);
// adaptive PI control; the variants (fixed steps, tstops, ragged rows, callbacks, ...) are compile-time options.
//
// Compile-time configuration (set by the shim before this file):
//   B200_N, B200_NP      state / parameter dimension
//   B200_F32             0: double, 1: float
//   B200_ALG             1 Tsit5, 2 Vern7, 3 Rosenbrock23, 4 Rodas5P, 5 DP5, 6 BS3,
//                        7 Rodas5, 8 Rodas4, 9 Rodas42, 10 Rodas4P, 11 Rodas4P2, 12 Vern6, 13 Vern8, 14 Vern9, 15 Rosenbrock32, 16 Rodas5Pe,
//                        17 AutoTsit5(Rosenbrock23()), 18 Rodas3P, 19 Rodas23W
//   B200_RHS(du,u,p,t)   user right-hand side (plus B200_JAC / B200_TGRAD for stiff)
//   B200_BLOCK, B200_MINBLOCKS   launch bounds
#pragma once
#include "b200_base.cuh"
#include "b200_detmath.cuh"

#define B200_ALG_TSIT5 1
#define B200_ALG_VERN7 2
#define B200_ALG_ROS23 3
#define B200_ALG_RODAS5P 4
#define B200_ALG_DP5 5
#define B200_ALG_BS3 6
#define B200_ALG_RODAS5 7
#define B200_ALG_RODAS4 8
#define B200_ALG_RODAS42 9
#define B200_ALG_RODAS4P 10
#define B200_ALG_RODAS4P2 11
#define B200_ALG_VERN6 12
#define B200_ALG_VERN8 13
#define B200_ALG_VERN9 14
#define B200_ALG_ROS32 15
#define B200_ALG_RODAS5PE 16
#define B200_ALG_AUTOTSIT5_ROS23 17
#define B200_ALG_RODAS3P 18
#define B200_ALG_RODAS23W 19
#define B200_COMPOSITE (B200_ALG == B200_ALG_AUTOTSIT5_ROS23)
#define B200_IS_RODAS (B200_ALG == B200_ALG_RODAS5P || B200_ALG == B200_ALG_RODAS5PE || B200_ALG == B200_ALG_RODAS3P || B200_ALG == B200_ALG_RODAS23W || (B200_ALG >= B200_ALG_RODAS5 && B200_ALG <= B200_ALG_RODAS4P2))
// (the composite algorithm counts as Rosenbrock-type here: it needs jac/tgrad and reports njacs / nw / nsolve)
#define B200_IS_ROSENBROCK (B200_ALG == B200_ALG_ROS23 || B200_ALG == B200_ALG_ROS32 || B200_IS_RODAS || B200_COMPOSITE)

#ifndef B200_COOP
#define B200_COOP 0           // 1: lane-group kernel (b200_coop.cuh): B200_L lanes per trajectory, component-form RHS
#endif

#ifndef B200_WIDE
#define B200_WIDE 0           // 1: stage derivatives in shared memory (b200_vern7_wide.cuh); B200_WIDE_NT threads per CTA own a trajectory
#endif
#if B200_WIDE && (B200_COOP || B200_ALG != B200_ALG_VERN7)
#error "the shared-memory stage kernel (B200ODE_OPT_SMEM_STAGES) is available for Vern7, one thread per trajectory"
#endif

#if B200_COOP
// defined below, after B200Params (b200_coop.cuh)
// the one-thread-per-trajectory initial-dt kernel evaluates the component form in a loop
#define B200_USER_RHS(du, u, p, t) do { for (int b200_i = 0; b200_i < B200_N; ++b200_i) (du)[b200_i] = B200_USER_RHS_COMP(b200_i, (u), (p), (t)); } while (0)
#elif B200_ALG == B200_ALG_TSIT5
#define B200_RHS(du, u, p, t) B200_USER_RHS(du, u, p, t)
#include "b200_tsit5.cuh"
typedef B200Tsit5 B200Stepper;
#elif B200_ALG == B200_ALG_VERN7 && B200_WIDE
// stage derivatives in shared memory, one inlined RHS (b200_vern7_wide.cuh): wide states
#include "b200_vern7_wide.cuh"
typedef B200Vern7Wide B200Stepper;
#elif B200_ALG == B200_ALG_VERN7
#define B200_RHS(du, u, p, t) B200_USER_RHS(du, u, p, t)
#include "b200_vern7.cuh"
typedef B200Vern7 B200Stepper;
#elif B200_COMPOSITE
#define B200_RHS(du, u, p, t) B200_USER_RHS(du, u, p, t)
#include "b200_composite.cuh"
typedef B200AutoTsit5Ros23 B200Stepper;
#elif B200_IS_ROSENBROCK
#define B200_RHS(du, u, p, t) B200_USER_RHS(du, u, p, t)
#include "b200_rosenbrock.cuh"
#if B200_ALG == B200_ALG_ROS23 || B200_ALG == B200_ALG_ROS32
typedef B200Ros23 B200Stepper;
#else
typedef B200Rodas5P B200Stepper;
#endif
#elif B200_ALG == B200_ALG_VERN6 || B200_ALG == B200_ALG_VERN8 || B200_ALG == B200_ALG_VERN9
#define B200_RHS(du, u, p, t) B200_USER_RHS(du, u, p, t)
#include "b200_verner_gen.cuh"
#if B200_ALG == B200_ALG_VERN6
typedef B200Vern6 B200Stepper;
#elif B200_ALG == B200_ALG_VERN8
typedef B200Vern8 B200Stepper;
#else
typedef B200Vern9 B200Stepper;
#endif
#elif B200_ALG == B200_ALG_DP5 || B200_ALG == B200_ALG_BS3
#define B200_RHS(du, u, p, t) B200_USER_RHS(du, u, p, t)
#include "b200_lowrk.cuh"
#if B200_ALG == B200_ALG_DP5
typedef B200DP5 B200Stepper;
#else
typedef B200BS3 B200Stepper;
#endif
#endif

// PI controller exponents: beta2_default = 2//(5 order), beta1_default = 7//(10 order)
// (OrdinaryDiffEqCore alg_utils.jl:766,788), overridden for DP5 (LowOrderRK alg_utils.jl:37-39);
// QT(rational) is the correctly rounded quotient
#if B200_ALG == B200_ALG_DP5
#define B200_BETA2 ((real)(4.0 / 100.0))
#define B200_BETA1 ((real)(17.0 / 100.0))
#else
#define B200_BETA2 ((real)(2.0 / (5.0 * B200Stepper::order())))
#define B200_BETA1 ((real)(7.0 / (10.0 * B200Stepper::order())))
#endif

// ReturnCode values exported to the host (include/b200ode.h)
#define B200_RC_DEFAULT 0
#define B200_RC_SUCCESS 1
#define B200_RC_MAXITERS 2
#define B200_RC_DTLESSTHANMIN 3
#define B200_RC_UNSTABLE 4
#define B200_RC_DTNAN 5
#define B200_RC_TERMINATED 6      // terminate!(integrator) from a callback

struct B200Params {
    long long N;              // trajectories handled by this launch
    const real* u0;           // element (i,c) at u0[i*u0_ts + c*u0_cs]; u0_ts = 0 when shared
    long long u0_ts, u0_cs;
    const real* p;
    long long p_ts, p_cs;
    real t0, tf;
    real reltol, abstol;
    real dt_user;             // 0 => automatic (dt0[] filled by b200_initdt)
    real dtmin, dtmax;
    long long maxiters;
    const real* saveat;       // grid times in (t0, tf], ascending
    int nsaveat;
    int save_start, save_end;
    int nslots;               // rows per trajectory in us (0 => no time series output)
    real* dt0;                // [N]
    real* u_final;            // element (i,c) at u_final[i*uf_ts + c*uf_cs]
    long long uf_ts, uf_cs;
    real* t_final;            // [N]
    real* us;                 // [N][nslots][n]
    int* naccept; int* nreject; int* nf; int* retcode; int* nsaved;
    int* njacs; int* nw; int* nsolve;
    unsigned long long* work_counter;
    int flags;
    int tol_const;            // |t0| <= |tf|: the tstop tolerance 100*eps(max(|t|,|tf|)) is the constant below
    real tol100_tf;           // 100*eps(|tf|)
    // save_everystep programs (-DB200_EVERYSTEP=1) only: ragged per-step output.  Trajectory i owns rows
    // row_offsets[i] .. row_offsets[i+1]-1 of us[.][n] / ts_rag[.]; NULL = counting pass (nsaved[] only).
    const long long* row_offsets;
    real* ts_rag;
    real* dts_rag;            // step size the stages of the step ending at each row were computed with (0 for other rows)
    // tstops programs (-DB200_TSTOPS=1) only: opts.tstops as initialize_tstops builds it (solve.jl:1021-1040):
    // ascending, strictly inside (t0, tf), duplicates kept, tf appended last
    const real* tstops;
    int ntstops;
    // fastpower(qoldinit = 1e-4, beta2) and its correctly rounded reciprocal: the controller state every trajectory
    // starts from (setup_controller_cache, controllers.jl:793-803), computed once by the host
    real fpe0, rfpe0;
    // d_discontinuities (tstops programs only): opts.d_discontinuities as reinit_d_discontinuities! builds it
    // (solve.jl:1185-1197): every entry >= t0, ascending; the entries inside (t0, tf) are also stops of `tstops`
    const real* disc;
    int ndisc;
    // per-trajectory time spans (programs compiled with -DB200_TSPANS=1): (t0_i, tf_i) pairs, NULL otherwise
    const double* tspans;
    int dtmax_default;        // 1: opts.dtmax was not given, i.e. dtmax = tf_i - t0_i per trajectory (solve.jl:152)
    // fused ordered gather (B200DeviceResult.peer_u_final): the final state also goes, at the trajectory's global index,
    // into the result arrays of npeer ranks (this GPU's own and its NVLink peers')
    real* peer_out[8];
    int npeer, peer_world, peer_rank;
    long long peer_block;
};

#define B200_FLAG_STATIC_SCHEDULE 1   // one trajectory per thread, no refill (A/B baseline)
#define B200_FLAG_NO_STEP_ROWS 2      // save_everystep programs: ragged rows without the per-step rows (saveat + callback rows)

// save_idxs (solve.jl kwarg; _savevalues! integrator_utils.jl:368-375, ode_interpolant(Θ, integrator, idxs, ...)):
// -DB200_SAVE_IDXS=i0,i1,... (0-based) makes every saved row hold only those components, in that order.
// A compile-time list keeps the state in registers (no dynamic indexing of u[]).
#ifdef B200_SAVE_IDXS
template <int... I> struct B200IdxList { static constexpr int n = (int)sizeof...(I); };
#define B200_SAVE_TAB_DECL constexpr int b200_save_tab[] = {B200_SAVE_IDXS};
#define B200_NSAVE (B200IdxList<B200_SAVE_IDXS>::n)
#define B200_SAVE_COMP(c) b200_save_tab[c]
#else
#define B200_SAVE_TAB_DECL
#define B200_NSAVE B200_N
#define B200_SAVE_COMP(c) (c)
#endif

#ifndef B200_ADAPTIVE
#define B200_ADAPTIVE 1       // 0: adaptive = false — fixed dt = opts.dt (dtcache), every step accepted, no controller
#endif
#if !B200_ADAPTIVE && B200_COOP
#error "adaptive=false is not available in the lane-group kernel"
#endif

#ifndef B200_TSTOPS
#define B200_TSTOPS 0         // 1: the tstops keyword (several stop times); 0: tstops = {tf}
#endif

// B200_REVERSE = 1 (B200ODE_OPT_REVERSE_TIME): tspan[2] < tspan[1], tdir = -1 (solve.jl:273).  The kernels integrate the
// mirrored problem du/ds = -f(u, p, -s) over (-t0, -tf): every operation of the integrator is an IEEE operation that
// commutes with negation (round-to-nearest is symmetric), the reference's tdir bookkeeping multiplies both sides of each
// comparison by tdir, so the mirrored run reproduces the reverse run bit for bit — states, stage derivatives (negated),
// step sizes (negated), statistics.  The shim wraps the user's functions (RHS and Jacobian negated, every user function
// sees t = -s: B200_USER_T) and hands the kernels mirrored times; the kernels store times through B200_USER_T.  The two
// places where the reference itself is not symmetric in tdir (fix_dt_at_bounds!'s dtmin clamp, check_error's tstop
// comparison) are compiled as the reference writes them for tdir < 0 (see b200_traj_iterate).
#if B200_REVERSE && (B200_COOP || B200_WIDE)
#error "reverse-time integration (B200ODE_OPT_REVERSE_TIME) is available in the one-thread-per-trajectory kernel"
#endif
#if B200_TSTOPS && B200_COOP
#error "tstops are not available in the lane-group kernel"
#endif

#ifndef B200_EVERYSTEP
#define B200_EVERYSTEP 0      // 1: save_everystep = true (integrator_utils.jl:385-411), ragged rows
#endif
#ifndef B200_CALLBACKS
#define B200_CALLBACKS 0      // 1: the program carries a CallbackSet (device/b200_callbacks.cuh; Tsit5)
#endif
#ifndef B200_TSPANS
#define B200_TSPANS 0         // 1: every trajectory has its own (t0, tf) (prob_func changed tspan): B200Params.tspans
#endif
#if B200_TSPANS && (B200_COOP || B200_TSTOPS || B200_CALLBACKS || B200_WIDE)
#error "per-trajectory time spans are not combined with tstops / d_discontinuities, callbacks, the lane-group or the stage kernel"
#endif
#if B200_TSPANS
#define B200_T0 (T.t0)
#define B200_TF (T.tf)
#define B200_DTMAX (T.dtmax)
#else
#define B200_T0 (P.t0)
#define B200_TF (P.tf)
#define B200_DTMAX (P.dtmax)
#endif
#ifndef B200_STAGE_ROWS
#define B200_STAGE_ROWS 0     // 1: saveat rows are packed through a per-warp shared-memory queue and interpolated at full lane occupancy (Tsit5)
#endif
#if B200_STAGE_ROWS && (B200_ALG != B200_ALG_TSIT5 || B200_EVERYSTEP || B200_CALLBACKS || defined(B200_SAVE_IDXS) || B200_COOP || B200_TSPANS)
#error "the staged saveat queue serves Tsit5 with a rectangular saveat output (no save_everystep / save_idxs / callbacks)"
#endif
#if defined(B200_ISOUT) && B200_COOP
#error "isoutofdomain is not available in the lane-group kernel"
#endif
#if B200_VECTOR_TOL && B200_COOP
#error "per-component tolerances are not available in the lane-group kernel"
#endif
#if B200_CALLBACKS && B200_COOP
#error "callbacks are not available in the lane-group kernel"
#endif
#if B200_WIDE && (B200_EVERYSTEP || B200_CALLBACKS)
#error "save_everystep / callbacks are not available in the shared-memory stage kernel (no interpolant: k11..k16 are not stored)"
#endif
#if B200_CALLBACKS && B200_NCC > 0 && B200_ALG != B200_ALG_TSIT5
#error "continuous callbacks are available for Tsit5 (discrete callbacks: every stepper)"
#endif
#if (B200_EVERYSTEP || defined(B200_SAVE_IDXS)) && B200_COOP
#error "save_everystep / save_idxs are not available in the lane-group kernel"
#endif

#ifndef B200_BLOCK
#define B200_BLOCK 128
#endif
#ifndef B200_MINBLOCKS
#define B200_MINBLOCKS 1
#endif

#if B200_COOP
#if B200_ALG != B200_ALG_VERN7 && B200_ALG != B200_ALG_ROS23
#error "the lane-group kernel is available for Vern7 and Rosenbrock23"
#endif
#if B200_ALG == B200_ALG_ROS23
struct B200CoopStepperTag { static B200_D int order() { return 2; } static B200_D real qsteady_min() { return (real)1; } static B200_D real qsteady_max() { return (real)1.2; } };
#else
struct B200CoopStepperTag { static B200_D int order() { return 7; } static B200_D real qsteady_min() { return (real)1; } static B200_D real qsteady_max() { return (real)1; } };
#endif
typedef B200CoopStepperTag B200Stepper;       // order / qsteady for the controller helpers (the stepper itself follows)
#endif

// ---------------------------------------------------------------------------
// ODE_DEFAULT_NORM for a static vector: sqrt(sum(abs2,u)/n), left fold, no fusion
// (lib/DiffEqBase/src/common_defaults.jl:102-107).
B200_D real b200_rms(const real* v) {
    real acc = v[0] * v[0];
#pragma unroll
    for (int i = 1; i < B200_N; ++i) acc = acc + v[i] * v[i];
    return b200_sqrt(b200_div_const(acc, (real)B200_N, (real)1 / (real)B200_N));
}

// _ode_initdt_oop (initdt.jl:346-459), g === nothing, forward time.
B200_D real b200_initdt_one(const real* u0, const real* p, real t, real dtmax_tdir, real abstol, real reltol,
                            real opts_dtmin, int order) {
    real dtmin = b200_nextfloat(b200_max(opts_dtmin, b200_eps(t)));
    real smalldt = b200_max(dtmin, (real)1e-6);
    real sk[B200_N], f0[B200_N], tmp[B200_N];
#pragma unroll
    for (int i = 0; i < B200_N; ++i) sk[i] = b200_fma(b200_abs(u0[i]), B200_RTOL_AT(i, reltol), B200_ATOL_AT(i, abstol));
#pragma unroll
    for (int i = 0; i < B200_N; ++i) tmp[i] = u0[i] / sk[i];
    real d0 = b200_rms(tmp);
    B200_USER_RHS(f0, u0, p, t);
    bool anynan = false;
#pragma unroll
    for (int i = 0; i < B200_N; ++i) anynan = anynan || b200_isnan(f0[i]);
    if (anynan) return dtmin;
#pragma unroll
    for (int i = 0; i < B200_N; ++i) tmp[i] = f0[i] / sk[i];
    real d1 = b200_rms(tmp);
    if (b200_isnan(d1)) return dtmin;
    real dt0;
    // comparisons against the exact rationals 1//10^5 (both binary64 and binary32
    // operands, widened exactly to double, compare like this against the double literal)
    if ((double)d0 < 1e-5 || (double)d1 < 1e-5) dt0 = smalldt;
    else dt0 = (d0 / d1) / (real)100;
    dt0 = b200_min(dt0, dtmax_tdir);
    real u1[B200_N], f1[B200_N];
#pragma unroll
    for (int i = 0; i < B200_N; ++i) u1[i] = b200_fma(dt0, f0[i], u0[i]);
    B200_USER_RHS(f1, u1, p, t + dt0);
    bool alleq = true;
#pragma unroll
    for (int i = 0; i < B200_N; ++i) alleq = alleq && (f0[i] == f1[i]);
    if (alleq) return b200_max(dtmin, (real)100 * dt0);
#pragma unroll
    for (int i = 0; i < B200_N; ++i) tmp[i] = (f1[i] - f0[i]) / sk[i];
    real d2 = b200_rms(tmp) / dt0;
    real m = b200_max(d1, d2);
    real dt1;
    if ((double)m < 1e-15) {            // m <= 1//10^15 (exact rational)
        dt1 = b200_max(smalldt, dt0 * (real)0.001);
    } else if (!b200_isfinite(m)) {
        dt1 = (real)0;                  // log10(Inf)=Inf -> 10^-Inf = 0
    } else {
        real l = (real)b200_log10_cr((double)m);
        real e = -((real)2 + l) / (real)order;
        dt1 = (real)b200_exp10_cr((double)e);
    }
    return b200_max(dtmin, b200_min(b200_min((real)100 * dt0, dt1), dtmax_tdir));
}

extern "C" __global__ void __launch_bounds__(256) b200_initdt(B200Params P) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    real u0[B200_N], p[B200_NP > 0 ? B200_NP : 1];
#pragma unroll
    for (int c = 0; c < B200_N; ++c) u0[c] = P.u0[i * P.u0_ts + c * P.u0_cs];
#pragma unroll
    for (int c = 0; c < B200_NP; ++c) p[c] = P.p[i * P.p_ts + c * P.p_cs];
    // _determine_initdt: dtmax = min(|opts.dtmax|, |first_tstop - t|)
    // _determine_initdt: dtmax = min(|opts.dtmax|, |first_tstop - t|) (integrator_interface.jl:643-647)
#if B200_TSPANS
    const real t0i = (real)P.tspans[2 * i], tfi = (real)P.tspans[2 * i + 1];
    const real dtmax_o = P.dtmax_default ? (tfi - t0i) : P.dtmax;
    real dtmax = b200_min(b200_abs(dtmax_o), b200_abs(tfi - t0i));
    P.dt0[i] = b200_initdt_one(u0, p, t0i, dtmax, P.abstol, P.reltol, P.dtmin, B200Stepper::order());
#else
#if B200_TSTOPS
    real dtmax = b200_min(b200_abs(P.dtmax), b200_abs(P.tstops[0] - P.t0));
#else
    real dtmax = b200_min(b200_abs(P.dtmax), b200_abs(P.tf - P.t0));
#endif
    P.dt0[i] = b200_initdt_one(u0, p, P.t0, dtmax, P.abstol, P.reltol, P.dtmin, B200Stepper::order());
#endif
}

#if B200_ADAPTIVE
// ---- stepsize_controller! / step_accept_controller! / step_reject_controller! of the PI controller -------------
// (lib/OrdinaryDiffEqCore/src/integrators/controllers.jl:805-843), all scalar work of one loopfooter! in one
// straight-line block:
//   q11 = fastpower(EEst, beta1); q = q11 / fastpower(errold, beta2) / gamma, clamped to [1/qmax, 1/qmin]
//   accept:  (qsteady) q = 1;  errold = max(EEst, 1e-4);  dtnew = dt / q
//   reject:  dt = dt / min(1/qmin, q11/gamma)            (the reference does this in the next loopheader!)
// Accepted and rejected steps need ONE true division (numerator/denominator selected first), and the two
// fastpower calls share fastlog2(Float32(EEst)) whenever errold == EEst.  FAST = flagged branch-free math
// (b200_base.cuh); the caller repeats the block with FAST = false if the flag comes back set.
struct B200Ctl { real q11, fpe, rfpe, dtdiv, num; bool accept; };
// PI parameters that depend on the algorithm (compile-time constants except for a composite algorithm, whose
// branches bring their own: CompositeController, controllers.jl:1254-1338)
struct B200CtlCfg { real beta1, beta2, qsteady_min, qsteady_max; };
B200_D B200CtlCfg b200_ctl_cfg_static() {
    B200CtlCfg c; c.beta1 = B200_BETA1; c.beta2 = B200_BETA2;
    c.qsteady_min = B200Stepper::qsteady_min(); c.qsteady_max = B200Stepper::qsteady_max();
    return c;
}
template <bool FAST>
B200_D B200Ctl b200_controller_t(real EEst, real q11_old, real fpe, real rfpe, real dt, real dtpropose, bool tstop_flag,
                                 bool first, bool& bad, const B200CtlCfg cfg, bool isout = false) {
    const real qmin = (real)0.2, qmax = (real)10, gamma = (real)0.9;
    const real beta1 = cfg.beta1, beta2 = cfg.beta2;
    B200Ctl c;
    const real qmax_eff = first ? (real)10000 : qmax;
    // fastpower(EEst, beta1) and fastpower(max(EEst, 1e-4), beta2) from one logarithm
#if B200_F32
    const float ef = EEst;
#else
    const float ef = (float)EEst;
#endif
    const float lg = b200_fastlog2(ef);
    const bool small = ((real)1e-4 > EEst);            // errold = max_c(1e-4, EEst) picks the constant
    const float lg_old = small ? b200_fastlog2((float)(real)1e-4) : lg;
    real q11 = (real)b200_exp2_fast((float)beta1 * lg);
    q11 = (EEst == (real)0) ? (real)0 : q11;           // fastpower(0, y) = 0
    c.fpe = (real)b200_exp2_fast((float)beta2 * lg_old);
    c.rfpe = B200Math<FAST>::div((real)1, c.fpe, bad);
    real q = B200Math<FAST>::divc(q11, fpe, rfpe, bad);
    q = B200Math<FAST>::divc(q, gamma, (real)1 / gamma, bad);
    const real lo = (real)1 / qmax_eff, hi = (real)1 / qmin;
    q = q < lo ? lo : (q > hi ? hi : q);               // clamp under @fastmath
    const bool zero = (EEst == (real)0);               // iszero(EEst): q = inv(qmax), q11 untouched
    q = zero ? lo : q;
    c.q11 = zero ? q11_old : q11;
    // accept_step = !isout && accept_step_controller (integrator_utils.jl:612-618); isout = isoutofdomain(u, p, t + dt)
    c.accept = (EEst <= (real)1) & !isout;
    // step_accept_controller!: qsteady window
    const real qa = (cfg.qsteady_min <= q && q <= cfg.qsteady_max) ? (real)1 : q;
    // step_reject_controller!: dt /= min(inv(qmin), q11/gamma)
    const real qr = b200_min_c((real)1 / qmin, B200Math<FAST>::divc(c.q11, gamma, (real)1 / gamma, bad));
    // accepted steps divide the un-clipped dt (integrator_utils.jl:629-633 restores it first)
    c.num = (c.accept && tstop_flag) ? dtpropose : dt;
    c.dtdiv = B200Math<FAST>::div(c.num, c.accept ? qa : qr, bad);
    // handle_step_rejection! (integrator_utils.jl:135-136): a step rejected by isoutofdomain shrinks dt by qmin
    c.dtdiv = isout ? dt * qmin : c.dtdiv;
    return c;
}
#endif

#if B200_COOP
#include "b200_coop.cuh"
#endif

#if !B200_COOP
// ---------------------------------------------------------------------------
struct B200Traj {
    real u[B200_N], uprev[B200_N];
    real p[B200_NP > 0 ? B200_NP : 1];
    B200Stepper st;
    real t, tprev, dt, dtpropose;
    real q11, EEst;
    real fpe, rfpe;             // fastpower(errold, beta2) and its correctly rounded reciprocal (see iterate)
#if B200_COMPOSITE
    real q11_o, fpe_o, rfpe_o;  // the PI cache of the branch that is not running (swapped in when the algorithm switches)
#endif
    real next_save;             // saveat[save_idx] (or +Inf when the grid is exhausted)
    int naccept, nreject, nf;      // iter = naccept+nreject(+1), success_iter = naccept (see iterate)
    int save_idx, nsaved;
    int retcode;
    bool accept, tstop_flag;
    real* row;                  // next row of us[idx][.][:] (running pointer: no 64-bit index arithmetic per row)
#if B200_TSPANS
    real t0, tf, dtmax;         // this trajectory's span and opts.dtmax (default tf - t0)
#endif
#if B200_TSTOPS
    real tstop;                 // first(opts.tstops)
    int tstop_idx;
    int disc_idx;               // entries of opts.d_discontinuities already popped
#endif
#if B200_EVERYSTEP
    real* trow;                 // next entry of ts_rag
    real* drow;                 // next entry of dts_rag
    real last_t;                // sol.t[end]
    int cap;                    // rows this trajectory owns (0 in the counting pass)
#endif
#if B200_IS_ROSENBROCK
    int njacs, nw, nsolve;
#endif
#if B200_CALLBACKS
    int event_last;             // integrator.event_last_time: 1-based index of the continuous callback that ended the last step
    real last_event_error;      // integrator.last_event_error
    bool reeval_fsal, terminated;
#endif
};

B200_D void b200_emit(const B200Params& P, long long idx, B200Traj& T, real ts, const real* v, real dt_stages = (real)0) {
    B200_SAVE_TAB_DECL
#if B200_EVERYSTEP
    if (T.nsaved < T.cap) {
#pragma unroll
        for (int c = 0; c < B200_NSAVE; ++c) T.row[c] = v[B200_SAVE_COMP(c)];
        T.row += B200_NSAVE;
        *T.trow++ = B200_USER_T(ts);                 // reverse-time programs run in mirrored time (B200_REVERSE)
        *T.drow++ = B200_USER_T(dt_stages);
    }
    T.last_t = ts;
#else
    if (T.nsaved < P.nslots) {              // nslots == 0: no time series requested
#pragma unroll
        for (int c = 0; c < B200_NSAVE; ++c) T.row[c] = v[B200_SAVE_COMP(c)];
        T.row += B200_NSAVE;
    }
#endif
    T.nsaved += 1;
}


// modify_dt_for_tstops! (integrator_utils.jl:268-324), adaptive branch, tstops={tf}.
// tol100 = 100*eps(max(|t|,|tf|)) and dist = |tf - t| depend only on t.
#if B200_ADAPTIVE
B200_D void b200_modify_dt_for_tstops(const B200Params&, B200Traj& T, real dist, real tol100) {
    real orig = b200_abs(T.dt);
    T.dtpropose = orig;
    T.tstop_flag = !(orig + tol100 < dist);
    T.dt = b200_min_c(dist, orig);
}
#else
// non-adaptive branches (integrator_utils.jl:300-316; dtchangeable): always step with dtcache = opts.dt,
// shortened to the next stop; dtcache == 0 steps from stop to stop
B200_D void b200_modify_dt_for_tstops(const B200Params& P, B200Traj& T, real dist, real tol100) {
    const real dtcache = b200_abs(P.dt_user);
    if (dtcache == (real)0) { T.dt = dist; T.tstop_flag = true; return; }
    T.tstop_flag = !(dtcache + tol100 < dist);
    T.dt = b200_min_c(dist, dtcache);
}
#endif

B200_D void b200_traj_begin(const B200Params& P, long long idx, B200Traj& T) {
#pragma unroll
    for (int c = 0; c < B200_N; ++c) {
        real v = P.u0[idx * P.u0_ts + c * P.u0_cs];
        T.u[c] = v; T.uprev[c] = v;
    }
#pragma unroll
    for (int c = 0; c < B200_NP; ++c) T.p[c] = P.p[idx * P.p_ts + c * P.p_cs];
#if B200_TSPANS
    T.t0 = (real)P.tspans[2 * idx]; T.tf = (real)P.tspans[2 * idx + 1];
    T.dtmax = P.dtmax_default ? (T.tf - T.t0) : P.dtmax;
#endif
    T.t = B200_T0; T.tprev = B200_T0;
    T.nf = 0;
#if B200_IS_ROSENBROCK
    T.njacs = 0; T.nw = 0; T.nsolve = 0;
#endif
    T.nsaved = 0; T.save_idx = 0;
#if B200_EVERYSTEP
    T.cap = 0; T.row = nullptr; T.trow = nullptr; T.drow = nullptr; T.last_t = B200_T0;
    if (P.row_offsets != nullptr) {
        const long long o = P.row_offsets[idx];
        T.cap = (int)(P.row_offsets[idx + 1] - o);
        T.row = P.us + (size_t)o * B200_NSAVE;
        T.trow = P.ts_rag + o;
        T.drow = P.dts_rag + o;
    }
#else
    T.row = P.us + (size_t)idx * (size_t)P.nslots * B200_NSAVE;
#endif
    if (P.save_start) b200_emit(P, idx, T, T.t, T.u);      // solve.jl:809-824
    T.st.init(T.u, T.p, T.t, T.nf);                   // initialize!(integrator, cache)
#if B200_ADAPTIVE
    if (P.dt_user == (real)0) { T.dt = P.dt0[idx]; T.nf += 2; }   // auto_dt_reset!: nf += 2
    else T.dt = P.dt_user;
#else
    T.dt = P.dt_user;          // handle_dt!: the automatic initial dt is adaptive-only (solve.jl:968-985)
#endif
    T.dtpropose = T.dt;
    T.q11 = (real)1; T.EEst = (real)1;                            // setup_controller_cache (controllers.jl:793-803)
    T.fpe = P.fpe0; T.rfpe = P.rfpe0;       // errold = qoldinit = 1e-4; only fastpower(errold, beta2) is ever used
#if B200_COMPOSITE
    // the stiff branch's own PIControllerCache (q11 = 1, errold = qoldinit, beta2 = 2//(5*2))
    T.q11_o = (real)1;
    T.fpe_o = b200_fastpower((real)1e-4, (real)(2.0 / 10.0));
    T.rfpe_o = (real)1 / T.fpe_o;
#endif
    T.next_save = (P.nsaveat > 0) ? P.saveat[0] : b200_inf();
#if B200_TSPANS
    // the shared list holds absolute times: this trajectory's grid is its part inside (t0_i, tf_i]
    while (T.next_save <= T.t0) {
        T.save_idx += 1;
        T.next_save = (T.save_idx < P.nsaveat) ? P.saveat[T.save_idx] : b200_inf();
    }
#endif
#if B200_TSTOPS
    T.tstop_idx = 0; T.tstop = P.tstops[0];
    T.disc_idx = 0;
    // handle_starting_time_discontinuity! (solve.jl:887-901): a discontinuity at exactly t0 is popped, t moves one ulp
    // forward and a first-same-as-last stepper evaluates its first stage again on the new side (reset_fsal!: nf += 1).
    // The initial dt was determined at t0 itself, the start row carries t0.
    if (P.ndisc > 0 && P.disc[0] == T.t) {
        T.disc_idx = 1;
        T.t = b200_nextfloat_signed(T.t);
#if B200_COMPOSITE
        T.st.reset_fsal(T.u, T.p, T.t, T.nf);
#else
        T.st.init(T.u, T.p, T.t, T.nf);
#endif
    }
#endif
    T.naccept = 0; T.nreject = 0;
    T.accept = false; T.tstop_flag = false;
    T.retcode = B200_RC_DEFAULT;
#if B200_CALLBACKS
    T.event_last = 0; T.last_event_error = (real)0; T.reeval_fsal = false; T.terminated = false;
#endif
}

#if B200_CALLBACKS
#include "b200_callbacks.cuh"
#endif

// One pass of the while-loop body of solve!.  Returns true when the trajectory is finished.
// `amask` = the lanes of this warp that execute this call; the short accept/reject
// branches of loopheader! are re-converged explicitly before the stage loop (otherwise
// a single rejecting lane makes the warp run perform_step! twice).
B200_D bool b200_traj_iterate(const B200Params& P, long long idx, B200Traj& T, unsigned amask) {
    const real qmin = (real)0.2, qmax = (real)10, gamma = (real)0.9;
    const real beta1 = B200_BETA1;
    const real beta2 = B200_BETA2;
    // quantities of modify_dt_for_tstops! that depend only on t (t is finite here)
    // integrator.iter (before its increment) and integrator.success_iter are not stored:
    // iter = naccept + nreject and success_iter = naccept at every point they are read.
    const int iter0 = T.naccept + T.nreject;
#if B200_TSTOPS
    // update_fsal! (integrator_utils.jl:215-220), first branch: the step just accepted ended on the first entry of
    // opts.d_discontinuities — pop it, move t one ulp past it (shift_past_discontinuity!) and, below, refresh the first
    // stage of a first-same-as-last stepper at the new t instead of copying the last one.  (t does not enter update_uprev!
    // or `dt = dtpropose`, so shifting before them is the reference's order of effects.)
    bool disc_hit = false;
    if (P.ndisc > 0 && iter0 > 0 && T.accept && T.disc_idx < P.ndisc && P.disc[T.disc_idx] == T.t) {
        T.disc_idx += 1;
        T.t = b200_nextfloat_signed(T.t);
        disc_hit = true;
    }
    const real tstop = T.tstop;
    const real dist = b200_abs(tstop - T.t);
    const real at = b200_abs(T.t), atf = b200_abs(tstop);
    const real tol100 = (real)100 * b200_eps_finite(at > atf ? at : atf);
#else
    const real tstop = B200_TF;
    const real dist = b200_abs(B200_TF - T.t);
    const real at = b200_abs(T.t), atf = b200_abs(B200_TF);
    // tstop tolerance 100*eps(max(|t|,|tf|)): a launch constant whenever |t0| <= |tf|
    const real tol100 = P.tol_const ? P.tol100_tf : (real)100 * b200_eps_finite(at > atf ? at : atf);
#endif
    const real eps_t = b200_eps_finite(T.t);
    const real dtmin_t = eps_t > P.dtmin ? eps_t : P.dtmin;          // timedepentdtmin
    // ---- loopheader! ----
    if (iter0 > 0) {
        if (T.accept) {
            // apply_step!: update_uprev!, dt = dtpropose, update_fsal!, modify_dt_for_tstops!
#pragma unroll
            for (int c = 0; c < B200_N; ++c) T.uprev[c] = T.u[c];
            T.dt = T.dtpropose;
#if B200_TSTOPS
            if (disc_hit) {         // get_current_isfsal && reset_fsal! (init() of a stepper that is not FSAL is empty)
#if B200_COMPOSITE
                T.st.reset_fsal(T.u, T.p, T.t, T.nf);
#else
                T.st.init(T.u, T.p, T.t, T.nf);
#endif
            } else
#endif
#if B200_CALLBACKS
            // update_fsal! (integrator_utils.jl:215-239): reeval_fsal => reset_fsal!.  For every FSAL stepper init() IS
            // "fsalfirst = f(u, p, t); nf += 1" and steppers that are not FSAL have nothing to refresh (their init() is
            // empty); the composite algorithm re-evaluates in its running branch
#if B200_COMPOSITE
            if (T.reeval_fsal) T.st.reset_fsal(T.u, T.p, T.t, T.nf); else T.st.accept();
#else
            if (T.reeval_fsal) T.st.init(T.u, T.p, T.t, T.nf); else T.st.accept();
#endif
#else
            T.st.accept();
#endif
            b200_modify_dt_for_tstops(P, T, dist, tol100);
        }
        // (rejected step: step_reject_controller!'s dt /= min(inv(qmin), q11/gamma) was applied by the loopfooter
        //  that rejected it — one shared division, see b200_controller_t)
    }
#if B200_COMPOSITE
    // choose_algorithm!(integrator, integrator.cache) (integrator_utils.jl:121): after iter += 1, before the dt bounds
    if (T.st.choose(T.dt, T.uprev, T.p, T.t, T.nf)) {
        real x;
        x = T.q11; T.q11 = T.q11_o; T.q11_o = x;
        x = T.fpe; T.fpe = T.fpe_o; T.fpe_o = x;
        x = T.rfpe; T.rfpe = T.rfpe_o; T.rfpe_o = x;
    }
#endif
    // fix_dt_at_bounds!
    T.dt = b200_min_c(B200_DTMAX, T.dt);
#if !B200_REVERSE
    T.dt = b200_max_c(dtmin_t, T.dt);
#endif
    // (B200_REVERSE: for tdir < 0 the reference takes min(dt, dtmin) against the positive dtmin, integrator_utils.jl:1250-1254 —
    //  no lower bound on |dt|; only steps that check_error ends anyway can tell the difference)
    b200_modify_dt_for_tstops(P, T, dist, tol100);
    // ---- check_error ---- (flat predicates; the else-if order of the reference decides the code)
    const bool c_nan = b200_isnan(T.dt);
    const bool c_max = ((long long)iter0 + 1 > P.maxiters);
#if B200_ADAPTIVE
#if B200_REVERSE
    // check_error.jl:96 compares `t + dt < tdir * first(opts.tstops)` in both directions; in mirrored time that reads s + ds > stop
    const bool c_min = (b200_abs(T.dt) <= b200_abs(P.dtmin)) & (!T.accept | (T.t + T.dt > tstop));
#else
    const bool c_min = (b200_abs(T.dt) <= b200_abs(P.dtmin)) & (!T.accept | (T.t + T.dt < tstop));   // first(opts.tstops) (check_error.jl:93-99)
#endif
#else
    const bool c_min = false;
#endif
#if B200_ADAPTIVE
    const bool c_uns = (!T.accept) & (b200_abs(T.dt) <= eps_t);
#else
    const bool c_uns = false;           // the dtmin / eps(t) checks are adaptive-only (check_error.jl:91)
#endif
    bool bad = false;
#pragma unroll
    for (int c = 0; c < B200_N; ++c) bad = bad | !b200_isfinite(T.u[c]);
    const bool c_inf = T.accept & bad;
#if B200_COMPOSITE
    // `integrator.do_error_check && check_error!(integrator)` (solve.jl:909); loopfooter! sets the flag again
    const bool ok = !(T.st.do_error_check & (c_nan | c_max | c_min | c_uns | c_inf));
#else
    const bool ok = !(c_nan | c_max | c_min | c_uns | c_inf);
#endif
    if (!ok)
        T.retcode = c_nan ? B200_RC_DTNAN : (c_max ? B200_RC_MAXITERS : (c_min ? B200_RC_DTLESSTHANMIN : B200_RC_UNSTABLE));
    // ---- perform_step! / handle_tstop_step! ----
    const bool skip = T.tstop_flag && b200_abs(T.dt) < eps_t;   // integrator_utils.jl:326-333 (eps(|t|) == eps(t))
    __syncwarp(amask);
    if (ok && !skip) {
#if B200_IS_ROSENBROCK
        T.EEst = T.st.attempt(T.uprev, T.u, T.p, T.t, T.dt, P.reltol, P.abstol, T.nf, T.njacs, T.nw, T.nsolve,
                              (B200_EVERYSTEP || P.nslots > 0) && P.nsaveat > 0);
#else
        T.EEst = T.st.attempt(T.uprev, T.u, T.p, T.t, T.dt, P.reltol, P.abstol, T.nf);
#endif
    }
    __syncwarp(amask);
    if (!ok) return true;
    // ---- loopfooter! ----
#if B200_CALLBACKS
    T.reeval_fsal = false;                              // loopfooter_reset!
#endif
    const real ttmp = T.t + T.dt;
#if !B200_ADAPTIVE
    // not adaptive (integrator_utils.jl:650-659): every step is accepted, dtpropose = dt, no controller
    T.accept = true;
    real q = (real)1;
#else
    // stepsize_controller! + step_accept/reject_controller! (one straight-line block, see b200_controller_t)
    real q = (real)1;
    B200Ctl ctl;
    {
#if B200_COMPOSITE
        B200CtlCfg cfg;
        cfg.beta1 = T.st.beta1(); cfg.beta2 = T.st.beta2(); cfg.qsteady_min = (real)1; cfg.qsteady_max = T.st.qsteady_max_cur();
        T.st.do_error_check = true;                 // loopfooter! (integrator_utils.jl:599)
#else
        const B200CtlCfg cfg = b200_ctl_cfg_static();
#endif
#ifdef B200_ISOUT
        const bool isout = B200_ISOUT(T.u, T.p, B200_USER_T(ttmp)) != (real)0;      // opts.isoutofdomain(u, p, ttmp)
#else
        const bool isout = false;
#endif
        bool bad = false;
        ctl = b200_controller_t<true>(T.EEst, T.q11, T.fpe, T.rfpe, T.dt, T.dtpropose, T.tstop_flag, T.naccept == 0, bad, cfg, isout);
        if (bad) {      // cold, inline
            bool unused = false;
            ctl = b200_controller_t<false>(T.EEst, T.q11, T.fpe, T.rfpe, T.dt, T.dtpropose, T.tstop_flag, T.naccept == 0, unused, cfg, isout);
        }
    }
    T.q11 = ctl.q11;
    T.accept = ctl.accept;
#endif
    if (T.accept) {
        T.naccept += 1;
        T.tprev = T.t;
#if B200_EVERYSTEP
        const real dt_stages = T.dt;                // what perform_step! ran with (the dense pass recomputes the stages from it)
#endif
#if B200_ADAPTIVE
        T.dt = ctl.num;                             // tstop_flag: the un-clipped dt is restored (integrator_utils.jl:629-633)
#endif
        T.t = T.tstop_flag ? tstop : ttmp;          // fixed_t_for_tstop_error! (tstop_target)
        T.tstop_flag = false;
#if !B200_ADAPTIVE
        T.dtpropose = T.dt;
        (void)q; (void)beta1; (void)beta2; (void)qmin; (void)qmax; (void)gamma;
#else
        // step_accept_controller!: errold = max(EEst, qoldinit) — only fastpower(errold, beta2) and its reciprocal are kept
        T.fpe = ctl.fpe;
        T.rfpe = ctl.rfpe;
        const real dtnew = ctl.dtdiv;
        // calc_dt_propose!: eps at the NEW t
        const real eps_n = b200_eps_finite(T.t);
        T.dtpropose = b200_max_c(eps_n > P.dtmin ? eps_n : P.dtmin, b200_min_c(b200_abs(B200_DTMAX), b200_abs(dtnew)));
#endif
        // handle_callbacks! -> savevalues!
#if B200_CALLBACKS
        b200_handle_callbacks(P, idx, T);
        if (T.terminated) return true;                  // terminate! emptied the tstops
#else
#ifndef B200_NO_SAVEAT
#define B200_NO_SAVEAT 0        // 1: the program is only launched without a saveat grid (final states / start-end rows)
#endif
#if !B200_STAGE_ROWS && !(B200_NO_SAVEAT && !B200_EVERYSTEP)      // (staged programs interpolate in b200_stage_rows, after the step, with all lanes of the warp)
        {
            bool dense_ready = false;
            real rdt = (real)0;         // refined 1/dt, shared by the rows of this step
            if (T.next_save <= T.t) rdt = b200_rcp_refine(T.dt);
            while (T.next_save <= T.t) {
                const real curt = T.next_save;
                T.save_idx += 1;
                T.next_save = (T.save_idx < P.nsaveat) ? P.saveat[T.save_idx] : b200_inf();
                if (curt != T.t) {
                    if (!dense_ready) {
                        T.st.dense_prepare(T.uprev, T.u, T.p, T.tprev, T.dt);
                        dense_ready = true;
                    }
                    real th;
                    {   // Θ = (curt - tprev) / dt: flagged fast division (divisor part hoisted), the plain IEEE
                        // operation in the (cold, non-speculable) flagged case
                        bool bad = false;
                        th = b200_div_rcp(curt - T.tprev, T.dt, rdt, bad);
                        if (bad) th = b200_div_cold(curt - T.tprev, T.dt);
                    }
                    real out[B200_N];
                    T.st.interp(th, T.dt, T.uprev, T.u, out);
                    b200_emit(P, idx, T, curt, out);
                } else {
                    if (curt == B200_TF && !P.save_end) continue;   // skip_saveat_at_tspan_end
                    b200_emit(P, idx, T, T.t, T.u);
                }
            }
#if B200_EVERYSTEP
            // save_everystep && (isempty(sol.t) || (t !== sol.t[end] || iszero(dt)) && (save_end || t !== tspan[2]))
            if ((T.nsaved == 0 || ((T.t != T.last_t || T.dt == (real)0) && (P.save_end || T.t != B200_TF))))
                b200_emit(P, idx, T, T.t, T.u, dt_stages);
#endif
        }
#endif  // !B200_STAGE_ROWS && !B200_NO_SAVEAT
#endif  // !B200_CALLBACKS
#if B200_TSTOPS
        // handle_tstop! (integrator_utils.jl:1290-1314): pop every copy of a stop time that was reached;
        // the list ends with tf, whose pop ends the solve (the return below)
        while (T.t == T.tstop && T.tstop_idx + 1 < P.ntstops) {
            T.tstop_idx += 1;
            T.tstop = P.tstops[T.tstop_idx];
        }
#endif
    } else {
        T.nreject += 1;
#if B200_ADAPTIVE
        T.dt = ctl.dtdiv;                           // step_reject_controller!
#endif
    }
    // while tdir*t < first_tstop
    return !(T.t < B200_TF);
}

// cold path (failed trajectories only): kept out of line so it costs the hot loop no registers
__device__ __noinline__ void b200_zero_rows(real* us, long long idx, int from, int nslots) {
    for (int s = from; s < nslots; ++s) {
        real* dst = us + ((size_t)idx * (size_t)nslots + (size_t)s) * B200_NSAVE;
        for (int c = 0; c < B200_NSAVE; ++c) dst[c] = (real)0;
    }
}

// postamble! (+ writing the per-trajectory results)
B200_D void b200_traj_end(const B200Params& P, long long idx, B200Traj& T) {
    if (T.retcode == B200_RC_DEFAULT) T.retcode = B200_RC_SUCCESS;
    // solution_endpoint_match_cur_integrator! (integrator_utils.jl:540-587):
    //   save_end && (saveiter == 0 || sol.t[saveiter] != t &&
    //                (save_end_user === true || t in saveat || t == tspan[2] || isempty(saveat)))
    // P.save_end: 0 false, 1 default true, 2 explicit true.  A grid point equal to t
    // would already be the last saved row, so `t in saveat` adds nothing here.
    if (P.save_end) {
        bool emit;
        if (T.nsaved == 0) emit = true;
        else {
#if B200_EVERYSTEP
            const real last_t = T.last_t;
#else
            const real last_t = (T.save_idx > 0) ? P.saveat[T.save_idx - 1] : B200_T0;
#endif
            emit = (last_t != T.t) && (P.save_end == 2 || T.t == B200_TF || P.nsaveat == 0);
        }
        if (emit) b200_emit(P, idx, T, T.t, T.u);
    }
#if !B200_EVERYSTEP
    // a trajectory that failed leaves its remaining rows zero (the host does not pre-clear `us`)
    if (P.nslots > 0 && T.nsaved < P.nslots) b200_zero_rows(P.us, idx, T.nsaved, P.nslots);
#endif
#pragma unroll
    for (int c = 0; c < B200_N; ++c) P.u_final[idx * P.uf_ts + c * P.uf_cs] = T.u[c];
    if (P.npeer > 0) {
        // the ordered all-gather, fused: peer stores over NVLink straight into every rank's result, in global order
        const long long g = ((idx / P.peer_block) * P.peer_world + P.peer_rank) * P.peer_block + (idx % P.peer_block);
        for (int r = 0; r < P.npeer; ++r) {
            real* dst = P.peer_out[r] + g * B200_N;
#pragma unroll
            for (int c = 0; c < B200_N; ++c) dst[c] = T.u[c];
        }
    }
    P.t_final[idx] = B200_USER_T(T.t);
    P.naccept[idx] = T.naccept;
    P.nreject[idx] = T.nreject;
    P.nf[idx] = T.nf;
    P.retcode[idx] = T.retcode;
    P.nsaved[idx] = T.nsaved;
#if B200_IS_ROSENBROCK
    P.njacs[idx] = T.njacs; P.nw[idx] = T.nw; P.nsolve[idx] = T.nsolve;
#endif
}


#if B200_STAGE_ROWS
// ---------------------------------------------------------------------------
// savevalues! through shared memory (north_star "Output").  In the plain kernel every lane interpolates its own saveat
// rows inside the accepting step: the `while (next_save <= t)` loop runs 1.9 passes per warp-step at 13 of 32 lanes
// (18 lanes have one row, 4 a second, 0.6 a third), i.e. the interpolation's FP64 work is issued 2.4 times.  Here a
// lane only DEPOSITS: once per step a snapshot of what the interpolant needs (uprev, k1..k7, dt, tprev) and per row
// (grid time, destination, snapshot index) into a per-warp queue; whenever 32 rows are queued the whole warp —
// including lanes that are waiting for a new trajectory — drains them, one row per lane, at full occupancy.
// Layout per warp: snap[field][B200_SQ_SNAPS] (a deposit pass writes consecutive entries, a drain reads consecutive
// or identical entries: conflict-free), row ring curt[64] / dest[64] / sidx[64].  Rows keep their destination, so the
// order in which they are written does not matter; results are bit-identical to the in-step loop (same Theta sequence:
// b200_rcp_refine + b200_div_rcp, same interpolant nesting).  Reference: _savevalues! integrator_utils.jl:340-414.
#define B200_SQ_SNAPS 56
#define B200_SQ_ROWS 64
#define B200_SQ_FIELDS (8 * B200_N + 2)
#define B200_SQ_WARP_BYTES ((B200_SQ_FIELDS * B200_SQ_SNAPS * (int)sizeof(real) + B200_SQ_ROWS * ((int)sizeof(real) + 12) + 16 + 15) / 16 * 16)
extern __shared__ __align__(16) unsigned char b200_sq_smem[];
// The ring positions live in the queue's own shared-memory header between calls (four ints per warp), so the step loop
// carries no extra registers; a call loads them, works on warp-uniform copies and stores them back.
struct B200RowQ {
    real* snap; real* curt; unsigned long long* dest; int* sidx; int* hdr;
    int rhead, rcount, shead, scount;     // warp-uniform
};
B200_D void b200_rowq_bind(B200RowQ& Q) {
    unsigned char* base = b200_sq_smem + (size_t)(threadIdx.x >> 5) * B200_SQ_WARP_BYTES;
    Q.snap = reinterpret_cast<real*>(base);
    Q.dest = reinterpret_cast<unsigned long long*>(base + B200_SQ_FIELDS * B200_SQ_SNAPS * sizeof(real));
    Q.curt = reinterpret_cast<real*>(base + B200_SQ_FIELDS * B200_SQ_SNAPS * sizeof(real) + B200_SQ_ROWS * 8);
    Q.sidx = reinterpret_cast<int*>(base + B200_SQ_FIELDS * B200_SQ_SNAPS * sizeof(real) + B200_SQ_ROWS * (8 + sizeof(real)));
    Q.hdr = Q.sidx + B200_SQ_ROWS;
}
B200_D void b200_rowq_load(B200RowQ& Q) {
    b200_rowq_bind(Q);
    const int4 h = *reinterpret_cast<const int4*>(Q.hdr);
    Q.rhead = h.x; Q.rcount = h.y; Q.shead = h.z; Q.scount = h.w;
}
B200_D void b200_rowq_store(const B200RowQ& Q, unsigned lane) {
    __syncwarp();       // every lane has read the header (b200_rowq_load) before lane 0 overwrites it
    if (lane == 0) *reinterpret_cast<int4*>(Q.hdr) = make_int4(Q.rhead, Q.rcount, Q.shead, Q.scount);
    __syncwarp();
}
// the oldest `nrows` (<= 32) queued rows, one per lane; every lane of the warp calls this
#ifndef B200_SQ_NOINLINE
#define B200_SQ_NOINLINE 0
#endif
#if B200_SQ_NOINLINE
__device__ __noinline__ void b200_drain_rows(B200RowQ& Q, unsigned lane, int nrows) {
#else
B200_D void b200_drain_rows(B200RowQ& Q, unsigned lane, int nrows) {
#endif
    __syncwarp();
    if ((int)lane < nrows) {
        const int r = (Q.rhead + (int)lane) & (B200_SQ_ROWS - 1);
        const real curt = Q.curt[r];
        real* dst = reinterpret_cast<real*>(Q.dest[r]);
        const real* S = Q.snap + Q.sidx[r];
        const real dt = S[(8 * B200_N) * B200_SQ_SNAPS], tprev = S[(8 * B200_N + 1) * B200_SQ_SNAPS];
        real th;
        {   // Theta = (curt - tprev) / dt exactly as the in-step loop computes it
            bool bad = false;
            th = b200_div_rcp(curt - tprev, dt, b200_rcp_refine(dt), bad);
            if (bad) th = b200_div_cold(curt - tprev, dt);
        }
        real b[7];
        b200_tsit5_interp_weights(th, b);
#pragma unroll
        for (int i = 0; i < B200_N; ++i) {
            real acc = S[(B200_N + i) * B200_SQ_SNAPS] * b[0];
#pragma unroll
            for (int j = 1; j < 7; ++j) acc = b200_fma(S[(B200_N * (j + 1) + i) * B200_SQ_SNAPS], b[j], acc);
            dst[i] = b200_fma(dt, acc, S[i * B200_SQ_SNAPS]);
        }
    }
    Q.rhead = (Q.rhead + nrows) & (B200_SQ_ROWS - 1);
    Q.rcount -= nrows;
    if (Q.rcount > 0) {
        // snapshots older than every outstanding row's are free.  (Not simply "older than the oldest row's": the second
        // row of a step is queued after the first rows of its neighbours and refers to an older snapshot than they do.)
        int d = B200_SQ_SNAPS;
#pragma unroll
        for (int j = 0; j < B200_SQ_ROWS; j += 32) {
            if (j + (int)lane < Q.rcount) {
                int e = Q.sidx[(Q.rhead + j + (int)lane) & (B200_SQ_ROWS - 1)] - Q.shead;
                if (e < 0) e += B200_SQ_SNAPS;
                d = e < d ? e : d;
            }
        }
        const int rel = __reduce_min_sync(0xffffffffu, d);
        int h = Q.shead + rel; if (h >= B200_SQ_SNAPS) h -= B200_SQ_SNAPS;
        Q.shead = h; Q.scount -= rel;
    } else {
        int e = Q.shead + Q.scount; if (e >= B200_SQ_SNAPS) e -= B200_SQ_SNAPS;
        Q.shead = e; Q.scount = 0;
    }
    __syncwarp();
}
// after b200_traj_iterate, all 32 lanes: the rows of the step just accepted go into the queue
B200_D void b200_stage_rows(const B200Params& P, long long idx, B200Traj& T, bool active, unsigned lane, unsigned lt_mask) {
    bool pending = active && (T.next_save <= T.t);
    if (__ballot_sync(0xffffffffu, pending) == 0u) return;
    B200RowQ Q;
    b200_rowq_load(Q);
    int my_snap = -1;
    for (;;) {
        // room for one row and one snapshot per lane
        while (Q.rcount > B200_SQ_ROWS - 32 || Q.scount > B200_SQ_SNAPS - 32) {
            b200_drain_rows(Q, lane, Q.rcount < 32 ? Q.rcount : 32);
            if (my_snap >= 0) {     // a snapshot whose rows have all been written is gone: the next row deposits it again
                int pos = my_snap - Q.shead; if (pos < 0) pos += B200_SQ_SNAPS;
                if (pos >= Q.scount) my_snap = -1;
            }
        }
        bool store = false;
        real curt = (real)0;
        real* dst = nullptr;
        if (pending) {
            curt = T.next_save;
            T.save_idx += 1;
            T.next_save = (T.save_idx < P.nsaveat) ? P.saveat[T.save_idx] : b200_inf();
            if (curt != T.t) {
                store = (T.nsaved < P.nslots);
                dst = T.row;
                if (store) T.row += B200_N;
                T.nsaved += 1;
            } else if (!(curt == P.tf && !P.save_end)) {     // skip_saveat_at_tspan_end
                b200_emit(P, idx, T, T.t, T.u);
            }
        }
        const bool need_snap = store && my_snap < 0;
        const unsigned nm = __ballot_sync(0xffffffffu, need_snap);
        if (need_snap) {
            int sidx = Q.shead + Q.scount + __popc(nm & lt_mask);
            if (sidx >= B200_SQ_SNAPS) sidx -= B200_SQ_SNAPS;
            if (sidx >= B200_SQ_SNAPS) sidx -= B200_SQ_SNAPS;
            my_snap = sidx;
            real* S = Q.snap + sidx;
#pragma unroll
            for (int i = 0; i < B200_N; ++i) {
                S[i * B200_SQ_SNAPS] = T.uprev[i];
                S[(B200_N + i) * B200_SQ_SNAPS] = T.st.k1[i];
                S[(2 * B200_N + i) * B200_SQ_SNAPS] = T.st.k2[i];
                S[(3 * B200_N + i) * B200_SQ_SNAPS] = T.st.k3[i];
                S[(4 * B200_N + i) * B200_SQ_SNAPS] = T.st.k4[i];
                S[(5 * B200_N + i) * B200_SQ_SNAPS] = T.st.k5[i];
                S[(6 * B200_N + i) * B200_SQ_SNAPS] = T.st.k6[i];
                S[(7 * B200_N + i) * B200_SQ_SNAPS] = T.st.k7[i];
            }
            S[(8 * B200_N) * B200_SQ_SNAPS] = T.dt;
            S[(8 * B200_N + 1) * B200_SQ_SNAPS] = T.tprev;
        }
        Q.scount += __popc(nm);
        const unsigned sm = __ballot_sync(0xffffffffu, store);
        if (store) {
            const int r = (Q.rhead + Q.rcount + __popc(sm & lt_mask)) & (B200_SQ_ROWS - 1);
            Q.curt[r] = curt;
            Q.dest[r] = reinterpret_cast<unsigned long long>(dst);
            Q.sidx[r] = my_snap;
        }
        Q.rcount += __popc(sm);
        pending = pending && (T.next_save <= T.t);
        if (__ballot_sync(0xffffffffu, pending) == 0u) break;
    }
    b200_rowq_store(Q, lane);
}
// the end of the kernel: whatever is still queued
B200_D void b200_flush_rows(unsigned lane) {
    B200RowQ Q;
    b200_rowq_load(Q);
    while (Q.rcount > 0) b200_drain_rows(Q, lane, Q.rcount < 32 ? Q.rcount : 32);
    b200_rowq_store(Q, lane);
}
#endif  // B200_STAGE_ROWS
#ifndef B200_BLOCK
#define B200_BLOCK 128
#endif
#ifndef B200_MINBLOCKS
#define B200_MINBLOCKS 1
#endif
#ifndef B200_CHUNK_MAX
#define B200_CHUNK_MAX 64
#endif
#ifndef B200_REFILL_MIN
#define B200_REFILL_MIN 1
#endif

extern "C" __global__ void __launch_bounds__(B200_BLOCK, B200_MINBLOCKS) b200_integrate(B200Params P) {
    B200Traj T;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
#if B200_STAGE_ROWS
    {   // empty queue
        B200RowQ Q;
        b200_rowq_bind(Q);
        Q.rhead = 0; Q.rcount = 0; Q.shead = 0; Q.scount = 0;
        b200_rowq_store(Q, lane);
    }
#endif
#if B200_WIDE
    // only the first B200_WIDE_NT threads of the CTA have a shared-memory column: the others never take a trajectory
    const bool lane_on = threadIdx.x < B200_WIDE_NT;
    T.st.bind();
#else
    const bool lane_on = true;
#endif

    if (!B200_WIDE && (P.flags & B200_FLAG_STATIC_SCHEDULE)) {
        long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        bool live = idx < P.N;
        if (live) {
            b200_traj_begin(P, idx, T);
            live = (T.t < B200_TF);
            if (!live) b200_traj_end(P, idx, T);
        }
        for (;;) {
            const unsigned am = __ballot_sync(0xffffffffu, live);
            if (am == 0u) break;
            bool fin = false;
            if (live) fin = b200_traj_iterate(P, idx, T, am);
#if B200_STAGE_ROWS
            b200_stage_rows(P, idx, T, live, lane, lt_mask);
#endif
            if (live && fin) { b200_traj_end(P, idx, T); live = false; }
        }
#if B200_STAGE_ROWS
        b200_flush_rows(lane);
#endif
        return;
    }

    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long pool_next = 0, pool_end = 0;     // warp-uniform
    long long idx = -1;
    bool active = false;
    // "exhausted" (the global counter ran past N) is not stored: it is pool_next >= N.  A pool that ends exactly
    // at N was the last chunk the counter handed out, so nothing is left in that case either.  (A separate flag
    // was spilled to local memory and re-loaded at the top of every iteration.)
#define B200_EXHAUSTED (pool_next >= P.N)
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !active && lane_on);
        // B200_REFILL_MIN idle lanes are collected before the (single-lane, divergent) begin path runs: fewer passes
        // through it against lanes that wait a few steps (measured)
        if (need != 0u && !B200_EXHAUSTED && (__popc(need) >= B200_REFILL_MIN || need == 0xffffffffu || (P.N - pool_next) < 64)) {
            const int want = __popc(need);
            if (pool_end - pool_next < want) {
                // refill the warp pool: guided chunk (large while plenty of work remains,
                // shrinking towards the tail so warps finish together)
                long long base = 0, got_end = 0;
                if (lane == 0) {
                    unsigned long long cur = *((volatile unsigned long long*)P.work_counter);
                    long long rem = P.N - (long long)cur;
                    long long chunk = rem / (2 * total_warps);
                    if (chunk > B200_CHUNK_MAX) chunk = B200_CHUNK_MAX;
                    long long short_by = want - (pool_end - pool_next);
                    if (chunk < short_by) chunk = short_by;
                    base = (long long)atomicAdd(P.work_counter, (unsigned long long)chunk);
                    got_end = base + chunk;
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                got_end = __shfl_sync(0xffffffffu, got_end, 0);
                // the old pool (if any) is a contiguous range that ends where this one
                // may not begin; keep both ranges by handing out the old one first
                long long old_left = pool_end - pool_next;
                if (!active && lane_on) {
                    int rank = __popc(need & lt_mask);
                    long long cand = (rank < old_left) ? (pool_next + rank) : (base + (rank - old_left));
                    if (cand < P.N) { idx = cand; active = true; b200_traj_begin(P, idx, T); }
                }
                pool_next = base + (want - old_left);
                pool_end = got_end;
                if (pool_end > P.N) pool_end = P.N > pool_next ? P.N : pool_next;
            } else {
                if (!active && lane_on) {
                    int rank = __popc(need & lt_mask);
                    idx = pool_next + rank; active = true;
                    b200_traj_begin(P, idx, T);
                }
                pool_next += want;
            }
            // trajectories that are already finished at t0 >= tf
            if (active && !(T.t < B200_TF)) { b200_traj_end(P, idx, T); active = false; }
        }
        const unsigned am = __ballot_sync(0xffffffffu, active);
        if (am == 0u) {
            if (B200_EXHAUSTED || need == 0u) break;
            continue;
        }
        bool fin = false;
        if (active) fin = b200_traj_iterate(P, idx, T, am);
#if B200_STAGE_ROWS
        b200_stage_rows(P, idx, T, active, lane, lt_mask);
#endif
        if (active && fin) {
            b200_traj_end(P, idx, T);
            active = false;
        }
    }
#if B200_STAGE_ROWS
    b200_flush_rows(lane);
#endif
}
// (rows restricted by save_idxs cannot restart a step; Rosenbrock32's fsalfirst is f(uprev + dt k2) of the previous
// step, not f of the saved row, so its stages are not recomputable from (row, dt) either)
#if B200_EVERYSTEP && !defined(B200_SAVE_IDXS) && B200_ALG != B200_ALG_ROS32 && !B200_COMPOSITE && !B200_CALLBACKS
// ---------------------------------------------------------------------------
// sol(tq) for every trajectory, post hoc, from the ragged per-step rows — ode_interpolation
// (dense/generic_dense.jl:833-867: interval search :845-849, dt = ts[i+] - ts[i-], Θ :858-859,
// evaluate_interpolant :795-825 -> _ode_addsteps! + ode_interpolant).  The reference keeps every
// step's stage derivatives ks[i] (7-16 vectors per row, integrator_utils.jl:455-473); here they are
// RECOMPUTED from (u[i-], ts[i-], the step's own dt) — the same deterministic arithmetic, so the
// values are bit-identical — which costs one perform_step! per visited interval instead of 8-17x
// the HBM footprint.  Lazy extra stages (Vern7 k11..k16) use dt = ts[i+] - ts[i-] as the reference's
// post-hoc _ode_addsteps! does.  tq must be ascending (the reference sorts the queries first).
// Reverse-time programs (B200_REVERSE): the rows hold the caller's descending times, the kernel works on their mirror images
// (B200_TS / B200_DTS; ode_interpolation searches by tdir * t, generic_dense.jl:838-849) and takes tq descending.
#define B200_TS(i) B200_USER_T(ts[i])
#define B200_DTS(i) B200_USER_T(dts[i])
struct B200DenseParams {
    long long N;
    const real* p; long long p_ts, p_cs;
    const long long* row_offsets; const real* ts; const real* dts; const real* us;
    const real* tq; int M;
    real* out;                // [N][M][n]
    real reltol, abstol;      // only feed the (unused) error estimate of the recomputed step
};

extern "C" __global__ void __launch_bounds__(128) b200_dense_eval(B200DenseParams D) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D.N) return;
    real p[B200_NP > 0 ? B200_NP : 1];
#pragma unroll
    for (int c = 0; c < B200_NP; ++c) p[c] = D.p[idx * D.p_ts + c * D.p_cs];
    const long long a = D.row_offsets[idx];
    const int nrows = (int)(D.row_offsets[idx + 1] - a);
    const real* ts = D.ts + a;
    const real* dts = D.dts + a;
    const real* us = D.us + (size_t)a * B200_N;
    real* out = D.out + (size_t)idx * (size_t)D.M * B200_N;
    B200Stepper st;
    real uprev[B200_N], u[B200_N], scratch[B200_N];
    int hi = 1, cur = -1, nf = 0;
#if B200_IS_ROSENBROCK
    int njacs = 0, nw = 0, nsolve = 0;
#endif
    for (int j = 0; j < D.M; ++j) {
        const real t = B200_USER_T(D.tq[j]);
        real* o = out + (size_t)j * B200_N;
        if (nrows < 2) {      // a single row: i- = i+, dt = 0 => the interpolant collapses to that row
#pragma unroll
            for (int c = 0; c < B200_N; ++c) o[c] = nrows == 1 ? us[c] : (real)0;
            continue;
        }
        // i+ = min(lastindex, max(previous i+, first i with ts[i] >= t)); i- = i+ - 1
        while (hi < nrows - 1 && B200_TS(hi) < t) hi += 1;
        const int ip = hi, im = hi - 1;
        const real dt = B200_TS(ip) - B200_TS(im);
        const real th = (dt == (real)0) ? (real)1 : (t - B200_TS(im)) / dt;
        if (ip != cur) {
#pragma unroll
            for (int c = 0; c < B200_N; ++c) { uprev[c] = us[(size_t)im * B200_N + c]; u[c] = us[(size_t)ip * B200_N + c]; }
            st.init(uprev, p, B200_TS(im), nf);                                   // FSAL k1 = f(u[i-], ts[i-])
#if B200_IS_ROSENBROCK
            st.attempt(uprev, scratch, p, B200_TS(im), B200_DTS(ip), D.reltol, D.abstol, nf, njacs, nw, nsolve, true);
#else
            st.attempt(uprev, scratch, p, B200_TS(im), B200_DTS(ip), D.reltol, D.abstol, nf);
#endif
            st.dense_prepare(uprev, u, p, B200_TS(im), dt);
            cur = ip;
        }
        real val[B200_N];
        st.interp(th, dt, uprev, u, val);
#pragma unroll
        for (int c = 0; c < B200_N; ++c) o[c] = val[c];
    }
}
#endif  // B200_EVERYSTEP
#endif  // !B200_COOP
