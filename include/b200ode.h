/* b200ode.h — C ABI of the B200 ensemble ODE path.
 *
 * This is the drop-in boundary for
 *     solve(EnsembleProblem(prob; prob_func), alg, EnsembleB200(); trajectories, saveat, reltol, abstol, …)
 * i.e. the work the reference performs in SciMLBase's ensemble driver
 * (`__solve(::AbstractEnsembleProblem, alg, ensemblealg)` → `solve_batch` → one
 * `solve(prob_i, alg; kw...)` per trajectory — external to the reference tree, its
 * contract is exercised at /root/reference/lib/DiffEqBase/test/downstream/ensemble.jl:51-112)
 * and, per trajectory, everything below `__solve`
 * (/root/reference/lib/OrdinaryDiffEqCore/src/solve.jl:1-12, 128-884, 904-946).
 *
 * The Julia side (`julia/EnsembleB200.jl`) binds these entry
 * points with `ccall`; the Python host mirror binds them with ctypes.  Plain
 * pointers and sizes only; every function returns 0 on success or a negative
 * B200ODE_E* code, and `b200ode_last_error` gives the message.  Numerical failure
 * of a trajectory is NOT an error: it is reported per trajectory in `retcode[]`
 * (the reference's split between thrown ArgumentErrors and ReturnCodes,
 * /root/reference/lib/DiffEqBase/src/check_error.jl:70-135).
 */
#ifndef B200ODE_H
#define B200ODE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums ------------------------------------------------------------- */
/* alg: the reference algorithm whose perform_step! the kernel reproduces */
#define B200ODE_ALG_TSIT5 1        /* lib/OrdinaryDiffEqTsit5/src/tsit_perform_step.jl:140-186 */
#define B200ODE_ALG_VERN7 2        /* lib/OrdinaryDiffEqVerner/src/verner_rk_perform_step.jl:256-383 */
#define B200ODE_ALG_ROSENBROCK23 3 /* lib/OrdinaryDiffEqRosenbrock/src/rosenbrock_perform_step.jl:249-332 */
#define B200ODE_ALG_RODAS5P 4      /* lib/OrdinaryDiffEqRosenbrock/src/rosenbrock_perform_step.jl:431-559 */
#define B200ODE_ALG_DP5 5          /* lib/OrdinaryDiffEqLowOrderRK/src/low_order_rk_perform_step.jl:667-710 */
#define B200ODE_ALG_BS3 6          /* lib/OrdinaryDiffEqLowOrderRK/src/low_order_rk_perform_step.jl:13-36 */
/* the RodasTableau family: same stepper as Rodas5P (rosenbrock_perform_step.jl:431-559) over the tableaus of
 * lib/OrdinaryDiffEqRosenbrockTableaus/src/rosenbrock_tableaus.jl:26-220 */
#define B200ODE_ALG_RODAS5 7
#define B200ODE_ALG_RODAS4 8
#define B200ODE_ALG_RODAS42 9
#define B200ODE_ALG_RODAS4P 10
#define B200ODE_ALG_RODAS4P2 11

#define B200ODE_F64 0
#define B200ODE_F32 1

/* retcode[] values (mirror SciMLBase.ReturnCode as used by check_error.jl:77-117) */
#define B200ODE_RC_DEFAULT 0
#define B200ODE_RC_SUCCESS 1
#define B200ODE_RC_MAXITERS 2
#define B200ODE_RC_DTLESSTHANMIN 3
#define B200ODE_RC_UNSTABLE 4
#define B200ODE_RC_DTNAN 5

/* status codes */
#define B200ODE_OK 0
#define B200ODE_EINVAL (-1)     /* bad argument (the reference throws ArgumentError / KeywordArgError) */
#define B200ODE_ECOMPILE (-2)   /* NVRTC rejected the RHS/Jacobian source */
#define B200ODE_ECUDA (-3)      /* CUDA runtime error (no device, launch failure, out of memory) */
#define B200ODE_EUNSUPPORTED (-4)

/* layout of per-trajectory arrays handed to the *_device entry points */
#define B200ODE_LAYOUT_AOS 0    /* [trajectory][component]  (what Vector{SVector{n,T}} is) */
#define B200ODE_LAYOUT_SOA 1    /* [component][trajectory]  (coalesced device layout) */

/* opts.flags */
#define B200ODE_FLAG_STATIC_SCHEDULE 1  /* one thread = one trajectory, no lane refill (A/B baseline) */

/* extra_options of b200ode_compile that select a program variant */
#define B200ODE_OPT_EVERYSTEP "-DB200_EVERYSTEP=1"  /* save_everystep = true: ragged per-step rows (b200ode_solve_everystep) */
/* "-DB200_SAVE_IDXS=i0,i1,..." (0-based, no spaces): the save_idxs keyword (solve.jl; _savevalues!
 * integrator_utils.jl:368-375).  Every saved row — `us` of b200ode_solve[_device], the ragged rows, the mean/var
 * statistics — then holds only the listed components, in that order: replace n by the list length in those
 * shapes.  u_final stays the full state.  Not combinable with dense output. */
#define B200ODE_OPT_SAVE_IDXS_PREFIX "-DB200_SAVE_IDXS="

typedef struct b200ode_handle_s* b200ode_handle;     /* one per process per GPU */
typedef struct b200ode_program_s* b200ode_program;   /* one per (alg, dtype, n, np, RHS source) */

/* Replaces: ODEProblem fields harvested from prob_func(prob, ctx) for sim_id = 1..trajectories
 * (u0_i, p_i), plus the shared tspan.  All trajectories share (t0, tf). */
typedef struct {
    int64_t trajectories;
    const void* u0;      /* real[trajectories][n], or real[n] when u0_shared */
    int32_t u0_shared;
    const void* p;       /* real[trajectories][np], or real[np] when p_shared; may be NULL when np == 0 */
    int32_t p_shared;
    double t0, tf;       /* tf > t0 (forward time only) */
} B200Problem;

/* Replaces: the keyword arguments of solve that the path honours
 * (/root/reference/lib/OrdinaryDiffEqCore/src/solve.jl:134-161, defaults :377-399). */
typedef struct {
    double reltol;        /* <= 0: default 1e-3 */
    double abstol;        /* <= 0: default 1e-6 */
    double dt;            /* 0: automatic initial step (initdt.jl:346-459) */
    double dtmin;         /* default 0 */
    double dtmax;         /* <= 0: tf - t0 */
    int64_t maxiters;     /* <= 0: 1000000 */
    const double* saveat; /* explicit ascending (non-decreasing) grid, every entry in (t0, tf]; NULL/0: final state only.
                             (The caller expands `saveat = h` with the reference's range arithmetic,
                             solve.jl:1103-1124.) */
    int32_t nsaveat;
    int32_t save_start;   /* -1 default (true), 0, 1 */
    int32_t save_end;     /* -1 default (true), 0, 1 (explicit true) */
    int32_t flags;
    int32_t reserved;
} B200Opts;

/* Caller-allocated outputs.  Any pointer may be NULL if that output is not wanted,
 * except u_final.  Replaces the fields of each trajectory's ODESolution
 * (sol.u, sol.t, sol.retcode, sol.stats.{naccept,nreject,nf,njacs,nw,nsolve}). */
typedef struct {
    void* u_final;        /* real[trajectories][n]: state at t_final */
    double* t_final;      /* [trajectories] (== tf unless the trajectory failed) */
    void* us;             /* real[trajectories][nslots][n]; nslots = b200ode_nslots(...) */
    double* ts;           /* [nslots] shared time grid of the saved rows */
    int32_t* nsaved;      /* [trajectories] rows actually written (== nslots on success) */
    int32_t* naccept;
    int32_t* nreject;
    int32_t* nf;
    int32_t* njacs;       /* Rosenbrock only; zero otherwise */
    int32_t* nw;
    int32_t* nsolve;
    int32_t* retcode;
    double kernel_ms;     /* out: compute-stream time from the first kernel launch to the last kernel's end (CUDA
                             events).  With a saveat grid the launches are chunked; the D2H copies of finished chunks
                             run on a second stream and are not part of this figure. */
    double total_ms;      /* out: device time including H2D/D2H copies */
} B200Result;

/* Device-resident variant: same meaning, pointers are device pointers on the
 * handle's GPU, layouts selectable, asynchronous on `stream`. */
typedef struct {
    int64_t trajectories;
    const void* u0; int32_t u0_shared; int32_t u0_layout;
    const void* p;  int32_t p_shared;  int32_t p_layout;
    double t0, tf;
} B200DeviceProblem;

typedef struct {
    void* u_final; int32_t u_final_layout; int32_t pad0;
    double* t_final;      /* NOTE: real-typed on device: float[trajectories] for F32 programs */
    void* us;             /* real[trajectories][nslots][n] */
    int32_t* nsaved; int32_t* naccept; int32_t* nreject; int32_t* nf;
    int32_t* njacs; int32_t* nw; int32_t* nsolve; int32_t* retcode;
} B200DeviceResult;

typedef struct {
    int32_t regs_integrate, regs_initdt;
    int32_t local_bytes_integrate, local_bytes_initdt;
    int32_t smem_bytes_integrate;
    int32_t block, blocks_per_sm, grid;    /* launch configuration chosen for b200_integrate */
    int64_t cubin_bytes;
    double compile_ms;
} B200ProgramInfo;

/* ---- lifecycle ---------------------------------------------------------- */
int b200ode_create(b200ode_handle* out, int device_id);
int b200ode_destroy(b200ode_handle h);
/* message of the last failure on this thread (h may be NULL) */
const char* b200ode_last_error(b200ode_handle h);
const char* b200ode_version(void);

/* ---- compilation --------------------------------------------------------
 * rhs_src is C source as emitted by Symbolics `build_function(...; target = CTarget())`:
 *     void NAME(real* du, const real* u, const real* p, const real t) { du[0] = …; }
 * (real = double for F64, float for F32; `#include` lines are ignored).
 * jac_src:   void NAME(real* J, const real* u, const real* p, const real t)  — J column-major n×n
 * tgrad_src: void NAME(real* dT, const real* u, const real* p, const real t)
 * jac/tgrad are required for the Rosenbrock algorithms (the reference's has_jac /
 * has_tgrad branches, lib/OrdinaryDiffEqDifferentiation/src/derivative_utils.jl:188-189,337-338)
 * and ignored otherwise.  User code is compiled without floating-point contraction. */
int b200ode_compile(b200ode_handle h, b200ode_program* out, int alg, int dtype, int n, int np,
                    const char* rhs_src, const char* rhs_name,
                    const char* jac_src, const char* jac_name,
                    const char* tgrad_src, const char* tgrad_name,
                    const char* extra_options);
int b200ode_program_destroy(b200ode_program prog);
int b200ode_program_info(b200ode_program prog, B200ProgramInfo* info);

/* NVRTC only (no GPU needed): compile to an sm_100a cubin and return the compiler
 * log (register counts, spills).  Buffers are malloc'ed; free with b200ode_free. */
int b200ode_compile_only(int alg, int dtype, int n, int np,
                         const char* rhs_src, const char* rhs_name,
                         const char* jac_src, const char* jac_name,
                         const char* tgrad_src, const char* tgrad_name,
                         const char* extra_options,
                         void** cubin, size_t* cubin_bytes, char** log);
void b200ode_free(void* p);

/* ---- solving ------------------------------------------------------------ */
/* rows per trajectory of `us`/`ts` for these options (0 when no saveat grid) */
int b200ode_nslots(const B200Problem* prob, const B200Opts* opts);

/* host buffers in, host buffers out (blocking) */
int b200ode_solve(b200ode_handle h, b200ode_program prog, const B200Problem* prob, const B200Opts* opts,
                  B200Result* result);

/* device buffers in/out, asynchronous on `stream` (a cudaStream_t; NULL = default stream) */
int b200ode_solve_device(b200ode_handle h, b200ode_program prog, const B200DeviceProblem* prob,
                         const B200Opts* opts, B200DeviceResult* result, void* stream);

/* ---- save_everystep = true (the reference's default when saveat is empty, solve.jl:138) -----------------
 * Every accepted step's (t, u) is a row (_savevalues!, integrators/integrator_utils.jl:385-411), plus the
 * save_start row, any saveat rows (interpolated, in time order) and the end point
 * (solution_endpoint_match_cur_integrator!, :540-587).  Row counts differ per trajectory, so the output is
 * ragged: trajectory i owns rows row_offsets[i] .. row_offsets[i+1]-1 of ts[] / us[][n] — its sol.t and sol.u.
 * The program must have been compiled with B200ODE_OPT_EVERYSTEP in extra_options.
 * Two passes over the same deterministic integration: count, exclusive scan, fill (no per-trajectory cap,
 * no wasted HBM).  Dense `sol(t)` needs no stored k's: Tsit5/Vern7/Rosenbrock stages are recomputable from
 * consecutive rows. */
typedef struct {
    int64_t total_rows;
    int64_t* row_offsets;  /* [trajectories + 1]; malloc'ed by the call, release with b200ode_free */
    double* ts;            /* [total_rows];       "  */
    void* us;              /* real[total_rows][n]; "  */
} B200Ragged;

/* host buffers in; per-trajectory scalars into caller-allocated `result` (result.us / result.ts are ignored,
 * result.nsaved[i] = rows of trajectory i); ragged rows into `out` */
int b200ode_solve_everystep(b200ode_handle h, b200ode_program prog, const B200Problem* prob, const B200Opts* opts,
                            B200Result* result, B200Ragged* out);

/* device buffers, asynchronous on `stream`.  row_offsets == NULL: counting pass (result.nsaved only).
 * Otherwise row_offsets is a device int64[trajectories + 1] exclusive scan of those counts, result.us is
 * real[total_rows][n], ts and dts are real[total_rows] (real-typed on device); dts[r] is the step size the
 * stages of the step that ends at row r were computed with (what b200ode_dense_eval_device recomputes from). */
int b200ode_solve_everystep_device(b200ode_handle h, b200ode_program prog, const B200DeviceProblem* prob,
                                   const B200Opts* opts, B200DeviceResult* result, const int64_t* row_offsets,
                                   void* ts, void* dts, void* stream);

/* ---- dense output: sol(tq) for every trajectory, post hoc ----------------------------------------------
 * Replaces ODESolution's interpolation object — ode_interpolation(tvals, id, idxs, deriv, p)
 * (lib/OrdinaryDiffEqCore/src/dense/generic_dense.jl:833-867; interval rule ts[i-] < t <= ts[i+], extrapolating
 * from the first/last interval outside [t0, t_end]) — for idxs = nothing, deriv = Val{0}, continuity = :left.
 * The reference stores every step's stage derivatives (sol.k, integrator_utils.jl:455-473); this path stores
 * one extra scalar per row (dts) and recomputes the stages, bit-identically, inside the evaluation kernel.
 * tq: real[nq] device array, ascending; out: real[trajectories][nq][n] device array. */
int b200ode_dense_eval_device(b200ode_handle h, b200ode_program prog, int64_t trajectories, const void* p, int32_t p_shared,
                              int32_t p_layout, const int64_t* row_offsets, const void* ts, const void* dts,
                              const void* us, const void* tq, int32_t nq, void* out, const B200Opts* opts, void* stream);

/* host buffers: integrate with save_everystep (no saveat), evaluate sol_i(tq[j]) on the device and return only
 * out = real[trajectories][nq][n] plus the per-trajectory scalars. */
int b200ode_solve_dense(b200ode_handle h, b200ode_program prog, const B200Problem* prob, const B200Opts* opts,
                        const double* tq, int32_t nq, void* out, B200Result* result);

/* ---- ensemble reductions (the `reduction` of EnsembleProblem, on device) ---
 * out[c] = sum over trajectories of x(i, c) as double, deterministic order
 * (lib/DiffEqBase/test/downstream/ensemble.jl:100-108 is the reference's `u + sum(batch)`).
 * x is a device array real[count][n] (AOS) or real[n][count] (SOA); `out` is a
 * device array double[n].  The caller divides by the global trajectory count after
 * the cross-GPU all-reduce. */
int b200ode_reduce_sum_device(b200ode_handle h, int dtype, const void* x, int layout, int64_t count, int n,
                              double* out, void* stream);

/* EnsembleAnalysis.timeseries_steps_meanvar on the device (SciMLBase.EnsembleAnalysis, EXT; exercised at
 * /root/reference/lib/DiffEqBase/test/downstream/ensemble_analysis.jl:12-33): for every saved row s and
 * component c, mean[s][c] and the corrected sample variance var[s][c] over the trajectories of
 * us = real[count][nslots][n] (device).  mean/var are device arrays double[nslots][n]; var may be NULL.
 * Two passes in fixed order (deterministic).  This is SURVEY §8(f) row 1: ensemble statistics without
 * moving the trajectories across PCIe. */
int b200ode_timeseries_meanvar_device(b200ode_handle h, int dtype, const void* us, int64_t count, int nslots, int n,
                                      double* mean, double* var, void* stream);

/* b200ode_solve + the reduction above, host buffers: `result->us` is ignored (the rows stay in HBM);
 * mean/var are host arrays double[nslots][n].  Only the statistics and the per-trajectory scalars
 * cross PCIe. */
int b200ode_solve_meanvar(b200ode_handle h, b200ode_program prog, const B200Problem* prob, const B200Opts* opts,
                          B200Result* result, double* mean, double* var);

/* ---- host memory helpers (pinned buffers make the D2H of `us` run at PCIe speed) */
int b200ode_host_register(void* ptr, size_t bytes);
int b200ode_host_unregister(void* ptr);

/* ---- measurement ---------------------------------------------------------- */
/* Achieved FMA throughput of a register-resident dependent-chain microbenchmark
 * (the roofline denominator of this path).  dtype F64/F32; returns TFLOP/s. */
int b200ode_measure_fma_peak(b200ode_handle h, int dtype, double* tflops, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* B200ODE_H */
