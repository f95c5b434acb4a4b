"""Host-side mirror of the reference's ensemble interface for the B200 path.

    prob  = ODEProblem(ODEFunction(f; jac, tgrad), u0, tspan, p)
    eprob = EnsembleProblem(prob; prob_func, output_func, reduction, u_init, safetycopy)
    sim   = solve(eprob, Tsit5(), EnsembleB200(); trajectories, batch_size, saveat, reltol, abstol, ...)

mirrors `solve(::EnsembleProblem, alg, ensemblealg; kw...)` of SciMLBase (EXT; the contract the
reference exercises at /root/reference/lib/DiffEqBase/test/downstream/ensemble.jl:51-112 and
documents at /root/reference/NEWS.md:209-224): batches of `batch_size`, `prob_func(prob, ctx)`
per trajectory, `output_func(sol, ctx) -> (out, rerun)`, `reduction(u, batch, I) -> (u, converged)`.

What changes behind `EnsembleB200()`: nothing is solved on the host.  `prob_func` is evaluated
only to harvest (u0_i, p_i) into flat tables (or is a `TableProbFunc` that already is a table),
and one C-ABI call integrates the whole batch on the GPU.  There is no CPU fallback: any other
ensemble algorithm raises.
"""
import time

import numpy as np

from . import _lib, codegen, lowlevel, ranges


# ---- algorithms (names and traits of the reference) -------------------------------------------
class _Alg:
    alg_id = 0
    order = 0
    stiff = False

    def __repr__(self):
        return type(self).__name__ + "()"


class Tsit5(_Alg):          # lib/OrdinaryDiffEqTsit5
    alg_id, order = _lib.ALG_TSIT5, 5


class Vern7(_Alg):          # lib/OrdinaryDiffEqVerner (lazy = true)
    alg_id, order = _lib.ALG_VERN7, 7


class Rosenbrock23(_Alg):   # lib/OrdinaryDiffEqRosenbrock
    alg_id, order, stiff = _lib.ALG_ROSENBROCK23, 2, True


class Rodas5Pe(_Alg):
    alg_id, order, stiff = _lib.ALG_RODAS5PE, 5, True


class Rosenbrock32(_Alg):
    alg_id, order, stiff = _lib.ALG_ROSENBROCK32, 3, True


class Rodas5P(_Alg):
    alg_id, order, stiff = _lib.ALG_RODAS5P, 5, True


class Vern6(_Alg):
    alg_id, order = _lib.ALG_VERN6, 6


class Vern8(_Alg):
    alg_id, order = _lib.ALG_VERN8, 8


class Vern9(_Alg):
    alg_id, order = _lib.ALG_VERN9, 9


class Rodas5(_Alg):         # RodasTableau family (lib/OrdinaryDiffEqRosenbrockTableaus)
    alg_id, order, stiff = _lib.ALG_RODAS5, 5, True


class Rodas4(_Alg):
    alg_id, order, stiff = _lib.ALG_RODAS4, 4, True


class Rodas42(_Alg):
    alg_id, order, stiff = _lib.ALG_RODAS42, 4, True


class Rodas4P(_Alg):
    alg_id, order, stiff = _lib.ALG_RODAS4P, 4, True


class Rodas23W(_Alg):
    alg_id, order, stiff = _lib.ALG_RODAS23W, 3, True


class Rodas3P(_Alg):
    alg_id, order, stiff = _lib.ALG_RODAS3P, 3, True


class Rodas4P2(_Alg):
    alg_id, order, stiff = _lib.ALG_RODAS4P2, 4, True


class DP5(_Alg):            # lib/OrdinaryDiffEqLowOrderRK
    alg_id, order = _lib.ALG_DP5, 5


class BS3(_Alg):
    alg_id, order = _lib.ALG_BS3, 3


def AutoTsit5(stiff_alg):
    """AutoTsit5(Rosenbrock23()) = AutoAlgSwitch(Tsit5(), stiff_alg) with the default AutoSwitch parameters
    (lib/OrdinaryDiffEqTsit5/src/algorithms.jl:27-33, lib/OrdinaryDiffEqCore/src/composite_algs.jl:4-15)."""
    if not isinstance(stiff_alg, Rosenbrock23):
        raise NotImplementedError("AutoTsit5 is available with Rosenbrock23() as the stiff algorithm")
    return _AutoTsit5Rosenbrock23()


class _AutoTsit5Rosenbrock23(_Alg):
    alg_id, order, stiff = _lib.ALG_AUTOTSIT5_ROSENBROCK23, 5, True

    def __repr__(self):
        return "AutoTsit5(Rosenbrock23())"


class EnsembleAlgorithm:
    pass


class EnsembleB200(EnsembleAlgorithm):
    """The drop-in ensemble algorithm: all trajectories of a batch in one GPU launch."""

    def __init__(self, device=None):
        self.device = device


class _NoCPU(EnsembleAlgorithm):
    def __init__(self, *a, **k):
        raise NotImplementedError(
            type(self).__name__ + " is the reference's CPU path; this package only provides EnsembleB200() "
            "and has no CPU fallback")


class EnsembleSerial(_NoCPU):
    pass


class EnsembleThreads(_NoCPU):
    pass


class EnsembleDistributed(_NoCPU):
    pass


# ---- problem types ----------------------------------------------------------------------------
class CSource:
    """C source of a function `void name(real* out, const real* u, const real* p, const real t)`."""

    def __init__(self, source, name):
        self.source, self.name = source, name


# ---- callbacks (lib/DiffEqBase/src/callbacks.jl; constructors are SciMLBase's, EXT) --------------------------------------
class ContinuousCallback:
    """ContinuousCallback(condition, affect!, affect_neg! = affect!; rootfind = LeftRootFind, save_positions = (true, true),
    interp_points = 10, abstol = 10eps(), repeat_nudge = 1//100).  condition / affect! are CSource objects:
        condition:  real NAME(const real* u, const real* p, const real t)
        affect!:    void NAME(real* u, real* p, const real t, int* terminate)      (*terminate = 1 is terminate!(integrator))
    Pass affect=None or affect_neg=None for `nothing` (events of that direction are ignored)."""
    _same = object()

    def __init__(self, condition, affect, affect_neg=_same, rootfind="left", save_positions=(True, True),
                 interp_points=10, abstol=None, repeat_nudge=None):
        self.condition, self.affect = condition, affect
        self.affect_neg = affect if affect_neg is ContinuousCallback._same else affect_neg
        self.rootfind, self.save_positions, self.interp_points = rootfind, tuple(save_positions), interp_points
        self.abstol, self.repeat_nudge = abstol, repeat_nudge

    def spec(self):
        d = dict(kind="continuous", condition=(self.condition.source, self.condition.name),
                 affect=None if self.affect is None else (self.affect.source, self.affect.name),
                 affect_neg=None if self.affect_neg is None else (self.affect_neg.source, self.affect_neg.name),
                 rootfind=self.rootfind, save_positions=self.save_positions, interp_points=self.interp_points)
        if self.abstol is not None:
            d["abstol"] = self.abstol
        if self.repeat_nudge is not None:
            d["repeat_nudge"] = self.repeat_nudge
        return d


class DiscreteCallback:
    """DiscreteCallback(condition, affect!; save_positions = (true, true)); the condition holds when its C function
    returns a non-zero value."""

    def __init__(self, condition, affect, save_positions=(True, True)):
        self.condition, self.affect, self.save_positions = condition, affect, tuple(save_positions)

    def spec(self):
        return dict(kind="discrete", condition=(self.condition.source, self.condition.name),
                    affect=None if self.affect is None else (self.affect.source, self.affect.name),
                    save_positions=self.save_positions)


class CallbackSet:
    """CallbackSet(cb...): continuous callbacks are handled before the discrete ones, each group in the given order."""

    def __init__(self, *cbs):
        flat = []
        for c in cbs:
            flat.extend(c.callbacks if isinstance(c, CallbackSet) else [c])
        self.callbacks = ([c for c in flat if isinstance(c, ContinuousCallback)] +
                          [c for c in flat if isinstance(c, DiscreteCallback)])
        if len(self.callbacks) != len(flat):
            raise TypeError("CallbackSet takes ContinuousCallback / DiscreteCallback objects")


class ODEFunction:
    """ODEFunction(f; jac, tgrad).  `f` is either a Python callable f(u, p, t) -> [exprs] traced with
    sympy (stand-in for Symbolics tracing) or a CSource; jac/tgrad likewise, or None to have them
    derived symbolically from a traceable f."""

    def __init__(self, f, jac=None, tgrad=None, iip=False):
        self.f, self.jac, self.tgrad, self.iip = f, jac, tgrad, iip

    def sources(self, n, np_, f32, need_jac):
        def one(obj, builder, fname):
            if obj is None:
                return None
            if isinstance(obj, CSource):
                return (obj.source, obj.name)
            return builder(obj, n, np_, fname=fname, f32=f32, iip=self.iip)
        rhs = one(self.f, codegen.build_function_c, "diffeqf")
        jac = tg = None
        if need_jac:
            if self.jac is not None:
                # a Python jac(u, p, t) returns the n x n matrix (rows of rows / sympy Matrix): flattened column-major
                jac = one(self.jac, codegen.build_matrix_c, "diffeqjac")
            elif not isinstance(self.f, CSource):
                jac = codegen.build_jacobian_c(self.f, n, np_, f32=f32, iip=self.iip)
            else:
                raise ValueError("Rosenbrock methods need a Jacobian: ODEFunction(f; jac=...) with a CSource f")
            if self.tgrad is not None:
                tg = one(self.tgrad, codegen.build_function_c, "diffeqtgrad")
            elif not isinstance(self.f, CSource):
                tg = codegen.build_tgrad_c(self.f, n, np_, f32=f32, iip=self.iip)
        return rhs, jac, tg


class ODEProblem:
    def __init__(self, f, u0, tspan, p=None, **kwargs):
        self.f = f if isinstance(f, ODEFunction) else ODEFunction(f)
        self.u0 = np.asarray(u0)
        self.tspan = (tspan[0], tspan[1])
        self.p = None if p is None else np.asarray(p)
        self.kwargs = dict(kwargs)      # prob.kwargs are merged under solve's (lib/DiffEqBase/src/solve.jl:49-74)


def remake(prob, u0=None, p=None, tspan=None):
    return ODEProblem(prob.f, prob.u0 if u0 is None else u0, prob.tspan if tspan is None else tspan,
                      prob.p if p is None else p, **prob.kwargs)


class EnsembleContext:
    def __init__(self, sim_id, repeat=1, rng=None):
        self.sim_id, self.repeat, self.rng = sim_id, repeat, rng


class TableProbFunc:
    """prob_func given as tables: trajectory i (1-based sim_id) gets u0[i-1] and/or p[i-1].
    Equivalent to `(prob, ctx) -> remake(prob; u0 = U0[ctx.sim_id], p = P[ctx.sim_id])` without
    a million host-side calls."""

    def __init__(self, u0=None, p=None, tspan=None):
        self.u0 = None if u0 is None else np.asarray(u0)
        self.p = None if p is None else np.asarray(p)
        self.tspan = None if tspan is None else np.asarray(tspan, dtype=np.float64)     # (N, 2): remake(prob; tspan = ...)

    def __call__(self, prob, ctx):
        i = ctx.sim_id - 1
        return remake(prob, u0=None if self.u0 is None else self.u0[i], p=None if self.p is None else self.p[i],
                      tspan=None if self.tspan is None else tuple(self.tspan[i]))


class EnsembleProblem:
    def __init__(self, prob, prob_func=None, output_func=None, reduction=None, u_init=None, safetycopy=None):
        self.prob = prob
        self.prob_func = prob_func
        self.output_func = output_func
        self.reduction = reduction
        self.u_init = u_init
        self.safetycopy = safetycopy    # irrelevant here: the problem is never mutated


# ---- solutions --------------------------------------------------------------------------------
class DEStats:
    def __init__(self, nf, naccept, nreject, njacs=0, nw=0, nsolve=0):
        self.nf, self.naccept, self.nreject = int(nf), int(naccept), int(nreject)
        self.njacs, self.nw, self.nsolve = int(njacs), int(nw), int(nsolve)


class ODESolution:
    def __init__(self, t, u, retcode, stats, prob=None, alg=None, interp=None):
        self.t, self.u, self.retcode, self.stats, self.prob, self.alg = t, u, retcode, stats, prob, alg
        self._interp = interp
        self.dense = interp is not None

    def __call__(self, t):
        """sol(t): the stepper's own dense output (ode_interpolation, dense/generic_dense.jl:833-867), evaluated
        on the device from this trajectory's steps.  Available when the solve saved every step and no saveat
        (the reference's default `dense`, solve.jl:144-145)."""
        if self._interp is None:
            raise NotImplementedError("this solution has no dense output (solve with save_everystep and no saveat)")
        scalar = np.isscalar(t)
        tq = np.atleast_1d(np.asarray(t, dtype=np.float64))
        # ode_interpolation evaluates the queries in the order the integration met them (sorted by tdir * t, generic_dense.jl:838)
        tdir = -1.0 if (len(self.t) > 1 and self.t[-1] < self.t[0]) else 1.0
        order = np.argsort(tdir * tq, kind="stable")
        vals = self._interp(tq[order])
        out = np.empty_like(vals)
        out[order] = vals
        return out[0] if scalar else out

    def __getitem__(self, i):
        return self.u[i]

    def __len__(self):
        return len(self.t)


class _LazySolutions:
    """Vector{ODESolution} over the flat result arrays (materialised per index on demand)."""

    def __init__(self, res, t0, alg, has_grid, u0=None, save_start=True, save_end=True, dense=None, save_idxs=None):
        self.res, self.t0, self.alg, self.has_grid, self.save_idxs = res, t0, alg, has_grid, save_idxs
        self.u0, self.save_start, self.save_end = u0, save_start, save_end
        self.dense = dense      # callable (i, tq ascending) -> [len(tq), n], or None

    def __len__(self):
        return self.res["u_final"].shape[0]

    def __getitem__(self, i):
        r = self.res
        if i < 0:
            i += len(self)
        if "row_offsets" in r:
            # save_everystep = true: this trajectory's slice of the ragged rows is its sol.t / sol.u
            a, b = int(r["row_offsets"][i]), int(r["row_offsets"][i + 1])
            t, u = r["ts"][a:b], r["us"][a:b]
        elif self.has_grid:
            k = int(r["nsaved"][i])
            t = np.array(r["ts"][:k])
            u = r["us"][i, :k]
            if k > 0 and r["retcode"][i] != _lib.RC_SUCCESS:
                t[-1] = min(t[-1], r["t_final"][i])
        else:
            # save_everystep = false and no saveat: sol.t = [t0, t_end] (save_start / save_end)
            ts, us = [], []
            sel = (lambda v: v) if self.save_idxs is None else (lambda v: np.asarray(v)[self.save_idxs])
            if self.save_start:
                ts.append(self.t0 if np.ndim(self.t0) == 0 else self.t0[i])
                us.append(sel(self.u0 if self.u0.ndim == 1 else self.u0[i]))
            if self.save_end:
                ts.append(r["t_final"][i])
                us.append(sel(r["u_final"][i]))
            t = np.array(ts)
            u = np.stack(us) if us else np.zeros((0, r["u_final"].shape[1]), dtype=r["u_final"].dtype)
        st = DEStats(r["nf"][i], r["naccept"][i], r["nreject"][i], r["njacs"][i], r["nw"][i], r["nsolve"][i])
        interp = None if self.dense is None else (lambda tq, i=i: self.dense(i, tq))
        return ODESolution(t, u, _lib.RETCODE_NAMES[int(r["retcode"][i])], st, alg=self.alg, interp=interp)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class EnsembleSolution:
    def __init__(self, u, elapsed, converged, arrays=None, dense_all=None):
        self.u, self.elapsed_time, self.converged, self.arrays = u, elapsed, converged, arrays
        self._dense_all = dense_all

    def at(self, tq):
        """[sol(tq) for sol in ensemble] as one array [trajectories, len(tq), n] in a single device pass
        (b200ode_solve_dense); tq in the order the integration meets the times (ascending; descending for a reversed tspan)."""
        if self._dense_all is None:
            raise NotImplementedError("dense output needs save_everystep, no saveat, and the default reduction")
        return self._dense_all(np.asarray(tq, dtype=np.float64))

    def __getitem__(self, i):
        return self.u[i]

    def __len__(self):
        return len(self.u)


# ---- solve ------------------------------------------------------------------------------------
_ALLOWED_KW = {"trajectories", "batch_size", "saveat", "save_start", "save_end", "save_everystep", "save_idxs", "tstops", "d_discontinuities", "reltol",
               "abstol", "dt", "dtmin", "dtmax", "maxiters", "adaptive", "dense", "dtype", "flags", "save_on", "callback"}
# accepted and ignored: they do not change the numbers (logging / progress / error-statistics switches of solve.jl:166-181)
_IGNORED_KW = {"verbose", "progress", "progress_steps", "progress_name", "progress_message", "progress_id",
               "timeseries_errors", "dense_errors", "alias", "userdata"}
_program_cache = {}
_handles = {}


def _handle(device):
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _handles:
        _handles[device] = _lib.Handle(device)
    return _handles[device]


def get_program(handle, alg, fn, n, np_, f32, everystep=False, save_idxs=None, tstops=False, adaptive=True, callbacks=None,
                vector_tol=False, smem_stages=False, tspans=False, reverse=False):
    rhs, jac, tg = fn.sources(n, np_, f32, alg.stiff)
    extra = []
    if reverse:
        extra.append(_lib.OPT_REVERSE_TIME)
    if tspans:
        extra.append(_lib.OPT_TSPANS)
    if smem_stages:
        extra.append(_lib.OPT_SMEM_STAGES)
    if everystep:
        extra.append(_lib.OPT_EVERYSTEP)
    if tstops:
        extra.append(_lib.OPT_TSTOPS)
    if not adaptive:
        extra.append(_lib.OPT_FIXED_DT)
    if vector_tol:
        extra.append(_lib.OPT_VECTOR_TOL)
    if save_idxs is not None:
        extra.append(_lib.opt_save_idxs(save_idxs))
    extra = " ".join(extra) or None
    key = (handle.device, alg.alg_id, f32, n, np_, rhs, jac, tg, extra, repr(callbacks))
    if key not in _program_cache:
        _program_cache[key] = handle.compile(alg.alg_id, _lib.F32 if f32 else _lib.F64, n, np_, rhs[0], rhs[1],
                                             jac[0] if jac else None, jac[1] if jac else None,
                                             tg[0] if tg else None, tg[1] if tg else None,
                                             extra_options=extra, callbacks=callbacks)
    return _program_cache[key]


def _harvest(eprob, I, repeat=1):
    """Evaluate prob_func for sim_ids in I -> (u0 table or shared vector, p table or shared vector, per-trajectory
    (t0, tf) table or None when every trajectory keeps the problem's tspan)."""
    prob, pf = eprob.prob, eprob.prob_func
    if pf is None:
        return prob.u0, prob.p, None
    if isinstance(pf, TableProbFunc):
        idx = np.asarray(I) - 1
        u0 = prob.u0 if pf.u0 is None else pf.u0[idx]
        p = prob.p if pf.p is None else pf.p[idx]
        return u0, p, (None if pf.tspan is None else np.ascontiguousarray(pf.tspan[idx]))
    u0s, ps, spans = [], [], []
    for i in I:
        q = pf(prob, EnsembleContext(int(i), repeat))
        spans.append((float(q.tspan[0]), float(q.tspan[1])))
        u0s.append(np.asarray(q.u0))
        ps.append(None if q.p is None else np.asarray(q.p))
    u0 = np.stack(u0s)
    p = None if ps[0] is None else np.stack(ps)
    base = (float(prob.tspan[0]), float(prob.tspan[1]))
    return u0, p, (None if all(sp == base for sp in spans) else np.array(spans, dtype=np.float64))


def solve(eprob, alg, ensemblealg=None, **kw):
    """solve(EnsembleProblem, alg, EnsembleB200(); trajectories, ...) -> EnsembleSolution."""
    if not isinstance(eprob, EnsembleProblem):
        raise TypeError("this package accelerates the EnsembleProblem path only")
    if not isinstance(ensemblealg, EnsembleB200):
        raise NotImplementedError("only EnsembleB200() is provided (no CPU fallback)")
    if not isinstance(alg, _Alg):
        raise TypeError("alg must be one of Tsit5(), Vern7(), DP5(), BS3(), Rosenbrock23(), Rodas4/42/4P/4P2/5/5P()")
    prob = eprob.prob
    kw = dict(prob.kwargs, **kw)                      # merge_problem_kwargs: solve's kwargs win
    kw = {k: v for k, v in kw.items() if k not in _IGNORED_KW}
    bad = set(kw) - _ALLOWED_KW
    if bad:
        # the reference throws for unrecognised keywords (lib/DiffEqBase/src/solve.jl:79-93)
        raise TypeError("unsupported keyword arguments for the EnsembleB200 path: %s" % sorted(bad))
    if "trajectories" not in kw:
        raise TypeError("trajectories is required")
    has_saveat = kw.get("saveat", None) is not None and not (hasattr(kw["saveat"], "__len__") and len(kw["saveat"]) == 0)
    # the reference's default is save_everystep = isempty(saveat) (solve.jl:138): ragged per-step output
    everystep = bool(kw.get("save_everystep", not has_saveat))
    adaptive = bool(kw.get("adaptive", True))
    if not adaptive and kw.get("dt") is None and (kw.get("tstops") is None or len(kw["tstops"]) == 0):
        # solve.jl:277-280
        raise ValueError("Fixed timestep methods require a choice of dt or choosing the tstops")
    dense_kw = kw.get("dense", None)     # default: save_everystep && isempty(saveat) (solve.jl:144-145)
    N = int(kw["trajectories"])
    batch_size = int(kw.get("batch_size", N)) if N > 0 else 1
    f32 = kw.get("dtype", None)
    if f32 is None:
        f32 = (prob.u0.dtype == np.float32)
    else:
        f32 = (np.dtype(f32) == np.float32)
    n = int(prob.u0.shape[-1])
    np_ = 0 if prob.p is None else int(np.asarray(prob.p).shape[-1])
    grid = ranges.saveat_grid(kw.get("saveat", None), prob.tspan)
    # save_start / save_end defaults depend on the keywords as given (solve.jl:141-143,596-599) ...
    save_start, save_end = ranges.resolve_save_flags(kw.get("saveat", None), prob.tspan, everystep,
                                                     kw.get("save_start"), kw.get("save_end"))
    if not kw.get("save_on", True):
        # ... while save_on = false only silences _savevalues! (integrator_utils.jl:342): no saveat / per-step rows,
        # the start row and the end point are still stored
        grid, everystep = [], False
    # save_idxs: component indices of the saved rows, 0-based here (the Julia binding converts from 1-based)
    save_idxs = kw.get("save_idxs", None)
    if save_idxs is not None:
        save_idxs = [int(i) for i in (save_idxs if hasattr(save_idxs, "__len__") else [save_idxs])]
        if not save_idxs or min(save_idxs) < 0 or max(save_idxs) >= n:
            raise ValueError("save_idxs out of range for a state of length %d" % n)
    tstops = kw.get("tstops", None)
    tstops = None if tstops is None or len(tstops) == 0 else [float(x) for x in tstops]
    # d_discontinuities (solve.jl:136): stops at which t is moved one ulp on and the first stage is evaluated again
    discs = kw.get("d_discontinuities", None)
    discs = None if discs is None or len(discs) == 0 else [float(x) for x in discs]
    if discs is not None and kw.get("callback") is not None:
        raise NotImplementedError("d_discontinuities are not combined with callbacks")
    # (the dense pass re-integrates with the default end-point handling: a solve with save_end = false keeps sol.t without
    #  the row at tf, which the dense rows would include — declined rather than answered from different rows)
    dense_ok = (everystep and not grid and save_start and save_idxs is None and tstops is None and discs is None and save_end is not False
                and alg.alg_id not in (_lib.ALG_ROSENBROCK32, _lib.ALG_AUTOTSIT5_ROSENBROCK23) and dense_kw is not False)      # dense = save_everystep && isempty(saveat) (solve.jl:144)
    if dense_kw and not dense_ok:
        raise NotImplementedError("dense=true is served for save_everystep solves without saveat / save_idxs / tstops "
                                  "(the stages are recomputed from the saved steps); not for Rosenbrock32")
    # callback = ContinuousCallback / DiscreteCallback / CallbackSet: events on the device (Tsit5).  Rows forced by
    # save_positions make the output ragged even when save_everystep = false.
    cb_specs, cb_flags, ragged = None, 0, everystep
    if kw.get("callback") is not None:
        cbset = kw["callback"] if isinstance(kw["callback"], CallbackSet) else CallbackSet(kw["callback"])
        if cbset.callbacks:
            if not isinstance(alg, Tsit5) and any(isinstance(c, ContinuousCallback) for c in cbset.callbacks):
                raise NotImplementedError("continuous callbacks are available with Tsit5() (discrete callbacks: every algorithm)")
            cb_specs = [c.spec() for c in cbset.callbacks]
            if any(any(c.save_positions) for c in cbset.callbacks) and not everystep:
                ragged, cb_flags = True, _lib.FLAG_NO_STEP_ROWS
            dense_ok = False
            if dense_kw:
                raise NotImplementedError("dense=true is not available with callbacks")
    handle = _handle(ensemblealg.device)
    # abstol / reltol given as vectors: one tolerance per component (solve.jl:377-399)
    vector_tol = np.ndim(kw.get("reltol")) > 0 or np.ndim(kw.get("abstol")) > 0
    for name in ("reltol", "abstol"):
        if np.ndim(kw.get(name)) > 0 and len(kw[name]) != n:
            raise ValueError("%s must be a number or a vector with one entry per component" % name)
    # Vern7 on a wide state with nothing but start / end rows asked for: the stage derivatives do not fit a thread's
    # registers, so the kernel that keeps them in shared memory runs (B200ODE_OPT_SMEM_STAGES; bit-identical results)
    t0_, tf_ = float(prob.tspan[0]), float(prob.tspan[1])
    # tspan[2] < tspan[1]: tdir = -1 (solve.jl:273) — a program compiled with B200ODE_OPT_REVERSE_TIME (mirrored-time kernels)
    reverse = tf_ < t0_
    # (n >= 24: measured crossover on a cheap-RHS chain system, scripts/time_wide_threshold.py — below it the plain kernel's
    #  local-memory stage vectors, served from L1 at full occupancy, are faster than 112-256 threads with shared-memory stages)
    smem_stages = (alg.alg_id == _lib.ALG_VERN7 and n >= 24 and not ragged and cb_specs is None and not reverse
                   and all(not (t0_ < float(g) < tf_) for g in (grid or [])))
    program = get_program(handle, alg, prob.f, n, np_, f32, ragged, save_idxs, tstops is not None or discs is not None, adaptive, cb_specs,
                          vector_tol, smem_stages, False, reverse)

    def run(u0, p, ntraj, flags=0, spans=None):
        common = dict(trajectories=ntraj, reltol=kw.get("reltol"), abstol=kw.get("abstol"), dt=kw.get("dt"),
                      dtmin=kw.get("dtmin"), dtmax=kw.get("dtmax"), maxiters=kw.get("maxiters"),
                      saveat=grid if grid else None, save_start=save_start, save_end=save_end,
                      flags=flags | cb_flags, tstops=tstops, d_discontinuities=discs)
        if spans is None:
            if ragged:
                return lowlevel.solve_host_everystep(program, u0, p, prob.tspan, **common)
            return lowlevel.solve_host(program, u0, p, prob.tspan, **common)
        # prob_func changed tspan: every trajectory integrates over its own span (B200Problem.tspans).  The rows of
        # different trajectories share no grid, so only final states (+ start / end rows) or the ragged output exist;
        # a saveat list is taken as absolute times (each trajectory keeps the entries inside its span), a saveat step is not
        # expressible as one list
        if tstops is not None or discs is not None or cb_specs is not None:
            raise NotImplementedError("per-trajectory tspan is not combined with tstops, d_discontinuities or callbacks")
        if reverse or np.any(spans[:, 1] <= spans[:, 0]):
            raise NotImplementedError("per-trajectory tspan: forward spans only")
        sv = kw.get("saveat", None)
        if sv is not None and not hasattr(sv, "__len__"):
            raise NotImplementedError("per-trajectory tspan with saveat = step: pass the times as a list")
        if grid and not ragged:
            raise NotImplementedError("per-trajectory tspan with saveat needs the ragged output (save_everystep = true)")
        prog_s = get_program(handle, alg, prob.f, n, np_, f32, ragged, save_idxs, False, adaptive, None, vector_tol, False, True)
        if ragged:
            common["saveat"] = None if sv is None or len(sv) == 0 else sorted(float(x) for x in sv)
            return lowlevel.solve_host_everystep(prog_s, u0, p, spans, **common)
        return lowlevel.solve_host(prog_s, u0, p, spans, **common)

    tol_kw = dict(reltol=kw.get("reltol"), abstol=kw.get("abstol"), dt=kw.get("dt"), dtmin=kw.get("dtmin"),
                  dtmax=kw.get("dtmax"), maxiters=kw.get("maxiters"))

    def dense_of(u0, p):
        u0 = np.asarray(u0); p_ = None if p is None else np.asarray(p)

        def one(i, tq):
            ui = u0 if u0.ndim == 1 else u0[i:i + 1]
            pi = p_ if (p_ is None or p_.ndim == 1) else p_[i:i + 1]
            return lowlevel.solve_host_dense(program, ui, pi, prob.tspan, tq, trajectories=1, **tol_kw)["dense"][0]
        return one
    batches_in = []

    t_start = time.perf_counter()
    reduction, output_func = eprob.reduction, eprob.output_func
    u_acc = [] if eprob.u_init is None else eprob.u_init
    converged = False
    all_arrays = []
    for b0 in range(0, N, max(batch_size, 1)):
        I = np.arange(b0 + 1, min(b0 + batch_size, N) + 1)
        u0, p, spans = _harvest(eprob, I)
        res = run(u0, p, len(I), kw.get("flags", 0), spans)
        all_arrays.append(res)
        mk = lambda r, u0_, p_=None, sp_=None: _LazySolutions(r, prob.tspan[0] if sp_ is None else sp_[:, 0], alg,
                                                              bool(grid) and sp_ is None, np.asarray(u0_),
                                                              save_start, save_end is None or save_end,
                                                              dense=dense_of(u0_, p_) if (dense_ok and sp_ is None) else None,
                                                              save_idxs=save_idxs)
        batches_in.append((u0, p, len(I)) if spans is None else (u0, p, len(I), spans))
        sols = mk(res, u0, p, spans)
        if output_func is None:
            batch = sols
        else:
            batch = []
            for k, i in enumerate(I):
                out, rerun = output_func(sols[k], EnsembleContext(int(i), 1))
                repeat = 1
                while rerun:
                    # re-solve this trajectory alone with repeat+1 (the driver's rerun loop)
                    repeat += 1
                    u0r, pr, spr = _harvest(eprob, [int(i)], repeat)
                    r1 = run(u0r, pr, 1, 0, spr)
                    out, rerun = output_func(mk(r1, u0r, pr, spr)[0], EnsembleContext(int(i), repeat))
                batch.append(out)
        if reduction is None:
            if isinstance(u_acc, list) and not u_acc and output_func is None and b0 + batch_size >= N and b0 == 0:
                u_acc = batch                      # single batch, default reduction: keep the lazy view
            else:
                u_acc = list(u_acc) + list(batch)
        else:
            u_acc, converged = reduction(u_acc, batch, I)
            if converged:
                break
    elapsed = time.perf_counter() - t_start
    dense_all = None
    if dense_ok and reduction is None and output_func is None and all(len(b) == 3 for b in batches_in):
        def dense_all(tq):
            return np.concatenate([lowlevel.solve_host_dense(program, u0_, p_, prob.tspan, tq, trajectories=cnt, **tol_kw)["dense"]
                                   for (u0_, p_, cnt) in batches_in])
    return EnsembleSolution(u_acc, elapsed, converged, arrays=all_arrays, dense_all=dense_all)
