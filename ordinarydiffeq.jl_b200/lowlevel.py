"""Array-level entry points over the C ABI: host buffers (numpy) and device buffers (torch).

`solve_host` is what `solve(EnsembleProblem, alg, EnsembleB200(); ...)` lowers to once
`prob_func` has been harvested into flat (u0, p) tables; `solve_device` is the same call
with everything already resident in HBM (torch tensors are used purely as device
memory owners here).
"""
import ctypes as C

import numpy as np

from . import _lib


def _np_real(dtype):
    return np.float32 if dtype == _lib.F32 else np.float64


def _per_trajectory_spans(tspan):
    """None for a shared (t0, tf); the contiguous float64 (N, 2) array when tspan holds one span per trajectory."""
    a = np.asarray(tspan, dtype=np.float64)
    if a.ndim == 1:
        return None
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError("tspan must be (t0, tf) or an (N, 2) array of per-trajectory spans")
    return np.ascontiguousarray(a)


def solve_host(program, u0, p, tspan, trajectories=None, reltol=None, abstol=None, dt=None, dtmin=None, dtmax=None,
               maxiters=None, saveat=None, save_start=None, save_end=None, flags=0, out=None, _meanvar=None, tstops=None, _mean=None,
               d_discontinuities=None):
    """Host arrays in, host arrays out (b200ode_solve; b200ode_multi_solve for a MultiProgram).

    u0: (N, n) or (n,) shared; p: (N, np) or (np,) shared or None.
    saveat: explicit ascending grid in (t0, tf] or None.
    out: optional dict of preallocated (e.g. pinned) numpy arrays keyed like the result.
    """
    L = _lib.lib()
    rdt = _np_real(program.dtype)
    n, npar = program.n, program.np
    u0 = np.ascontiguousarray(u0, dtype=rdt)
    u0_shared = (u0.ndim == 1)
    if p is None:
        p_arr, p_shared = None, True
    else:
        p_arr = np.ascontiguousarray(p, dtype=rdt)
        p_shared = (p_arr.ndim == 1)
    tspans = _per_trajectory_spans(tspan)       # (N, 2) array: every trajectory its own (t0_i, tf_i)
    if trajectories is None:
        if not u0_shared:
            trajectories = u0.shape[0]
        elif p_arr is not None and not p_shared:
            trajectories = p_arr.shape[0]
        elif tspans is not None:
            trajectories = tspans.shape[0]
        else:
            raise ValueError("trajectories must be given when both u0 and p are shared")
    N = int(trajectories)
    if u0.shape[-1] != n or (not u0_shared and u0.shape[0] != N):
        raise ValueError("u0 has shape %s, expected (%d, %d) or (%d,)" % (u0.shape, N, n, n))
    if npar > 0:
        if p_arr is None or p_arr.shape[-1] != npar or (not p_shared and p_arr.shape[0] != N):
            raise ValueError("p has wrong shape for np=%d" % npar)
    if tspans is not None and tspans.shape[0] != N:
        raise ValueError("tspan has %d rows for %d trajectories" % (tspans.shape[0], N))
    t0, tf = (float(tspan[0]), float(tspan[1])) if tspans is None else (float(tspans[:, 0].min()), float(tspans[:, 1].max()))
    opts, keep = _lib.make_opts(reltol, abstol, dt, dtmin, dtmax, maxiters, saveat, save_start, save_end, flags, tstops, d_discontinuities)
    prob = _lib.B200Problem()
    prob.trajectories = N
    prob.u0 = u0.ctypes.data
    prob.u0_shared = int(u0_shared)
    prob.p = p_arr.ctypes.data if p_arr is not None else None
    prob.p_shared = int(p_shared)
    prob.t0, prob.tf = t0, tf
    prob.tspans = tspans.ctypes.data if tspans is not None else None
    if getattr(program, "multi", False):
        # (a MultiProgram holds one program per device; the F64 rule differs from the F32 one only for grid points
        # within a float ulp of tf)
        nslots = L.b200ode_nslots(C.byref(prob), C.byref(opts))
    else:
        nslots = L.b200ode_nslots_program(program._p, C.byref(prob), C.byref(opts))
    if tspans is not None:
        nslots = 0          # per-trajectory spans: no rectangular rows (final states and statistics only; rows via the ragged output)
    out = {} if out is None else out

    def buf(name, shape, dtype):
        a = out.get(name)
        if a is None:
            a = np.empty(shape, dtype=dtype)
            out[name] = a
        assert a.shape == tuple(shape) and a.dtype == dtype and a.flags["C_CONTIGUOUS"], name
        return a

    res = _lib.B200Result()
    res.u_final = buf("u_final", (N, n), rdt).ctypes.data
    res.t_final = buf("t_final", (N,), np.float64).ctypes.data
    if nslots > 0:
        if _meanvar is None:
            res.us = buf("us", (N, nslots, program.nsave), rdt).ctypes.data
        res.ts = buf("ts", (nslots,), np.float64).ctypes.data
    for name in ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode"):
        setattr(res, name, buf(name, (N,), np.int32).ctypes.data)
    if _meanvar is not None:
        mean = buf("mean", (nslots, program.nsave), np.float64)
        var = buf("var", (nslots, program.nsave), np.float64) if _meanvar[1] else None
        _lib.check(L.b200ode_solve_meanvar(program.handle._h, program._p, C.byref(prob), C.byref(opts), C.byref(res),
                                           C.c_void_p(mean.ctypes.data), C.c_void_p(var.ctypes.data) if var is not None else None))
    elif _mean is not None:
        mean = buf("mean", (n,), np.float64)
        _lib.check(L.b200ode_multi_reduce_mean(program.handle._h, program._p, C.byref(prob), C.byref(opts),
                                               mean.ctypes.data_as(C.POINTER(C.c_double)), C.byref(res)))
    elif getattr(program, "multi", False):
        _lib.check(L.b200ode_multi_solve(program.handle._h, program._p, C.byref(prob), C.byref(opts), C.byref(res)))
    else:
        _lib.check(L.b200ode_solve(program.handle._h, program._p, C.byref(prob), C.byref(opts), C.byref(res)))
    out["kernel_ms"] = res.kernel_ms
    out["total_ms"] = res.total_ms
    out["nslots"] = nslots
    return out


def _marshal_ragged(program, u0, p, tspan, trajectories):
    """Problem + scalar result structs shared by the save_everystep / dense entry points."""
    if not program.everystep:
        raise ValueError("program was not compiled with OPT_EVERYSTEP")
    rdt = _np_real(program.dtype)
    n, npar = program.n, program.np
    u0 = np.ascontiguousarray(u0, dtype=rdt)
    u0_shared = (u0.ndim == 1)
    p_arr = None if p is None else np.ascontiguousarray(p, dtype=rdt)
    p_shared = True if p_arr is None else (p_arr.ndim == 1)
    tspans = _per_trajectory_spans(tspan)
    if trajectories is None:
        if not u0_shared:
            trajectories = u0.shape[0]
        elif p_arr is not None and not p_shared:
            trajectories = p_arr.shape[0]
        elif tspans is not None:
            trajectories = tspans.shape[0]
        else:
            raise ValueError("trajectories must be given when both u0 and p are shared")
    N = int(trajectories)
    if u0.shape[-1] != n or (not u0_shared and u0.shape[0] != N):
        raise ValueError("u0 has shape %s, expected (%d, %d) or (%d,)" % (u0.shape, N, n, n))
    if npar > 0 and (p_arr is None or p_arr.shape[-1] != npar or (not p_shared and p_arr.shape[0] != N)):
        raise ValueError("p has wrong shape for np=%d" % npar)
    if tspans is not None and tspans.shape[0] != N:
        raise ValueError("tspan has %d rows for %d trajectories" % (tspans.shape[0], N))
    prob = _lib.B200Problem()
    prob.trajectories = N
    prob.u0 = u0.ctypes.data; prob.u0_shared = int(u0_shared)
    prob.p = p_arr.ctypes.data if p_arr is not None else None; prob.p_shared = int(p_shared)
    if tspans is None:
        prob.t0, prob.tf = float(tspan[0]), float(tspan[1])
    else:
        prob.t0, prob.tf = float(tspans[:, 0].min()), float(tspans[:, 1].max())
        prob.tspans = tspans.ctypes.data
    out = {"u_final": np.empty((N, n), dtype=rdt), "t_final": np.empty((N,), dtype=np.float64)}
    res = _lib.B200Result()
    res.u_final = out["u_final"].ctypes.data
    res.t_final = out["t_final"].ctypes.data
    for name in ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode"):
        out[name] = np.empty((N,), dtype=np.int32)
        setattr(res, name, out[name].ctypes.data)
    return N, n, rdt, prob, res, out, (u0, p_arr, tspans)


def solve_host_dense(program, u0, p, tspan, tq, trajectories=None, reltol=None, abstol=None, dt=None, dtmin=None,
                     dtmax=None, maxiters=None, flags=0):
    """sol_i(tq[j]) for every trajectory (b200ode_solve_dense): integrates with save_everystep, evaluates the
    dense output on the device, returns dict with dense[N, len(tq), n] and the per-trajectory scalars."""
    L = _lib.lib()
    N, n, rdt, prob, res, out, keep_in = _marshal_ragged(program, u0, p, tspan, trajectories)
    opts, keep = _lib.make_opts(reltol, abstol, dt, dtmin, dtmax, maxiters, None, None, None, flags)
    tq = np.ascontiguousarray(tq, dtype=np.float64)
    out["dense"] = np.empty((N, len(tq), n), dtype=rdt)
    _lib.check(L.b200ode_solve_dense(program.handle._h, program._p, C.byref(prob), C.byref(opts),
                                     C.c_void_p(tq.ctypes.data), len(tq), C.c_void_p(out["dense"].ctypes.data),
                                     C.byref(res)))
    out["kernel_ms"] = res.kernel_ms
    out["total_ms"] = res.total_ms
    return out


def solve_host_everystep(program, u0, p, tspan, trajectories=None, reltol=None, abstol=None, dt=None, dtmin=None,
                         dtmax=None, maxiters=None, saveat=None, save_start=None, save_end=None, flags=0, tstops=None, d_discontinuities=None):
    """save_everystep = true (b200ode_solve_everystep): returns the per-trajectory scalars plus the ragged
    rows — row_offsets[N+1], ts[total], us[total, n]; trajectory i's sol.t / sol.u are
    ts[row_offsets[i]:row_offsets[i+1]] and the same slice of us."""
    L = _lib.lib()
    N, n, rdt, prob, res, out, keep_in = _marshal_ragged(program, u0, p, tspan, trajectories)
    opts, keep = _lib.make_opts(reltol, abstol, dt, dtmin, dtmax, maxiters, saveat, save_start, save_end, flags, tstops, d_discontinuities)
    rag = _lib.B200Ragged()
    _lib.check(L.b200ode_solve_everystep(program.handle._h, program._p, C.byref(prob), C.byref(opts), C.byref(res),
                                         C.byref(rag)))
    try:
        total = int(rag.total_rows)
        if N > 0:
            out["row_offsets"] = np.ctypeslib.as_array(C.cast(rag.row_offsets, C.POINTER(C.c_int64)), (N + 1,)).copy()
        else:
            out["row_offsets"] = np.zeros((1,), dtype=np.int64)
        if total > 0:
            out["ts"] = np.ctypeslib.as_array(C.cast(rag.ts, C.POINTER(C.c_double)), (total,)).copy()
            ct = C.c_float if rdt == np.float32 else C.c_double
            w = program.nsave
            out["us"] = np.ctypeslib.as_array(C.cast(rag.us, C.POINTER(ct)), (total * w,)).reshape(total, w).copy()
        else:
            out["ts"] = np.zeros((0,), dtype=np.float64)
            out["us"] = np.zeros((0, program.nsave), dtype=rdt)
    finally:
        for ptr in (rag.row_offsets, rag.ts, rag.us):
            if ptr:
                L.b200ode_free(C.c_void_p(ptr))
    out["kernel_ms"] = res.kernel_ms
    out["total_ms"] = res.total_ms
    return out


def solve_host_meanvar(program, u0, p, tspan, saveat, trajectories=None, want_var=True, **kw):
    """EnsembleAnalysis.timeseries_steps_meanvar without moving the trajectories to the host
    (b200ode_solve_meanvar): returns dict with ts, mean[nslots][n], var[nslots][n] and the
    per-trajectory scalars (u_final, retcode, counters)."""
    return solve_host(program, u0, p, tspan, trajectories=trajectories, saveat=saveat, _meanvar=(True, want_var), **kw)


def solve_host_mean(program, u0, p, tspan, trajectories=None, **kw):
    """Ensemble mean of u(tf) over all trajectories, sharded over the devices of a MultiProgram
    (b200ode_multi_reduce_mean); the per-trajectory results are returned as well."""
    if not getattr(program, "multi", False):
        raise ValueError("solve_host_mean needs a MultiProgram (MultiHandle.compile)")
    return solve_host(program, u0, p, tspan, trajectories=trajectories, _mean=True, **kw)


def nslots_for(tspan, saveat, save_start=None, save_end=None):
    L = _lib.lib()
    opts, keep = _lib.make_opts(saveat=saveat, save_start=save_start, save_end=save_end)
    prob = _lib.B200Problem()
    prob.t0, prob.tf = float(tspan[0]), float(tspan[1])
    return L.b200ode_nslots(C.byref(prob), C.byref(opts))


class DeviceBuffers:
    """Device-resident inputs/outputs for `solve_device` (torch tensors own the memory)."""

    def __init__(self, program, N, nslots, device, u0_shared=False, p_shared=False, layout=_lib.LAYOUT_AOS):
        import torch
        rt = torch.float32 if program.dtype == _lib.F32 else torch.float64
        n, npar = program.n, program.np
        self.N, self.nslots, self.layout = N, nslots, layout
        self.u0_shared, self.p_shared = u0_shared, p_shared
        shape_u0 = (n,) if u0_shared else ((N, n) if layout == _lib.LAYOUT_AOS else (n, N))
        shape_p = (max(npar, 1),) if p_shared else ((N, max(npar, 1)) if layout == _lib.LAYOUT_AOS else (max(npar, 1), N))
        self.u0 = torch.empty(shape_u0, dtype=rt, device=device)
        self.p = torch.empty(shape_p, dtype=rt, device=device)
        self.u_final = torch.empty((N, n) if layout == _lib.LAYOUT_AOS else (n, N), dtype=rt, device=device)
        self.t_final = torch.empty((N,), dtype=rt, device=device)
        self.us = torch.empty((N, nslots, program.nsave), dtype=rt, device=device) if nslots > 0 else None
        self.i32 = torch.zeros((8, N), dtype=torch.int32, device=device)
        names = ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode")
        for k, name in enumerate(names):
            setattr(self, name, self.i32[k])


def solve_device(program, bufs, tspan, reltol=None, abstol=None, dt=None, dtmin=None, dtmax=None, maxiters=None,
                 saveat=None, save_start=None, save_end=None, flags=0, stream=None, tstops=None, d_discontinuities=None,
                 first=0, count=None, peer_out=None):
    """Launch the ensemble on buffers already in HBM (b200ode_solve_device); asynchronous.
    first / count: only trajectories first .. first+count-1 of the buffers (array-of-structures layout) — lets a caller
    split one ensemble into several launches, e.g. to overlap a collective on the finished part with the rest."""
    L = _lib.lib()
    opts, keep = _lib.make_opts(reltol, abstol, dt, dtmin, dtmax, maxiters, saveat, save_start, save_end, flags, tstops, d_discontinuities)
    first = int(first)
    count = bufs.N - first if count is None else int(count)
    if first != 0 or count != bufs.N:
        if bufs.layout != _lib.LAYOUT_AOS:
            raise ValueError("a sub-range launch needs the array-of-structures layout")
        if first < 0 or count < 0 or first + count > bufs.N:
            raise ValueError("sub-range outside the buffers")

    def row(t, shared=False):      # device address of row `first`
        return t.data_ptr() if (shared or first == 0) else t[first:].data_ptr()
    dp = _lib.B200DeviceProblem()
    dp.trajectories = count
    dp.u0 = row(bufs.u0, bufs.u0_shared); dp.u0_shared = int(bufs.u0_shared); dp.u0_layout = bufs.layout
    dp.p = row(bufs.p, bufs.p_shared); dp.p_shared = int(bufs.p_shared); dp.p_layout = bufs.layout
    dp.t0, dp.tf = float(tspan[0]), float(tspan[1])
    dr = _lib.B200DeviceResult()
    dr.u_final = row(bufs.u_final); dr.u_final_layout = bufs.layout
    dr.t_final = row(bufs.t_final)
    dr.us = row(bufs.us) if bufs.us is not None else None
    for name in ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode"):
        setattr(dr, name, row(getattr(bufs, name)))
    if peer_out is not None:
        # (device pointers of every rank's result [total, n], world, rank, block): the kernel stores each final state at
        # its global index into all of them (B200DeviceResult.peer_u_final) — the ordered gather, fused
        ptrs, world, rank, block = peer_out
        if first != 0:
            raise ValueError("the fused gather numbers the launch's trajectories from the start of the shard")
        for r, ptr in enumerate(ptrs):
            dr.peer_u_final[r] = int(ptr)
        dr.npeers, dr.peer_world, dr.peer_rank, dr.peer_block = len(ptrs), int(world), int(rank), int(block)
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    _lib.check(L.b200ode_solve_device(program.handle._h, program._p, C.byref(dp), C.byref(opts), C.byref(dr),
                                      C.c_void_p(stream)))


def reduce_sum_device(handle, dtype, x, layout, count, n, out, stream=None):
    """out (torch float64 [n], device) = per-component sum over trajectories of x."""
    L = _lib.lib()
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    _lib.check(L.b200ode_reduce_sum_device(handle._h, dtype, C.c_void_p(x.data_ptr()), layout, int(count), int(n),
                                           C.c_void_p(out.data_ptr()), C.c_void_p(stream)))


def timeseries_meanvar_device(handle, dtype, us, mean, var=None, stream=None):
    """mean/var (torch float64 [nslots, n], device) over the trajectories of us [N, nslots, n]."""
    L = _lib.lib()
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    N, nslots, n = us.shape
    _lib.check(L.b200ode_timeseries_meanvar_device(handle._h, dtype, C.c_void_p(us.data_ptr()), int(N), int(nslots), int(n),
                                                   C.c_void_p(mean.data_ptr()),
                                                   C.c_void_p(var.data_ptr()) if var is not None else None, C.c_void_p(stream)))


def solve_everystep_device(program, bufs, tspan, row_offsets=None, ts=None, dts=None, us=None, reltol=None, abstol=None,
                           dt=None, dtmin=None, dtmax=None, maxiters=None, saveat=None, save_start=None, save_end=None,
                           flags=0, stream=None, tstops=None, d_discontinuities=None):
    """Device-resident save_everystep (b200ode_solve_everystep_device); asynchronous.

    Counting pass: row_offsets=None — fills bufs.nsaved.  Fill pass: row_offsets = int64[N+1] exclusive scan of those
    counts (e.g. torch.cumsum on the device), ts/dts = real[total], us = real[total, nsave], all device tensors."""
    L = _lib.lib()
    opts, keep = _lib.make_opts(reltol, abstol, dt, dtmin, dtmax, maxiters, saveat, save_start, save_end, flags, tstops, d_discontinuities)
    dp = _lib.B200DeviceProblem()
    dp.trajectories = bufs.N
    dp.u0 = bufs.u0.data_ptr(); dp.u0_shared = int(bufs.u0_shared); dp.u0_layout = bufs.layout
    dp.p = bufs.p.data_ptr(); dp.p_shared = int(bufs.p_shared); dp.p_layout = bufs.layout
    dp.t0, dp.tf = float(tspan[0]), float(tspan[1])
    dr = _lib.B200DeviceResult()
    dr.u_final = bufs.u_final.data_ptr(); dr.u_final_layout = bufs.layout
    dr.t_final = bufs.t_final.data_ptr()
    dr.us = us.data_ptr() if us is not None else None
    for name in ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode"):
        setattr(dr, name, getattr(bufs, name).data_ptr())
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    _lib.check(L.b200ode_solve_everystep_device(
        program.handle._h, program._p, C.byref(dp), C.byref(opts), C.byref(dr),
        C.c_void_p(row_offsets.data_ptr()) if row_offsets is not None else None,
        C.c_void_p(ts.data_ptr()) if ts is not None else None,
        C.c_void_p(dts.data_ptr()) if dts is not None else None, C.c_void_p(stream)))


def dense_eval_device(program, N, p, row_offsets, ts, dts, us, tq, out, p_shared=False, layout=_lib.LAYOUT_AOS,
                      reltol=None, abstol=None, stream=None):
    """sol_i(tq[j]) from device-resident ragged rows (b200ode_dense_eval_device); tq ascending, out = real[N, len(tq), n]."""
    L = _lib.lib()
    opts, keep = _lib.make_opts(reltol, abstol)
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    _lib.check(L.b200ode_dense_eval_device(
        program.handle._h, program._p, int(N), C.c_void_p(p.data_ptr()) if p is not None else None, int(p_shared),
        int(layout), C.c_void_p(row_offsets.data_ptr()), C.c_void_p(ts.data_ptr()), C.c_void_p(dts.data_ptr()),
        C.c_void_p(us.data_ptr()), C.c_void_p(tq.data_ptr()), int(tq.numel()), C.c_void_p(out.data_ptr()),
        C.byref(opts), C.c_void_p(stream)))
