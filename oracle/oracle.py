"""ctypes front end of the CPU oracle (oracle.cpp).  TEST INFRASTRUCTURE ONLY — see the
header of oracle.cpp.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the package.

The user's RHS/Jacobian C source (the same text the GPU path hands to NVRTC) is compiled
here with gcc, without floating-point contraction, into a small shared object whose
function pointers are passed to oracle_solve.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_BUILD = os.path.join(_HERE, "_build")


ALG_TSIT5, ALG_VERN7, ALG_ROSENBROCK23, ALG_RODAS5P, ALG_DP5, ALG_BS3 = 1, 2, 3, 4, 5, 6
ALG_RODAS5, ALG_RODAS4, ALG_RODAS42, ALG_RODAS4P, ALG_RODAS4P2 = 7, 8, 9, 10, 11
ALG_VERN6, ALG_VERN8, ALG_VERN9, ALG_ROSENBROCK32, ALG_RODAS5PE, ALG_VERN7_GENERATED = 12, 13, 14, 15, 16, 102
ALG_AUTOTSIT5_ROSENBROCK23, ALG_RODAS3P, ALG_RODAS23W = 17, 18, 19


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".inc"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT)
    return _LIB


class OracleArgs(C.Structure):
    _fields_ = [("alg", C.c_int), ("dtype", C.c_int), ("n", C.c_int), ("np", C.c_int),
                ("rhs", C.c_void_p), ("jac", C.c_void_p), ("tgrad", C.c_void_p),
                ("N", C.c_longlong),
                ("u0", C.c_void_p), ("u0_shared", C.c_int),
                ("p", C.c_void_p), ("p_shared", C.c_int),
                ("t0", C.c_double), ("tf", C.c_double),
                ("reltol", C.c_double), ("abstol", C.c_double), ("dt", C.c_double), ("dtmin", C.c_double),
                ("dtmax", C.c_double), ("maxiters", C.c_longlong),
                ("saveat", C.c_void_p), ("nsaveat", C.c_int),
                ("save_start", C.c_int), ("save_end", C.c_int),
                ("linsolve", C.c_int), ("nthreads", C.c_int),
                ("u_final", C.c_void_p), ("t_final", C.c_void_p), ("us", C.c_void_p), ("nslots", C.c_int),
                ("nsaved", C.c_void_p), ("naccept", C.c_void_p), ("nreject", C.c_void_p), ("nf", C.c_void_p),
                ("njacs", C.c_void_p), ("nw", C.c_void_p), ("nsolve", C.c_void_p), ("retcode", C.c_void_p),
                ("save_everystep", C.c_int), ("row_offsets", C.c_void_p), ("ts_rag", C.c_void_p),
                ("save_idxs", C.c_void_p), ("nsave_idxs", C.c_int),
                ("tstops", C.c_void_p), ("ntstops", C.c_int), ("fixed_dt", C.c_int),
                ("cbs", C.c_void_p), ("ncb", C.c_int), ("abstol_v", C.c_void_p), ("reltol_v", C.c_void_p),
                ("disc", C.c_void_p), ("ndisc", C.c_int), ("tspans", C.c_void_p),
                ("mirror_of_reverse", C.c_int)]


class OracleCallback(C.Structure):
    _fields_ = [("kind", C.c_int), ("condition", C.c_void_p), ("affect", C.c_void_p), ("affect_neg", C.c_void_p),
                ("rootfind", C.c_int), ("interp_points", C.c_int), ("abstol", C.c_double), ("repeat_nudge", C.c_double),
                ("save_before", C.c_int), ("save_after", C.c_int)]


RC_SUCCESS, RC_MAXITERS, RC_DTLESSTHANMIN, RC_UNSTABLE, RC_DTNAN = 1, 2, 3, 4, 5
RC_TERMINATED = 6


def _callback_sources(callbacks):
    out = []
    for cb in callbacks or []:
        for key in ("condition", "affect", "affect_neg"):
            v = cb.get(key)
            if v is not None:
                out.append(v[0])
    return out


def _callback_array(user, callbacks):
    """callbacks: list of dicts — kind ("discrete" | "continuous"), condition (src, name), affect (src, name) or None,
    affect_neg (continuous: defaults to affect; pass False for `nothing`), rootfind ("left" | "right" | "none"),
    interp_points, abstol, repeat_nudge, save_positions.  Continuous callbacks must come first (CallbackSet order)."""
    arr = (OracleCallback * len(callbacks))()
    for i, cb in enumerate(callbacks):
        c = arr[i]
        cont = cb["kind"] == "continuous"
        c.kind = {"continuous": 1, "discrete": 0, "isoutofdomain": 2}[cb["kind"]]
        c.condition = fn_ptr(user, cb["condition"][1])
        aff = cb.get("affect")
        c.affect = fn_ptr(user, aff[1]) if aff else None
        neg = cb.get("affect_neg", aff if cont else None)
        c.affect_neg = fn_ptr(user, neg[1]) if (cont and neg) else None
        c.rootfind = {"none": 0, "left": 1, "right": 2}[cb.get("rootfind", "left")]
        c.interp_points = cb.get("interp_points", 10)
        c.abstol = cb.get("abstol", 10 * 2.0 ** -52)
        c.repeat_nudge = cb.get("repeat_nudge", 0.01)
        sp = cb.get("save_positions", (True, True))
        c.save_before, c.save_after = int(bool(sp[0])), int(bool(sp[1]))
    return arr


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.oracle_solve.argtypes = [C.POINTER(OracleArgs)]
        L.oracle_dense_solve.argtypes = [C.POINTER(OracleArgs), C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_fastpower.restype = C.c_double
        L.oracle_fastpower.argtypes = [C.c_double, C.c_double]
        L.oracle_fastpower_f32.restype = C.c_float
        L.oracle_fastpower_f32.argtypes = [C.c_float, C.c_float]
        L.oracle_norm.restype = C.c_double
        L.oracle_norm.argtypes = [C.c_void_p, C.c_int]
        L.oracle_eps.restype = C.c_double
        L.oracle_eps.argtypes = [C.c_double]
        L.oracle_log10.restype = C.c_double
        L.oracle_log10.argtypes = [C.c_double]
        L.oracle_exp10.restype = C.c_double
        L.oracle_exp10.argtypes = [C.c_double]
        L.oracle_initdt.restype = C.c_double
        L.oracle_initdt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                    C.c_double, C.c_double, C.c_int]
        _lib = L
    return _lib


_user_libs = {}


def compile_user(sources):
    """sources: list of C source strings.  Returns a CDLL with the functions."""
    text = "\n".join(s for s in sources if s)
    key = hashlib.sha1(text.encode()).hexdigest()[:16]
    if key in _user_libs:
        return _user_libs[key]
    os.makedirs(_BUILD, exist_ok=True)
    c_path = os.path.join(_BUILD, "user_%s.c" % key)
    so_path = os.path.join(_BUILD, "user_%s.so" % key)
    if not os.path.exists(so_path):
        with open(c_path, "w") as f:
            f.write("#include <math.h>\n" + text)
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", so_path,
                        c_path, "-lm"], check=True)
    L = C.CDLL(so_path)
    _user_libs[key] = L
    return L


def fn_ptr(user_lib, name):
    return C.cast(getattr(user_lib, name), C.c_void_p).value


def nslots_for(t0, tf, saveat, save_start=None, save_end=None):
    """Rows per trajectory (mirrors solve.jl:141-142,596-599 + skip_saveat_at_tspan_end)."""
    if saveat is None or len(saveat) == 0:
        return 0
    ss = 1 if (save_start is None or save_start) else 0
    se = 1 if (save_end is None or save_end) else 0
    at_tf = sum(1 for s in saveat if s == tf)        # every copy of tf is skipped when save_end = false
    slots = ss + len(saveat)
    if at_tf and not se:
        slots -= at_tf
    if not at_tf and se:
        slots += 1
    return slots


def solve(alg, rhs, u0, p, tspan, n, np_, trajectories=None, f32=False, jac=None, tgrad=None, reltol=None,
          abstol=None, dt=None, dtmin=None, dtmax=None, maxiters=None, saveat=None, save_start=None, save_end=None,
          linsolve=0, nthreads=0, save_everystep=False, dense_tq=None, save_idxs=None, tstops=None, adaptive=True,
          callbacks=None, ragged_saveat=False, d_discontinuities=None, mirror_of_reverse=False):
    """rhs/jac/tgrad: (source, name) tuples.  Arrays as in lowlevel.solve_host.
    save_everystep=True returns ragged rows (row_offsets, ts, us[total, n]) like lowlevel.solve_host_everystep."""
    L = lib()
    rdt = np.float32 if f32 else np.float64
    user = compile_user([rhs[0], jac[0] if jac else None, tgrad[0] if tgrad else None] + _callback_sources(callbacks))
    u0 = np.ascontiguousarray(u0, dtype=rdt)
    u0_shared = u0.ndim == 1
    p_arr = None if p is None else np.ascontiguousarray(p, dtype=rdt)
    p_shared = True if p_arr is None else p_arr.ndim == 1
    tsp = np.asarray(tspan, dtype=np.float64)
    tspans = np.ascontiguousarray(tsp) if tsp.ndim == 2 else None        # (N, 2): per-trajectory spans
    if trajectories is None:
        if not u0_shared:
            trajectories = u0.shape[0]
        elif p_arr is not None and not p_shared:
            trajectories = p_arr.shape[0]
        else:
            trajectories = tspans.shape[0]
    N = int(trajectories)
    if tspans is None:
        t0, tf = float(tspan[0]), float(tspan[1])
    else:
        assert tspans.shape == (N, 2)
        t0, tf = float(tspans[:, 0].min()), float(tspans[:, 1].max())
    grid = None if saveat is None or len(saveat) == 0 else np.ascontiguousarray(saveat, dtype=np.float64)
    # (per-trajectory spans: no rectangular rows — final states, statistics, or the ragged output)
    nslots = nslots_for(t0, tf, grid, save_start, save_end) if tspans is None else 0
    idxs = None if save_idxs is None else np.ascontiguousarray(save_idxs, dtype=np.int32)
    w = n if idxs is None else len(idxs)          # components per saved row (save_idxs, 0-based)
    out = {
        "u_final": np.zeros((N, n), dtype=rdt), "t_final": np.zeros((N,), dtype=rdt),
        "us": np.zeros((N, nslots, w), dtype=rdt) if nslots > 0 else None,
    }
    for k in ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode"):
        out[k] = np.zeros((N,), dtype=np.int32)
    a = OracleArgs()
    a.alg, a.dtype, a.n, a.np = alg, int(f32), n, np_
    a.rhs = fn_ptr(user, rhs[1])
    a.jac = fn_ptr(user, jac[1]) if jac else None
    a.tgrad = fn_ptr(user, tgrad[1]) if tgrad else None
    a.N = N
    a.u0 = u0.ctypes.data; a.u0_shared = int(u0_shared)
    a.p = p_arr.ctypes.data if p_arr is not None else None; a.p_shared = int(p_shared)
    a.t0, a.tf = t0, tf
    a.tspans = tspans.ctypes.data if tspans is not None else None
    # abstol / reltol may be vectors (one entry per component)
    tolv = {}
    for key, val in (("abstol", abstol), ("reltol", reltol)):
        if val is not None and np.ndim(val) > 0:
            tolv[key] = np.ascontiguousarray(val, dtype=np.float64)
            assert tolv[key].shape == (n,)
            setattr(a, key + "_v", tolv[key].ctypes.data)
    a.reltol = 0.0 if (reltol is None or "reltol" in tolv) else reltol
    a.abstol = 0.0 if (abstol is None or "abstol" in tolv) else abstol
    a.dt = dt or 0.0; a.dtmin = dtmin or 0.0
    a.dtmax = dtmax or 0.0; a.maxiters = maxiters or 0
    a.saveat = grid.ctypes.data if grid is not None else None
    a.nsaveat = 0 if grid is None else len(grid)
    a.save_start = -1 if save_start is None else int(bool(save_start))
    a.save_end = -1 if save_end is None else int(bool(save_end))
    a.linsolve = linsolve; a.nthreads = nthreads
    if idxs is not None:
        a.save_idxs = idxs.ctypes.data; a.nsave_idxs = len(idxs)
    a.fixed_dt = 0 if adaptive else 1
    a.mirror_of_reverse = int(bool(mirror_of_reverse))      # test switch (Opts::mirror_of_reverse in oracle.cpp)
    cb_arr = None
    if callbacks:
        cb_arr = _callback_array(user, callbacks)
        a.cbs = C.cast(cb_arr, C.c_void_p).value; a.ncb = len(callbacks)
    stops = None
    if tstops is not None and len(tstops) > 0:
        stops = np.ascontiguousarray(tstops, dtype=np.float64)
        a.tstops = stops.ctypes.data; a.ntstops = len(stops)
    discs = None
    if d_discontinuities is not None and len(d_discontinuities) > 0:
        discs = np.ascontiguousarray(d_discontinuities, dtype=np.float64)
        a.disc = discs.ctypes.data; a.ndisc = len(discs)
    a.u_final = out["u_final"].ctypes.data; a.t_final = out["t_final"].ctypes.data
    a.us = out["us"].ctypes.data if out["us"] is not None else None
    a.nslots = nslots
    for k in ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode"):
        setattr(a, k, out[k].ctypes.data)
    if dense_tq is not None:
        # dense = true (default when save_everystep and no saveat): sol_i(tq[j]) from the stored k arrays
        tq = np.ascontiguousarray(dense_tq, dtype=np.float64)
        dense = np.zeros((N, len(tq), n), dtype=rdt)
        a.save_everystep = 1; a.us = None; a.nslots = 0; a.row_offsets = None; a.ts_rag = None
        rc = L.oracle_dense_solve(C.byref(a), tq.ctypes.data, len(tq), dense.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle_dense_solve failed: %d" % rc)
        out["dense"] = dense
        out["t_final"] = out["t_final"].astype(np.float64)
        return out
    if save_everystep or ragged_saveat:
        # counting pass, exclusive scan, fill pass (the same two passes the GPU path makes).  ragged_saveat: ragged rows
        # (saveat + the rows callbacks force) without the per-step rows, i.e. save_everystep = false with callbacks
        a.save_everystep = 1 if save_everystep else 2; a.us = None; a.nslots = 0; a.row_offsets = None; a.ts_rag = None
        rc = L.oracle_solve(C.byref(a))
        if rc != 0:
            raise RuntimeError("oracle_solve failed: %d" % rc)
        offs = np.zeros((N + 1,), dtype=np.int64)
        np.cumsum(out["nsaved"], out=offs[1:])
        total = int(offs[-1])
        us = np.zeros((max(total, 1), w), dtype=rdt)
        ts = np.zeros((max(total, 1),), dtype=rdt)
        a.us = us.ctypes.data; a.row_offsets = offs.ctypes.data; a.ts_rag = ts.ctypes.data
        rc = L.oracle_solve(C.byref(a))
        if rc != 0:
            raise RuntimeError("oracle_solve failed: %d" % rc)
        out["row_offsets"] = offs; out["us"] = us[:total]; out["ts"] = ts[:total].astype(np.float64)
        out["t_final"] = out["t_final"].astype(np.float64)
        return out
    rc = L.oracle_solve(C.byref(a))
    if rc != 0:
        raise RuntimeError("oracle_solve failed: %d" % rc)
    out["nslots"] = nslots
    if nslots > 0:
        ts = []
        if save_start is None or save_start:
            ts.append(t0)
        for s in grid:
            if s == tf and not (save_end is None or save_end):
                continue
            ts.append(float(rdt(s)))
        if len(ts) < nslots:
            ts.append(tf)
        out["ts"] = np.array(ts)
    out["t_final"] = out["t_final"].astype(np.float64)
    return out
