"""ctypes binding of include/b200ode.h (libb200ode.so, built in-tree by csrc/build.py).

There is deliberately no fallback: if the shared library is missing, or no sm_100
device is present when a solve is requested, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200ode.so")

ALG_TSIT5, ALG_VERN7, ALG_ROSENBROCK23, ALG_RODAS5P, ALG_DP5, ALG_BS3 = 1, 2, 3, 4, 5, 6
ALG_RODAS5, ALG_RODAS4, ALG_RODAS42, ALG_RODAS4P, ALG_RODAS4P2 = 7, 8, 9, 10, 11
ALG_VERN6, ALG_VERN8, ALG_VERN9, ALG_ROSENBROCK32, ALG_RODAS5PE = 12, 13, 14, 15, 16
ALG_AUTOTSIT5_ROSENBROCK23, ALG_RODAS3P, ALG_RODAS23W = 17, 18, 19
F64, F32 = 0, 1
LAYOUT_AOS, LAYOUT_SOA = 0, 1
FLAG_STATIC_SCHEDULE = 1
FLAG_NO_STEP_ROWS = 2
RC_DEFAULT, RC_SUCCESS, RC_MAXITERS, RC_DTLESSTHANMIN, RC_UNSTABLE, RC_DTNAN, RC_TERMINATED = range(7)
RETCODE_NAMES = {0: "Default", 1: "Success", 2: "MaxIters", 3: "DtLessThanMin", 4: "Unstable", 5: "DtNaN", 6: "Terminated"}
CB_DISCRETE, CB_CONTINUOUS, CB_ISOUTOFDOMAIN = 0, 1, 2
OK, EINVAL, ECOMPILE, ECUDA, EUNSUPPORTED = 0, -1, -2, -3, -4


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200ode error %d: %s" % (code, msg))
        self.code = code


class B200Problem(C.Structure):
    _fields_ = [("trajectories", C.c_int64), ("u0", C.c_void_p), ("u0_shared", C.c_int32),
                ("p", C.c_void_p), ("p_shared", C.c_int32), ("t0", C.c_double), ("tf", C.c_double),
                ("tspans", C.c_void_p)]


class B200Opts(C.Structure):
    _fields_ = [("reltol", C.c_double), ("abstol", C.c_double), ("dt", C.c_double), ("dtmin", C.c_double),
                ("dtmax", C.c_double), ("maxiters", C.c_int64), ("saveat", C.POINTER(C.c_double)),
                ("nsaveat", C.c_int32), ("save_start", C.c_int32), ("save_end", C.c_int32),
                ("flags", C.c_int32), ("reserved", C.c_int32),
                ("tstops", C.POINTER(C.c_double)), ("ntstops", C.c_int32), ("reserved2", C.c_int32),
                ("abstol_vec", C.POINTER(C.c_double)), ("reltol_vec", C.POINTER(C.c_double)),
                ("d_discontinuities", C.POINTER(C.c_double)), ("nd_discontinuities", C.c_int32), ("reserved3", C.c_int32)]


class B200Result(C.Structure):
    _fields_ = [("u_final", C.c_void_p), ("t_final", C.c_void_p), ("us", C.c_void_p), ("ts", C.c_void_p),
                ("nsaved", C.c_void_p), ("naccept", C.c_void_p), ("nreject", C.c_void_p), ("nf", C.c_void_p),
                ("njacs", C.c_void_p), ("nw", C.c_void_p), ("nsolve", C.c_void_p), ("retcode", C.c_void_p),
                ("kernel_ms", C.c_double), ("total_ms", C.c_double)]


class B200DeviceProblem(C.Structure):
    _fields_ = [("trajectories", C.c_int64), ("u0", C.c_void_p), ("u0_shared", C.c_int32), ("u0_layout", C.c_int32),
                ("p", C.c_void_p), ("p_shared", C.c_int32), ("p_layout", C.c_int32),
                ("t0", C.c_double), ("tf", C.c_double), ("tspans", C.c_void_p)]


class B200DeviceResult(C.Structure):
    _fields_ = [("u_final", C.c_void_p), ("u_final_layout", C.c_int32), ("pad0", C.c_int32),
                ("t_final", C.c_void_p), ("us", C.c_void_p),
                ("nsaved", C.c_void_p), ("naccept", C.c_void_p), ("nreject", C.c_void_p), ("nf", C.c_void_p),
                ("njacs", C.c_void_p), ("nw", C.c_void_p), ("nsolve", C.c_void_p), ("retcode", C.c_void_p),
                ("peer_u_final", C.c_void_p * 8), ("npeers", C.c_int32), ("peer_world", C.c_int32), ("peer_rank", C.c_int32),
                ("pad1", C.c_int32), ("peer_block", C.c_int64)]


class B200ProgramInfo(C.Structure):
    _fields_ = [("regs_integrate", C.c_int32), ("regs_initdt", C.c_int32),
                ("local_bytes_integrate", C.c_int32), ("local_bytes_initdt", C.c_int32),
                ("smem_bytes_integrate", C.c_int32), ("block", C.c_int32), ("blocks_per_sm", C.c_int32),
                ("grid", C.c_int32), ("cubin_bytes", C.c_int64), ("compile_ms", C.c_double)]


class B200Ragged(C.Structure):
    _fields_ = [("total_rows", C.c_int64), ("row_offsets", C.c_void_p), ("ts", C.c_void_p), ("us", C.c_void_p)]


class B200CallbackSrc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("rootfind", C.c_int32),
                ("condition_src", C.c_char_p), ("condition_name", C.c_char_p),
                ("affect_src", C.c_char_p), ("affect_name", C.c_char_p),
                ("affect_neg_src", C.c_char_p), ("affect_neg_name", C.c_char_p),
                ("interp_points", C.c_int32), ("save_before", C.c_int32), ("save_after", C.c_int32), ("reserved", C.c_int32),
                ("abstol", C.c_double), ("repeat_nudge", C.c_double)]


def callback_array(callbacks):
    """List of callback dicts -> (B200CallbackSrc array, n).  Keys: kind ("discrete" | "continuous"), condition (src, name),
    affect (src, name) or None, affect_neg (continuous: defaults to affect; False/None-with-key = nothing),
    rootfind ("left" | "right" | "none"), interp_points, abstol, repeat_nudge, save_positions."""
    arr = (B200CallbackSrc * len(callbacks))()
    for i, cb in enumerate(callbacks):
        c = arr[i]
        cont = cb["kind"] == "continuous"
        c.kind = {"continuous": CB_CONTINUOUS, "discrete": CB_DISCRETE, "isoutofdomain": CB_ISOUTOFDOMAIN}[cb["kind"]]
        c.condition_src, c.condition_name = _b(cb["condition"][0]), _b(cb["condition"][1])
        aff = cb.get("affect")
        if aff:
            c.affect_src, c.affect_name = _b(aff[0]), _b(aff[1])
        neg = cb.get("affect_neg", aff if cont else None)
        if cont and neg:
            c.affect_neg_src, c.affect_neg_name = _b(neg[0]), _b(neg[1])
        c.rootfind = {"none": 0, "left": 1, "right": 2}[cb.get("rootfind", "left")]
        c.interp_points = int(cb.get("interp_points", -1))
        c.abstol = float(cb.get("abstol", -1.0))
        c.repeat_nudge = float(cb.get("repeat_nudge", -1.0))
        sp = cb.get("save_positions", (True, True))
        c.save_before, c.save_after = int(bool(sp[0])), int(bool(sp[1]))
    return arr, len(callbacks)


OPT_EVERYSTEP = "-DB200_EVERYSTEP=1"
OPT_TSTOPS = "-DB200_TSTOPS=1"
OPT_TSPANS = "-DB200_TSPANS=1"
OPT_REVERSE_TIME = "-DB200_REVERSE=1"      # tspan[2] < tspan[1] (tdir = -1): mirrored-time kernels
OPT_VECTOR_TOL = "-DB200_VECTOR_TOL=1"
OPT_FIXED_DT = "-DB200_ADAPTIVE=0"
OPT_COMPONENT_RHS = "-DB200_COOP=1"
OPT_SMEM_STAGES = "-DB200_WIDE=1"
OPT_STAGED_SAVEAT = "-DB200_STAGE_ROWS=1"


def opt_save_idxs(idxs):
    """extra_options token selecting the saved components (0-based)."""
    return "-DB200_SAVE_IDXS=" + ",".join(str(int(i)) for i in idxs)


# every symbol include/b200ode.h declares (tests/test_abi.py checks the export list)
EXPORTS = [
    "b200ode_create", "b200ode_destroy", "b200ode_last_error", "b200ode_version",
    "b200ode_compile", "b200ode_program_destroy", "b200ode_program_info",
    "b200ode_compile_only", "b200ode_free", "b200ode_nslots", "b200ode_nslots_program", "b200ode_solve", "b200ode_solve_device",
    "b200ode_reduce_sum_device", "b200ode_timeseries_meanvar_device", "b200ode_solve_meanvar", "b200ode_host_register", "b200ode_host_unregister",
    "b200ode_measure_fma_peak", "b200ode_solve_everystep", "b200ode_solve_everystep_device",
    "b200ode_dense_eval_device", "b200ode_solve_dense", "b200ode_selftest_fastmath",
    "b200ode_multi_create", "b200ode_multi_destroy", "b200ode_multi_device_count", "b200ode_multi_compile",
    "b200ode_multi_program_destroy", "b200ode_multi_solve", "b200ode_multi_reduce_mean",
    "b200ode_compile_callbacks", "b200ode_compile_only_callbacks", "b200ode_struct_size",
]

_lib = None


def lib():
    """Load libb200ode.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libb200ode.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, cp, i32, i64, dbl = C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_double
    L.b200ode_create.argtypes = [C.POINTER(vp), i32]
    L.b200ode_destroy.argtypes = [vp]
    L.b200ode_last_error.argtypes = [vp]
    L.b200ode_last_error.restype = cp
    L.b200ode_version.restype = cp
    L.b200ode_compile.argtypes = [vp, C.POINTER(vp), i32, i32, i32, i32, cp, cp, cp, cp, cp, cp, cp]
    L.b200ode_compile_callbacks.argtypes = [vp, C.POINTER(vp), i32, i32, i32, i32, cp, cp, cp, cp, cp, cp, vp, i32, cp]
    L.b200ode_compile_only_callbacks.argtypes = [i32, i32, i32, i32, cp, cp, vp, i32, cp,
                                                 C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp)]
    L.b200ode_program_destroy.argtypes = [vp]
    L.b200ode_program_info.argtypes = [vp, C.POINTER(B200ProgramInfo)]
    L.b200ode_compile_only.argtypes = [i32, i32, i32, i32, cp, cp, cp, cp, cp, cp, cp,
                                       C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp)]
    L.b200ode_free.argtypes = [vp]
    L.b200ode_free.restype = None
    L.b200ode_nslots.argtypes = [C.POINTER(B200Problem), C.POINTER(B200Opts)]
    L.b200ode_nslots_program.argtypes = [vp, C.POINTER(B200Problem), C.POINTER(B200Opts)]
    L.b200ode_solve.argtypes = [vp, vp, C.POINTER(B200Problem), C.POINTER(B200Opts), C.POINTER(B200Result)]
    L.b200ode_solve_device.argtypes = [vp, vp, C.POINTER(B200DeviceProblem), C.POINTER(B200Opts),
                                       C.POINTER(B200DeviceResult), vp]
    L.b200ode_reduce_sum_device.argtypes = [vp, i32, vp, i32, i64, i32, vp, vp]
    L.b200ode_timeseries_meanvar_device.argtypes = [vp, i32, vp, i64, i32, i32, vp, vp, vp]
    L.b200ode_solve_meanvar.argtypes = [vp, vp, C.POINTER(B200Problem), C.POINTER(B200Opts), C.POINTER(B200Result), vp, vp]
    L.b200ode_solve_everystep.argtypes = [vp, vp, C.POINTER(B200Problem), C.POINTER(B200Opts), C.POINTER(B200Result),
                                          C.POINTER(B200Ragged)]
    L.b200ode_solve_everystep_device.argtypes = [vp, vp, C.POINTER(B200DeviceProblem), C.POINTER(B200Opts),
                                                 C.POINTER(B200DeviceResult), vp, vp, vp, vp]
    L.b200ode_dense_eval_device.argtypes = [vp, vp, i64, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp, C.POINTER(B200Opts), vp]
    L.b200ode_solve_dense.argtypes = [vp, vp, C.POINTER(B200Problem), C.POINTER(B200Opts), vp, i32, vp,
                                      C.POINTER(B200Result)]
    L.b200ode_host_register.argtypes = [vp, C.c_size_t]
    L.b200ode_host_unregister.argtypes = [vp]
    L.b200ode_measure_fma_peak.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.b200ode_selftest_fastmath.argtypes = [vp, i64, C.c_uint64, C.POINTER(i64), C.POINTER(i64)]
    L.b200ode_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i32]
    L.b200ode_multi_destroy.argtypes = [vp]
    L.b200ode_multi_device_count.argtypes = [vp]
    L.b200ode_multi_compile.argtypes = [vp, C.POINTER(vp), i32, i32, i32, i32, cp, cp, cp, cp, cp, cp, cp]
    L.b200ode_multi_program_destroy.argtypes = [vp]
    L.b200ode_multi_solve.argtypes = [vp, vp, C.POINTER(B200Problem), C.POINTER(B200Opts), C.POINTER(B200Result)]
    L.b200ode_multi_reduce_mean.argtypes = [vp, vp, C.POINTER(B200Problem), C.POINTER(B200Opts), C.POINTER(dbl),
                                            C.POINTER(B200Result)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().b200ode_last_error(None)
        raise B200Error(rc, msg.decode("utf-8", "replace") if msg else "")


def _b(s):
    return None if s is None else s.encode("utf-8")


def compile_only(alg, dtype, n, np_, rhs_src, rhs_name, jac_src=None, jac_name=None, tgrad_src=None,
                 tgrad_name=None, extra_options=None, callbacks=None):
    """NVRTC-compile without a GPU; returns (cubin_bytes, log)."""
    L = lib()
    cubin = C.c_void_p()
    size = C.c_size_t()
    log = C.c_void_p()
    if callbacks:
        arr, ncb = callback_array(callbacks)
        rc = L.b200ode_compile_only_callbacks(alg, dtype, n, np_, _b(rhs_src), _b(rhs_name), C.cast(arr, C.c_void_p), ncb,
                                              _b(extra_options), C.byref(cubin), C.byref(size), C.byref(log))
    else:
        rc = L.b200ode_compile_only(alg, dtype, n, np_, _b(rhs_src), _b(rhs_name), _b(jac_src), _b(jac_name),
                                    _b(tgrad_src), _b(tgrad_name), _b(extra_options),
                                    C.byref(cubin), C.byref(size), C.byref(log))
    log_s = C.string_at(log.value).decode("utf-8", "replace") if log.value else ""
    if log.value:
        L.b200ode_free(log)
    if rc != 0:
        if cubin.value:
            L.b200ode_free(cubin)
        raise B200Error(rc, log_s or L.b200ode_last_error(None).decode())
    data = C.string_at(cubin.value, size.value)
    L.b200ode_free(cubin)
    return data, log_s


class Handle:
    """One per process per GPU (b200ode_create)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().b200ode_create(C.byref(self._h), int(device)))
        self.device = int(device)

    def close(self):
        if self._h:
            lib().b200ode_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compile(self, alg, dtype, n, np_, rhs_src, rhs_name, jac_src=None, jac_name=None, tgrad_src=None,
                tgrad_name=None, extra_options=None, callbacks=None):
        return Program(self, alg, dtype, n, np_, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name,
                       extra_options, callbacks)

    def measure_fma_peak(self, dtype=F64):
        tf, mhz = C.c_double(), C.c_double()
        check(lib().b200ode_measure_fma_peak(self._h, dtype, C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value

    def selftest_fastmath(self, samples=1 << 24, seed=1):
        """(mismatches, flagged) of the branch-free division / square-root sequences vs the IEEE operators."""
        bad, fl = (C.c_int64 * 6)(), C.c_int64()
        check(lib().b200ode_selftest_fastmath(self._h, int(samples), int(seed), bad, C.byref(fl)))
        return list(bad), fl.value


class Program:
    def __init__(self, handle, alg, dtype, n, np_, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name,
                 extra_options, callbacks=None):
        self.handle = handle
        self.alg, self.dtype, self.n, self.np = alg, dtype, n, np_
        self.everystep = bool(extra_options) and OPT_EVERYSTEP in extra_options
        self.nsave = n               # components per saved row
        for tok in (extra_options or "").split():
            if tok.startswith("-DB200_SAVE_IDXS="):
                self.nsave = len(tok.split("=", 1)[1].split(","))
        self._p = C.c_void_p()
        self.callbacks = bool(callbacks)
        if callbacks:
            arr, ncb = callback_array(callbacks)
            check(lib().b200ode_compile_callbacks(handle._h, C.byref(self._p), alg, dtype, n, np_, _b(rhs_src), _b(rhs_name),
                                                  _b(jac_src), _b(jac_name), _b(tgrad_src), _b(tgrad_name),
                                                  C.cast(arr, C.c_void_p), ncb, _b(extra_options)))
        else:
            check(lib().b200ode_compile(handle._h, C.byref(self._p), alg, dtype, n, np_, _b(rhs_src), _b(rhs_name),
                                        _b(jac_src), _b(jac_name), _b(tgrad_src), _b(tgrad_name), _b(extra_options)))
        info = B200ProgramInfo()
        check(lib().b200ode_program_info(self._p, C.byref(info)))
        self.info = {k: getattr(info, k) for k, _ in B200ProgramInfo._fields_}

    def close(self):
        if self._p:
            lib().b200ode_program_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiHandle:
    """Several GPUs driven from this process (b200ode_multi_create).  devices: list of device ids, or None = all."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        if devices is None:
            check(lib().b200ode_multi_create(C.byref(self._h), None, 0))
        else:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            check(lib().b200ode_multi_create(C.byref(self._h), ids, len(devices)))
        self.ndev = lib().b200ode_multi_device_count(self._h)

    def compile(self, alg, dtype, n, np_, rhs_src, rhs_name, jac_src=None, jac_name=None, tgrad_src=None,
                tgrad_name=None, extra_options=None):
        return MultiProgram(self, alg, dtype, n, np_, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name,
                            extra_options)

    def close(self):
        if self._h:
            lib().b200ode_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiProgram:
    """One program per device of a MultiHandle; accepted by lowlevel.solve_host / solve_host_mean."""
    multi = True

    def __init__(self, handle, alg, dtype, n, np_, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name,
                 extra_options):
        self.handle = handle
        self.alg, self.dtype, self.n, self.np = alg, dtype, n, np_
        self.everystep = bool(extra_options) and OPT_EVERYSTEP in extra_options
        self.nsave = n
        for tok in (extra_options or "").split():
            if tok.startswith("-DB200_SAVE_IDXS="):
                self.nsave = len(tok.split("=", 1)[1].split(","))
        self._p = C.c_void_p()
        check(lib().b200ode_multi_compile(handle._h, C.byref(self._p), alg, dtype, n, np_, _b(rhs_src), _b(rhs_name),
                                          _b(jac_src), _b(jac_name), _b(tgrad_src), _b(tgrad_name), _b(extra_options)))

    def close(self):
        if self._p:
            lib().b200ode_multi_program_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_opts(reltol=None, abstol=None, dt=None, dtmin=None, dtmax=None, maxiters=None, saveat=None,
              save_start=None, save_end=None, flags=0, tstops=None, d_discontinuities=None):
    """Returns (B200Opts, keepalive)."""
    import numpy as np
    o = B200Opts()
    keep_tol = []
    for name, val in (("reltol", reltol), ("abstol", abstol)):
        if val is not None and np.ndim(val) > 0:        # a vector: one tolerance per component
            arr = np.ascontiguousarray(val, dtype=np.float64)
            keep_tol.append(arr)
            setattr(o, name + "_vec", arr.ctypes.data_as(C.POINTER(C.c_double)))
            setattr(o, name, 0.0)
        else:
            setattr(o, name, float(val) if val is not None else 0.0)
    o.dt = float(dt) if dt is not None else 0.0
    o.dtmin = float(dtmin) if dtmin is not None else 0.0
    o.dtmax = float(dtmax) if dtmax is not None else 0.0
    o.maxiters = int(maxiters) if maxiters is not None else 0
    keep = None
    if saveat is not None and len(saveat) > 0:
        keep = np.ascontiguousarray(saveat, dtype=np.float64)
        o.saveat = keep.ctypes.data_as(C.POINTER(C.c_double))
        o.nsaveat = int(keep.shape[0])
    else:
        o.saveat = None
        o.nsaveat = 0
    o.save_start = -1 if save_start is None else int(bool(save_start))
    o.save_end = -1 if save_end is None else int(bool(save_end))
    o.flags = int(flags)
    if tstops is not None and len(tstops) > 0:
        keep_t = np.ascontiguousarray(tstops, dtype=np.float64)
        o.tstops = keep_t.ctypes.data_as(C.POINTER(C.c_double))
        o.ntstops = int(keep_t.shape[0])
        keep = (keep, keep_t)
    else:
        o.tstops = None
        o.ntstops = 0
    if d_discontinuities is not None and len(d_discontinuities) > 0:
        keep_d = np.ascontiguousarray(d_discontinuities, dtype=np.float64)
        o.d_discontinuities = keep_d.ctypes.data_as(C.POINTER(C.c_double))
        o.nd_discontinuities = int(keep_d.shape[0])
        keep = (keep, keep_d)
    else:
        o.d_discontinuities = None
        o.nd_discontinuities = 0
    return o, (keep, keep_tol)
