"""Julia float-range arithmetic, host side.

`saveat = h` makes the reference build `(t0 + h):h:tf` (lib/OrdinaryDiffEqCore/src/solve.jl:1111-1115),
a `StepRangeLen` whose elements are computed in twice precision (Julia Base range.jl /
twiceprecision.jl, EXT to the reference tree).  When start/step/stop have small exact rational
forms (e.g. 0.1:0.1:10.0) element k is the correctly rounded (start_n + k step_n)/den — *not*
start + k*step in floating point.  The Julia binding simply `collect`s the range;
this module reproduces the same values for the Python host mirror and the tests.
"""
from fractions import Fraction
import math

_M = 1 << 24   # maxintfloat(Float32): rat() narrows Float64 to Float32


def _rat(x):
    """Base.rat: continued-fraction rational approximation (returns (n, d); d == 0 on failure)."""
    y = x
    a, d = 1, 1
    b, c = 0, 0
    while abs(y) <= _M:
        f = math.trunc(y)
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if not (max(abs(a), abs(b)) <= _M):
            return c, d
        if b != 0 and a / b == x:
            break
        if y == 0:
            break
        y = 1.0 / y
    return a, b


def _isbetween(a, x, b):
    return (a <= x <= b) or (b <= x <= a)


def julia_range(start, step, stop):
    """collect(start:step:stop) for Float64 arguments."""
    start, step, stop = float(start), float(step), float(stop)
    if step == 0:
        raise ValueError("range step cannot be zero")
    step_n, step_d = _rat(step)
    if step_d != 0 and step_n / step_d == step:
        start_n, start_d = _rat(start)
        stop_n, stop_d = _rat(stop)
        if start_d != 0 and stop_d != 0 and start_n / start_d == start and stop_n / stop_d == stop:
            den = start_d * step_d // math.gcd(start_d, step_d)
            m = float(1 << 53)
            if den != 0 and abs(start * den) <= m and abs(step * den) <= m and den % start_d == 0 and den % step_d == 0:
                sn = round(start * den)
                tn = round(step * den)
                # number of steps that fit; a negative quotient means an empty range
                ln = max(0, math.floor(Fraction(den * stop_n - stop_d * sn, tn * stop_d)) + 1)
                if _isbetween(start, start + (ln - 1) * step, stop + step / 2) and \
                        not _isbetween(start, start + ln * step, stop):
                    return [float(Fraction(sn + k * tn, den)) for k in range(ln)]
    lf = (stop - start) / step
    if lf < 0:
        ln = 0
    elif lf == 0:
        ln = 1
    else:
        ln = int(round(lf)) + 1      # round half to even, like Julia's round(Int, x)
        stop2 = start + (ln - 1) * step
        ln -= int(start < stop < stop2) + int(start > stop > stop2)
    fs, ft = Fraction(start), Fraction(step)
    return [float(fs + k * ft) for k in range(ln)]


def saveat_grid(saveat, tspan):
    """initialize_saveat (lib/OrdinaryDiffEqCore/src/solve.jl:1103-1124).

    Returns the save times in the order the integrator meets them — ascending for tf > t0, descending for a reversed
    span (the reference's heap orders by tdir * t): a number h gives (t0 + tdir |h|):(tdir |h|):tf, a list keeps its
    entries with tdir t0 < tdir t <= tdir tf."""
    t0, tf = float(tspan[0]), float(tspan[1])
    tdir = -1.0 if tf < t0 else 1.0
    if saveat is None:
        return []
    if isinstance(saveat, (int, float)):
        h = tdir * abs(float(saveat))                   # directional_saveat
        return julia_range(t0 + h, h, tf)
    vals = [float(t) for t in saveat]
    kept = [t for t in vals if tdir * t0 < tdir * t <= tdir * tf]
    return kept if tdir > 0 else sorted(kept, reverse=True)


def resolve_save_flags(saveat, tspan, save_everystep, save_start=None, save_end=None):
    """Defaults of save_start / save_end (lib/OrdinaryDiffEqCore/src/solve.jl:141-143,596-599):

        save_start = save_everystep || isempty(saveat) || saveat isa Number || tspan[1] in saveat
        save_end   = (same rule with tspan[2]) when the caller passed nothing

    Returns (save_start, save_end) as the C ABI encodes them: save_start in {0, 1};
    save_end in {0, None (default true: save_end_user is not a Bool), 1 (explicit true)}."""
    t0, tf = float(tspan[0]), float(tspan[1])
    is_number = isinstance(saveat, (int, float))
    empty = saveat is None or (not is_number and len(saveat) == 0)

    def default(endpoint):
        return bool(save_everystep or empty or is_number or any(float(s) == endpoint for s in saveat))
    ss = default(t0) if save_start is None else bool(save_start)
    if save_end is None:
        se = None if default(tf) else False
    else:
        se = bool(save_end)
    return ss, se
