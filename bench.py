#!/usr/bin/env python
"""bench.py — the ensemble hot path on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one pass of the hot path over one batch: solve a 2^20-trajectory Lorenz/Tsit5 FP64
ensemble with saveat = 0.1 (BASELINE.json configs[1], the configuration `metric` is quoted on).
Prints ONE JSON line (rank 0):
  value      trajectories/s, whole job (weak scaling: every rank its own 2^20 trajectories), inputs resident in HBM,
             CUDA-event time, max over ranks
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H timed; with the
             cudaMemcpy-only D2H time of the same bytes beside it (the PCIe ceiling) and labelled variants
             (FP32, save_idxs, mean/var on the device)
  roofline   FP64-FMA roofline of b200_integrate (algorithmic flops / CUDA-event time vs the FMA
             peak measured live on this GPU) + the HBM side of the saveat stream
  cpu_baseline  the CPU oracle (port of the reference algorithm) on this box's host cores, bounded sample
  configs    the other BASELINE.json configurations (FP32, final-state only, Robertson Rodas5P / Rosenbrock23,
             Pleiades Vern7), each with value, roofline and cpu_baseline (short runs; N = 1 only)
  strong_1m  ONE 2^20-trajectory saveat ensemble sharded over the N ranks (interleaved blocks), per-rank kernel ms
  sweep_64m  BASELINE configs[4]: the 64 Mi-trajectory rho sweep sharded over the N ranks with the ordered gather of
             the final states and the ensemble mean inside the timed step, NCCL time separated
`--impl reference` times the CPU oracle alone (the reference is Julia and cannot run here).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# Algorithmic FP64/FP32 flops per unit of work (FMA = 2; add/mul/div/sqrt = 1; abs/max/compare/select = 0; the
# scalar controller is excluded) — SURVEY §8(d), derivations in DESIGN.md §4:
#   Tsit5/Lorenz      3*61 stage/solution/error combinations + 6*8 RHS + 9 residuals + 7 norm          = 247 / attempt
#   Tsit5 interpolant Theta 2 + Theta^2 1 + b1 7 + b2..b7 30 + 3*15                                     = 85 / row
#   Rosenbrock23/ROBER (n = 3: RHS 13, J 10, W 4, 3x3 inverse 38, mat-vec solve 15, norm 16)
#       4 + 52 + stage1 24 + stage2 44 + u,RHS 20 + stage3 42 + error 28                                = 214 / attempt
#   Rodas5P/ROBER     1 + 52 + 1 + 13 + 7 + 15 + sum_{s=1..7}(13 s + 40) + 48 + 6 + 16                   = 803 / attempt
#   Vern7/Pleiades    28*111 + 10*588 + 84 + 57                                                          = 9129 / attempt
F_INIT = 60.0                   # initdt: 2 RHS + 3 norms (Lorenz); negligible elsewhere
FLOPS = {("lorenz", "tsit5"): dict(step=247.0, row=85.0), ("lorenz_sweep", "tsit5"): dict(step=247.0, row=85.0),
         ("robertson", "ros23"): dict(step=214.0, row=0.0), ("robertson", "rodas5p"): dict(step=803.0, row=0.0),
         ("pleiades", "vern7"): dict(step=9129.0, row=0.0)}
FLOPS_NOTE = {("lorenz", "tsit5"): "247/attempted step + 85/interpolated row + 60/trajectory",
              ("lorenz_sweep", "tsit5"): "247/attempted step + 60/trajectory",
              ("robertson", "ros23"): "214/attempted step (RHS 13, J 10, W 4, 3x3 inverse 38, 3 mat-vec solves 15, norm 16)",
              ("robertson", "rodas5p"): "803/attempted step (8 stages: sum(13 s + 40), 8 mat-vec solves)",
              ("pleiades", "vern7"): "9129/attempted step (28*111 + 10*588 + 84 + 57)"}

WORKLOADS = {
    # name: (problem, alg, f32, N per GPU, saveat h or None, tspan, tolerances)
    "lorenz_tsit5_saveat_1m": dict(problem="lorenz", alg="tsit5", f32=False, N=1 << 20, saveat=0.1, tspan=(0.0, 10.0), tol={}),
    "lorenz_tsit5_saveat_1m_f32": dict(problem="lorenz", alg="tsit5", f32=True, N=1 << 20, saveat=0.1, tspan=(0.0, 10.0), tol={}),
    "lorenz_tsit5_final_1m": dict(problem="lorenz", alg="tsit5", f32=False, N=1 << 20, saveat=None, tspan=(0.0, 10.0), tol={}),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (10 k trajectories, reltol 1e-8) — latency-, not
    # throughput-bound on a GPU: 10 k trajectories are two warps per SM
    "lorenz_tsit5_10k_reltol1e-8": dict(problem="lorenz", alg="tsit5", f32=False, N=10000, saveat=None, tspan=(0.0, 10.0),
                                        tol=dict(reltol=1e-8)),
    "robertson_rodas5p_1m": dict(problem="robertson", alg="rodas5p", f32=False, N=1 << 20, saveat=None, tspan=(0.0, 1e5),
                                 tol=dict(reltol=1e-6, abstol=1e-8)),
    "robertson_rosenbrock23_1m": dict(problem="robertson", alg="ros23", f32=False, N=1 << 20, saveat=None, tspan=(0.0, 1e5),
                                      tol=dict(reltol=1e-6, abstol=1e-8)),
    # BASELINE.json configs[4]: one 64 Mi-trajectory parameter sweep sharded over the ranks (strong scaling),
    # NCCL all-gather of the final states into trajectory order + ensemble mean (all-reduce of partial sums)
    "lorenz_sweep_64m": dict(problem="lorenz_sweep", alg="tsit5", f32=False, N=1 << 26, saveat=None, tspan=(0.0, 10.0), tol={}),
    "pleiades_vern7_256k": dict(problem="pleiades", alg="vern7", f32=False, N=1 << 18, saveat=None, tspan=(0.0, 3.0),
                                tol=dict(reltol=1e-6, abstol=1e-8)),
    # not a BASELINE configuration: a mixed stiff / non-stiff ensemble for the switching algorithm —
    # Van der Pol with mu log-uniform in [0.5, 500]: lanes that stay in Tsit5, lanes that move to Rosenbrock23 for good and
    # lanes that switch back and forth share warps
    "vdp_autotsit5_mixed_256k": dict(problem="vdp_mixed", alg="autotsit5", f32=False, N=1 << 18, saveat=None, tspan=(0.0, 20.0),
                                     tol={}),
}
OTHER_CONFIGS = ["lorenz_tsit5_10k_reltol1e-8", "lorenz_tsit5_saveat_1m_f32", "lorenz_tsit5_final_1m", "robertson_rodas5p_1m", "robertson_rosenbrock23_1m",
                 "pleiades_vern7_256k", "vdp_autotsit5_mixed_256k"]
ALG_NAMES = {"tsit5": "ALG_TSIT5", "vern7": "ALG_VERN7", "ros23": "ALG_ROSENBROCK23", "rodas5p": "ALG_RODAS5P",
             "autotsit5": "ALG_AUTOTSIT5_ROSENBROCK23"}


def sources(pl, w, device=False):
    """(rhs, jac, tgrad, n, np, extra compile options).  device=True: the form the GPU program is compiled from
    (Pleiades: the pair-shared full-vector text for the shared-memory stage kernel, bit-identical to the reference loop);
    otherwise the plain full-vector form the CPU oracle compiles."""
    f32 = w["f32"]
    if w["problem"] in ("lorenz", "lorenz_sweep"):
        return pl.lorenz_source(f32), None, None, 3, 3, None
    if w["problem"] == "robertson":
        r, j, tg = pl.robertson_sources(f32)
        return r, j, tg, 3, 3, None
    if w["problem"] == "vdp_mixed":
        r, j, tg, n, np_, _, _ = pl.stiff_sources("vdp", f32)
        return r, j, tg, n, np_, None
    if w["problem"] == "pleiades":
        if device:
            return pl.pleiades_pairs_source(f32), None, None, 28, 0, "-DB200_WIDE=1 -DB200_WIDE_WINDOW=4"
        return pl.pleiades_source(f32, loops=True), None, None, 28, 0, None
    raise ValueError(w["problem"])


def inputs(pl, w, N, offset=0, idx=None, total=None):
    """u0 (shared or per trajectory) and p for trajectories offset..offset+N-1, or for the global indices `idx`."""
    f32 = w["f32"]
    if idx is None:
        idx = np.arange(offset, offset + N, dtype=np.uint64)
    idx = np.asarray(idx, dtype=np.uint64)
    if w["problem"] in ("lorenz", "lorenz_sweep"):
        p = np.empty((idx.shape[0], 3), dtype=np.float64)
        p[:, 0] = 10.0
        p[:, 2] = 8.0 / 3.0
        if w["problem"] == "lorenz_sweep":
            p[:, 1] = 14.0 + 28.0 * idx.astype(np.float64) / float(total)
        else:
            p[:, 1] = 28.0 * (0.5 + pl.splitmix64_uniform(idx, 0))
        return np.array([1.0, 0.0, 0.0]), (p.astype(np.float32) if f32 else p)
    if w["problem"] == "vdp_mixed":
        mu = 0.5 * (1000.0 ** pl.splitmix64_uniform(idx, 0))
        p = mu.reshape(-1, 1)
        return np.array([1.0, 1.0]), (p.astype(np.float32) if f32 else p)
    if w["problem"] == "robertson":
        base = np.array([0.04, 3.0e7, 1.0e4])
        p = np.empty((idx.shape[0], 3), dtype=np.float64)
        for j in range(3):
            p[:, j] = base[j] * (0.5 + pl.splitmix64_uniform(idx, j))
        return np.array([1.0, 0.0, 0.0]), (p.astype(np.float32) if f32 else p)
    return pl.pleiades_u0(idx.shape[0], offset=int(idx[0]), f32=f32), None


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons sampled DURING the timed region (NVML, ~2 ms period;
    the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain index list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            try:
                phys = int(vis.split(",")[index])
            except Exception:
                phys = index
        self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))

    def run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((float(sm), pw, int(rs)))
            except Exception:
                pass
            time.sleep(0.002)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        return self.summary()

    def summary(self):
        nv = self.nv
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for s in self.samples:
            bits |= s[2]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, v in names.items() if bits & v)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples), "power_w_max": max(s[1] for s in self.samples)}


def cpu_oracle_rate(pl, w, seconds, nthreads=0):
    """Time the CPU oracle on a bounded sample of the workload (same inputs, same options).
    Returns (traj_per_s, sample, threads, steps_per_traj)."""
    from oracle import oracle
    rhs, jac, tg, n, np_, _ = sources(pl, w)
    alg = getattr(oracle, ALG_NAMES[w["alg"]])
    import b200_import
    pkg = b200_import.load()
    grid = pkg.ranges.saveat_grid(w["saveat"], w["tspan"]) if w["saveat"] is not None else None
    if nthreads == 0:
        # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1; override it)
        try:
            nthreads = len(os.sched_getaffinity(0))
        except Exception:
            nthreads = os.cpu_count() or 1
    threads = nthreads
    total = w["N"]

    def run(N):
        u0, p = inputs(pl, w, N, 0, total=total)
        t = time.perf_counter()
        o = oracle.solve(alg, rhs, u0, p, w["tspan"], n, np_, f32=w["f32"], jac=jac, tgrad=tg, saveat=grid,
                         nthreads=nthreads, **w["tol"])
        return time.perf_counter() - t, o
    probe = 64 * threads
    run(min(probe, w["N"]))                    # warm: compiles the user source, spins up the OpenMP team
    dt, _ = run(min(probe, w["N"]))
    rate = min(probe, w["N"]) / max(dt, 1e-6)
    sample = int(min(w["N"], max(probe, rate * seconds)))
    dt, o = run(sample)
    return sample / dt, sample, threads, float((o["naccept"] + o["nreject"]).mean())


def cpu_baseline_entry(pl, w, seconds):
    v, sample, threads, spt = cpu_oracle_rate(pl, w, seconds)
    return {"value": v, "unit": "trajectories/s", "cores": threads, "kind": "port",
            "sample": "%d of the %d trajectories (same inputs/options), CPU oracle with OpenMP schedule(dynamic,64)" % (sample, w["N"]),
            "steps_per_trajectory": spt}


# ---------------------------------------------------------------------------------------------------------------------
class DeviceRun:
    """One workload resident on one GPU: program + device buffers, timed with CUDA events on torch's current stream
    (the stream b200ode_solve_device is launched on)."""

    def __init__(self, pkg, h, w, N, dev, idx=None, offset=0, total=None, extra_options=None):
        import torch
        self.pkg, self.h, self.w, self.N, self.dev, self.torch = pkg, h, w, N, dev, torch
        pl, ll = pkg.problems_library, pkg.lowlevel
        rhs, jac, tg, self.n, self.np_, extra = sources(pl, w, device=True)
        self.dtype = pkg.F32 if w["f32"] else pkg.F64
        if extra_options:
            extra = (extra + " " + extra_options) if extra else extra_options
        self.prog = h.compile(getattr(pkg, ALG_NAMES[w["alg"]]), self.dtype, self.n, self.np_, rhs[0], rhs[1],
                              jac[0] if jac else None, jac[1] if jac else None, tg[0] if tg else None, tg[1] if tg else None,
                              extra_options=extra)
        self.grid = pkg.ranges.saveat_grid(w["saveat"], w["tspan"]) if w["saveat"] is not None else None
        self.nslots = ll.nslots_for(w["tspan"], self.grid) if self.grid else 0
        self.u0, self.p = inputs(pl, w, N, offset=offset, idx=idx, total=total)
        self.rs = 4 if w["f32"] else 8
        self.bufs = ll.DeviceBuffers(self.prog, N, self.nslots, dev, u0_shared=(self.u0.ndim == 1), p_shared=(self.p is None))
        self.bufs.u0.copy_(torch.from_numpy(np.ascontiguousarray(self.u0, dtype=np.float32 if w["f32"] else np.float64)))
        if self.p is not None:
            self.bufs.p.copy_(torch.from_numpy(np.ascontiguousarray(self.p)))
        self.kw = dict(saveat=self.grid, **w["tol"])

    def step(self, first=0, count=None, peer_out=None):
        self.pkg.lowlevel.solve_device(self.prog, self.bufs, self.w["tspan"], first=first, count=count, peer_out=peer_out, **self.kw)

    def timed(self, steps, warmup, barrier):
        torch = self.torch
        for _ in range(warmup):
            self.step()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_wall = time.perf_counter()
        e0.record()
        for i in range(steps):
            ev[i][0].record()
            self.step()
            ev[i][1].record()
        e1.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
        return e0.elapsed_time(e1), [a.elapsed_time(b) for a, b in ev], t_wall

    def attempts(self):
        b = self.bufs
        assert bool((b.retcode == 1).all()), "bench: some trajectories did not reach tf"
        return int((b.naccept.to(self.torch.int64).sum() + b.nreject.to(self.torch.int64).sum()).item())

    def roofline(self, kern_ms, attempts, traffic=None):
        w, N = self.w, self.N
        key = (w["problem"], w["alg"])
        f = FLOPS[key]
        n_interp = N * (len(self.grid) - 1) if self.grid else 0      # the row at tf is a copy, not an interpolation
        flops = f["step"] * attempts + f["row"] * n_interp + F_INIT * N
        peak_tf, _ = self.h.measure_fma_peak(self.dtype)
        achieved = flops / (kern_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        n, rs = self.n, self.rs
        out_bytes = N * self.nslots * n * rs + N * (n * rs + rs + 4 * 8)
        in_bytes = N * self.np_ * rs + N * rs + (0 if self.u0.ndim == 1 else N * n * rs)
        return {"bound": "fp32_fma" if w["f32"] else "fp64_fma", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf,
                "bound_note": "compute-bound on the scalar FP64 (FP32) FMA pipe: the path has no contraction, so neither the "
                              "tensor-core nor the HBM roofline applies; the HBM side of the same launch is under roofline.hbm",
                "peak_source": "measured live: b200ode_measure_fma_peak (register-resident FMA chains); "
                               "MEASURED_PEAKS.json has no FP64/FP32 FMA figure",
                "algorithmic_flops_per_launch": flops, "flops_model": FLOPS_NOTE[key] + " (DESIGN.md §4)",
                "kernel_ms": kern_ms, "traffic": traffic,
                "traffic_source": None if traffic is None else "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch "
                                  "(profiles/r2_ncu_traffic.json: a static capture, not re-measured in this run)",
                "hbm": {"algorithmic_bytes_per_launch": out_bytes + in_bytes,
                        "achieved": (out_bytes + in_bytes) / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                        "unit": "GB/s", "frac": (out_bytes + in_bytes) / (kern_ms * 1e-3) / 1e9 / hbm_peak}}

    def close(self):
        self.bufs = None
        self.prog.close()


def ncu_traffic(workload):
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", name)))
            t = prof.get(workload, {}).get("dram_bytes_per_launch")
            if t is not None:
                return t
        except Exception:
            pass
    return None


def pinned(torch, shape, dt):
    return torch.empty(shape, dtype=dt, pin_memory=True).numpy()


def e2e_host(pkg, torch, run, steps, warmup, barrier, world, dev, dist, extra_kw=None, meanvar=False):
    """The same workload through b200ode_solve with HOST (pinned) buffers: H2D, kernels and D2H inside the timed region."""
    ll = pkg.lowlevel
    w, N, n = run.w, run.N, run.n
    prog = run.prog
    rdt = torch.float32 if w["f32"] else torch.float64
    nsave = prog.nsave
    out = {"u_final": pinned(torch, (N, n), rdt), "t_final": pinned(torch, (N,), torch.float64)}
    if run.nslots > 0:
        if not meanvar:
            out["us"] = pinned(torch, (N, run.nslots, nsave), rdt)
        out["ts"] = np.empty((run.nslots,), dtype=np.float64)
    for k in ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode"):
        out[k] = pinned(torch, (N,), torch.int32)
    u0_h = np.ascontiguousarray(run.u0, dtype=np.float32 if w["f32"] else np.float64)
    p_pin = None
    if run.p is not None:
        p_pin = pinned(torch, run.p.shape, rdt)
        p_pin[...] = run.p
    kw = dict(run.kw, **(extra_kw or {}))
    if meanvar:
        kw["_meanvar"] = (True, True)

    def step_host():
        return ll.solve_host(prog, u0_h, p_pin, w["tspan"], trajectories=N, out=out, **kw)
    for _ in range(max(1, min(warmup, 2))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        r = step_host()
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    h2d = u0_h.nbytes + (p_pin.nbytes if p_pin is not None else 0) + (len(run.grid) * run.rs if run.grid else 0)
    d2h = sum(v.nbytes for k, v in out.items() if isinstance(v, np.ndarray) and k != "ts")
    return {"value": world * N * steps / float(tt.item()), "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "device_ms_per_step": r["total_ms"],
            "kernel_ms_per_step": r["kernel_ms"]}, out


def d2h_ceiling(torch, run, out, barrier, world, dev, dist, reps=3):
    """cudaMemcpyAsync-only baseline: the D2H of the step's result bytes (device buffers -> the same pinned host
    buffers), all ranks at once — what the host side of PCIe allows for this many GPUs, with no kernels involved."""
    pairs = []
    b = run.bufs
    if run.nslots > 0 and "us" in out:
        pairs.append((torch.from_numpy(out["us"]), b.us))
    pairs.append((torch.from_numpy(out["u_final"]), b.u_final))
    for k in ("nsaved", "naccept", "nreject", "nf", "retcode"):
        pairs.append((torch.from_numpy(out[k]), getattr(b, k)))
    nbytes = sum(d.numel() * d.element_size() for d, _ in pairs)
    for d, s in pairs:
        d.copy_(s.reshape(d.shape), non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for d, s in pairs:
            d.copy_(s.reshape(d.shape), non_blocking=True)
    e1.record()
    barrier()
    tt = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    return {"d2h_only_ms_per_step": ms, "d2h_only_gb_s_per_gpu": nbytes / (ms * 1e-3) / 1e9,
            "d2h_only_gb_s_total": world * nbytes / (ms * 1e-3) / 1e9,
            "note": "cudaMemcpyAsync of the step's result bytes into the same pinned buffers on all ranks at once, no kernels: "
                    "the host-side PCIe ceiling of e2e at this GPU count"}


def gather_stats(torch, dist, world, dev, values):
    """min / max / all of one float per rank."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    if world == 1:
        return [t.tolist()]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def strong_section(args, pkg, h, torch, dist, rank, world, local_rank, dev, barrier):
    """ONE 2^20-trajectory Lorenz/Tsit5 saveat ensemble sharded over the ranks in interleaved blocks of 1024."""
    import importlib
    d = importlib.import_module("ordinarydiffeq_jl_b200.distributed")
    w = WORKLOADS["lorenz_tsit5_saveat_1m"]
    Ntot = w["N"]
    idx = d.shard_indices(Ntot, world, rank)
    run = DeviceRun(pkg, h, w, int(idx.shape[0]), dev, idx=idx)
    steps = max(3, min(args.steps, 10))
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total, step_ms, _ = run.timed(steps, max(3, min(args.warmup, 3)), barrier)
    clocks = sampler.finish()
    per_rank = gather_stats(torch, dist, world, dev, [np.mean(step_ms), run.attempts()])
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    kms = [x[0] for x in per_rank]
    out = {"metric": "trajectories/sec", "value": Ntot * steps / (ms * 1e-3), "unit": "trajectories/s", "scaling": "strong",
           "trajectories_total": Ntot, "trajectories_per_gpu": int(idx.shape[0]), "steps": steps, "ms_per_step": ms / steps,
           "partition": "interleaved blocks of 1024 trajectories (block b -> rank b % N); outputs stay sharded in HBM, no collective",
           "kernel_ms_per_rank": {"min": min(kms), "max": max(kms), "all": kms},
           "attempted_steps_per_rank": [int(x[1]) for x in per_rank],
           "load_imbalance": (max(kms) - min(kms)) / max(kms) if max(kms) > 0 else 0.0, "clocks": clocks}
    run.close()
    return out


def sweep_section(args, pkg, h, torch, dist, rank, world, local_rank, dev, barrier, N=None):
    """configs[4]: 64 Mi-trajectory rho sweep, interleaved shards; ordered gather of the final states (copy-free: one
    all_gather per round, written in place) + ensemble mean (device partial sums, all-reduce) inside the timed step."""
    import importlib
    d = importlib.import_module("ordinarydiffeq_jl_b200.distributed")
    ll = pkg.lowlevel
    w = WORKLOADS["lorenz_sweep_64m"]
    N = N or w["N"]
    block = d.sweep_block(N, world)
    idx = d.shard_indices(N, world, rank, block=block)
    m = int(idx.shape[0])
    run = DeviceRun(pkg, h, w, m, dev, idx=idx, total=N)
    part = torch.zeros(3, dtype=torch.float64, device=dev)
    full = torch.empty((N, 3), dtype=torch.float64, device=dev) if world > 1 else None
    in_place = world > 1 and (N % (world * block) == 0)
    steps = max(2, min(args.steps, 4))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    mean = None

    # Pipelined exchange: the rank's shard is integrated in `groups` launches of whole rounds; the all-gathers of a
    # finished group run on a second stream while the next group integrates (the gather writes disjoint slices of
    # `full`, the kernels read and write disjoint rows of the shard), so only the last group's exchange is exposed.
    rounds = (N // (world * block)) if in_place else 0
    groups = 4 if (in_place and rounds >= 8 and rounds % 4 == 0) else 1
    peer = None
    mode = os.environ.get("B200_SWEEP_GATHER", "fused")      # fused | copy | nccl
    if groups > 1 and mode in ("fused", "copy"):
        # push over NVLink peer memory with the copy engines (distributed.PeerGather); NCCL when symmetric memory is not
        # available (all ranks agree on that before anyone enters the collective rendezvous)
        peer, why = d.PeerGather.create(N, (3,), torch.float64, dev, block)
        if peer is not None:
            full = peer.full
        elif rank == 0:
            sys.stderr.write("bench: PeerGather unavailable (%s); NCCL all-gather after the integration\n" % why)
    if peer is None:
        groups = 1
    # "fused": the integration kernel itself stores every final state at its global index into every rank's result
    # (peer stores over NVLink, B200DeviceResult.peer_u_final) — one launch, no copies; "copy": the copy-engine push above
    fused = peer is not None and mode == "fused"
    peer_ptrs = ([t.data_ptr() for t in peer.peers], world, rank, block) if fused else None
    gev = [torch.cuda.Event() for _ in range(groups)]

    def step(i=None):
        nonlocal mean
        if i is not None:
            ev[i][0].record()
        if fused:
            run.step(peer_out=peer_ptrs)
            ll.reduce_sum_device(h, pkg.F64, run.bufs.u_final, pkg._lib.LAYOUT_AOS, m, 3, part)
            if i is not None:
                ev[i][1].record()
            mean = d.allreduce_mean(part, N)        # also the barrier after which every rank's `full` is complete
        elif peer is not None:
            cur = torch.cuda.current_stream()
            rpg = rounds // groups
            peer.begin(cur)
            for g in range(groups):
                run.step(first=g * rpg * block, count=rpg * block)
                gev[g].record(cur)
                peer.push(run.bufs.u_final, g * rpg, (g + 1) * rpg, gev[g])
            ll.reduce_sum_device(h, pkg.F64, run.bufs.u_final, pkg._lib.LAYOUT_AOS, m, 3, part)
            if i is not None:
                ev[i][1].record()
            peer.finish(cur)
            mean = d.allreduce_mean(part, N)        # also the barrier after which every rank's `full` is complete
        else:
            run.step()
            ll.reduce_sum_device(h, pkg.F64, run.bufs.u_final, pkg._lib.LAYOUT_AOS, m, 3, part)
            if i is not None:
                ev[i][1].record()
            if world > 1:
                if in_place:
                    d.gather_in_place(run.bufs.u_final, N, block, out=full)
                else:
                    full.copy_(d.gather_in_order(run.bufs.u_final, N, block=block))
                mean = d.allreduce_mean(part, N)
            else:
                mean = part / float(N)
        if i is not None:
            ev[i][2].record()
    for _ in range(3):
        step()
    barrier()
    if peer is not None:        # the pushed result equals NCCL's ordered all-gather, bit for bit (checked outside the timed region)
        torch.cuda.synchronize()
        ref = d.gather_in_place(run.bufs.u_final, N, block)
        torch.cuda.synchronize()
        assert torch.equal(ref, full), "PeerGather result differs from the NCCL all-gather"
        del ref
        barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    barrier()
    clocks = sampler.finish()
    kern = float(np.mean([a[0].elapsed_time(a[1]) for a in ev]))
    coll = float(np.mean([a[1].elapsed_time(a[2]) for a in ev]))
    per_rank = gather_stats(torch, dist, world, dev, [kern, coll])
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    attempts = run.attempts()
    if world > 1:
        # the gathered array is in global trajectory order: this rank's block 0 sits at global block `rank`
        assert torch.equal(full[rank * block:(rank + 1) * block], run.bufs.u_final[:block])
    kms = [x[0] for x in per_rank]
    cms = [x[1] for x in per_rank]
    out = {"metric": "trajectories/sec", "value": N * steps / (ms * 1e-3), "unit": "trajectories/s", "scaling": "strong",
           "trajectories_total": N, "trajectories_per_gpu": m, "steps": steps, "ms_per_step": ms / steps,
           "partition": "interleaved blocks of %d trajectories (block b -> rank b %% N)" % block,
           "collectives": "none (single GPU)" if world == 1 else
                          ("fused into the integration kernel: every finished trajectory's final state is stored at its global index "
                           "into every rank's result (symmetric memory, peer stores over NVLink, B200DeviceResult.peer_u_final): no gather "
                           "step, no copy; equal to the NCCL all-gather bit for bit (asserted before timing); + ncclAllReduce of the partial "
                           "sums, which also closes the step across ranks") if fused else
                          ("push over NVLink peer memory: every rank copies each finished block of its shard into its final place "
                           "in every rank's result (symmetric memory, device-to-peer cudaMemcpyAsync on the copy engines, %d blocks x "
                           "%d peers per step) while the next of %d groups of rounds integrates; equal to the NCCL all-gather bit for bit "
                           "(asserted before timing); + ncclAllReduce of the partial sums, which also closes the step across ranks. "
                           "collective_ms = what remains exposed after the last kernel" % (rounds, world, groups)) if peer is not None else
                          "%d x ncclAllGather of the final states written in place (no un-interleave copy) + ncclAllReduce of the "
                          "partial sums, over NVLink" % (N // (world * block) if in_place else 1),
           "kernel_ms_per_rank": {"min": min(kms), "max": max(kms), "all": kms},
           "collective_ms_per_rank": {"min": min(cms), "max": max(cms), "all": cms},
           "collective_share": (max(cms) / (ms / steps)) if ms > 0 else None,
           "gathered_bytes": N * 24, "load_imbalance": (max(kms) - min(kms)) / max(kms) if max(kms) > 0 else 0.0,
           "ensemble_mean_u_tf": [float(x) for x in mean.cpu()], "attempted_steps_this_rank": attempts, "clocks": clocks,
           "roofline": run.roofline(kern, attempts)}
    run.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="lorenz_tsit5_saveat_1m")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-scaling-sections", action="store_true", help="skip strong_1m / sweep_64m")
    ap.add_argument("--sweep-n", type=int, default=0, help="trajectories of the sweep section (default 2^26)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "trajectories/sec"
    config = {"workload": args.workload, "problem": w["problem"], "alg": w["alg"], "trajectories_per_gpu": w["N"],
              "saveat": w["saveat"], "tspan": list(w["tspan"]), "tolerances": w["tol"] or "defaults (reltol 1e-3, abstol 1e-6)",
              "partition": "independent trajectories per GPU, no data-path collective",
              "cache": "per-step working set (inputs + saveat output, GBs) >> 126 MB L2; no flush needed"}

    import b200_import
    pkg = b200_import.load()
    pl = pkg.problems_library

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return 0
        vals = []
        sample = threads = None
        per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(args.steps + args.warmup, 1)))
        for i in range(args.warmup + args.steps):
            r, sample, threads, _ = cpu_oracle_rate(pl, w, per_step)
            if i >= args.warmup:
                vals.append(r)
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": metric, "value": v, "unit": "trajectories/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if w["f32"] else "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "trajectories/s", "cores": threads, "kind": "port",
                                 "sample": "%d trajectories of the same workload per step (CPU oracle, OpenMP, "
                                           "all host threads; the reference is Julia and cannot run here)" % sample},
                "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU oracle")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")       # keep NCCL's version banner off stdout (ONE JSON line)
        dist.init_process_group("nccl", device_id=dev)
    h = pkg.Handle(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if w["problem"] == "lorenz_sweep":
        sec = sweep_section(args, pkg, h, torch, dist, rank, world, local_rank, dev, barrier, N=args.sweep_n or None)
        if rank == 0:
            line = dict(sec, n_gpus=world, warmup=3, higher_is_better=True, vs_baseline=None, dtype="f64", data="synthetic",
                        config=dict(config, trajectories_total=sec["trajectories_total"]), gpu_launches=4 * sec["steps"])
            print(json.dumps(line))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    N = w["N"]
    run = DeviceRun(pkg, h, w, N, dev, offset=rank * N)       # weak scaling: every rank integrates its own N trajectories
    for _ in range(args.warmup):
        run.step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_ms_total, step_ms, t_wall = run.timed(args.steps, 0, barrier)
    clocks = sampler.finish()
    attempts = run.attempts()
    t = torch.tensor([dev_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * N * args.steps / (ms_total * 1e-3)
    kern_ms = float(np.mean(step_ms))
    per_rank = gather_stats(torch, dist, world, dev, [kern_ms])
    roofline = run.roofline(kern_ms, attempts, traffic=ncu_traffic(args.workload)) if (w["problem"], w["alg"]) in FLOPS else None

    # ---- e2e: host (pinned) buffers through b200ode_solve, H2D and D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        k_e2e = max(2, min(args.steps, 5))
        e2e, out = e2e_host(pkg, torch, run, k_e2e, args.warmup, barrier, world, dev, dist)
        e2e["api"] = "b200ode_solve (C ABI, pinned host buffers)"
        e2e["copy_ceiling"] = d2h_ceiling(torch, run, out, barrier, world, dev, dist)
        ceil_ms = e2e["copy_ceiling"]["d2h_only_ms_per_step"]
        e2e["fraction_of_copy_ceiling"] = (ceil_ms / e2e["device_ms_per_step"]) if e2e["device_ms_per_step"] else None
        del out
        variants = {}
        if run.nslots > 0:
            # only EnsembleAnalysis statistics wanted: rows stay in HBM (b200ode_solve_meanvar)
            v, _ = e2e_host(pkg, torch, run, k_e2e, 1, barrier, world, dev, dist, meanvar=True)
            v["api"] = "b200ode_solve_meanvar (timeseries mean/var reduced on the device)"
            variants["meanvar_on_device"] = v
        e2e["variants"] = variants
        e2e["stats_only"] = variants.get("meanvar_on_device")

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_entry(pl, w, args.cpu_seconds)
    prog_info = run.prog.info
    run.close()

    # ---- the other BASELINE configurations, short runs (N = 1 only: they are single-GPU configurations)
    configs = []
    if world == 1 and not args.no_configs and args.workload == "lorenz_tsit5_saveat_1m":
        for name in OTHER_CONFIGS:
            wc = WORKLOADS[name]
            r = DeviceRun(pkg, h, wc, wc["N"], dev)
            smp = ClockSampler(local_rank)
            smp.start()
            tot, sms, _ = r.timed(3, 3, barrier)
            ck = smp.finish()
            att = r.attempts()
            entry = {"workload": name, "problem": wc["problem"], "alg": wc["alg"], "dtype": "f32" if wc["f32"] else "f64",
                     "trajectories": wc["N"], "saveat": wc["saveat"], "tolerances": wc["tol"] or "defaults",
                     "value": wc["N"] * 3 / (tot * 1e-3), "unit": "trajectories/s", "ms_per_step": tot / 3, "steps": 3, "warmup": 3,
                     "attempted_steps_per_launch": att, "clocks": ck, "program": r.prog.info,
                     "roofline": r.roofline(float(np.mean(sms)), att, traffic=ncu_traffic(name))
                     if (wc["problem"], wc["alg"]) in FLOPS else None}
            if wc["alg"] == "autotsit5":
                b = r.bufs
                stiff_att = int(b.nw.to(torch.int64).sum().item())
                entry["mix"] = {"attempts_rosenbrock23": stiff_att, "attempts_tsit5": att - stiff_att,
                                "trajectories_that_used_rosenbrock23": int((b.nw > 0).sum().item())}
            if name == "lorenz_tsit5_saveat_1m_f32" and not args.no_e2e:
                v, _ = e2e_host(pkg, torch, r, 3, 1, barrier, world, dev, dist)
                v["api"] = "b200ode_solve (C ABI, pinned host buffers), FP32"
                entry["e2e"] = v
            if not args.no_cpu_baseline:
                entry["cpu_baseline"] = cpu_baseline_entry(pl, wc, 3.0)
            r.close()
            configs.append(entry)
        if not args.no_e2e:
            # save_idxs variant of the headline workload: one component per saved row (a third of the row bytes)
            r = DeviceRun(pkg, h, w, N, dev, extra_options=pkg._lib.opt_save_idxs([0]))
            v, _ = e2e_host(pkg, torch, r, 3, 1, barrier, world, dev, dist)
            v["api"] = "b200ode_solve (C ABI, pinned host buffers), save_idxs = [1]"
            e2e["variants"]["save_idxs_first_component"] = v
            r.close()

    strong = sweep = None
    if not args.no_scaling_sections and args.workload == "lorenz_tsit5_saveat_1m":
        strong = strong_section(args, pkg, h, torch, dist, rank, world, local_rank, dev, barrier)
        sweep = sweep_section(args, pkg, h, torch, dist, rank, world, local_rank, dev, barrier, N=args.sweep_n or None)

    if rank == 0:
        launches = 2 * args.steps
        line = {"metric": metric, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if w["f32"] else "f64", "data": "synthetic", "config": config,
                "clocks": clocks, "gpu_launches": launches,
                "gpu_launches_note": "b200_initdt + b200_integrate per step of the timed headline region (NVRTC-compiled, "
                                     "loaded by libb200ode.so); the other sections launch the same two kernels per step",
                "attempted_steps_per_launch": attempts, "wall_s_timed_region": t_wall,
                "kernel_ms_per_rank": {"min": min(x[0] for x in per_rank), "max": max(x[0] for x in per_rank)},
                "program": prog_info}
        if e2e:
            line["e2e"] = e2e
        if roofline:
            line["roofline"] = roofline
        if cpu:
            line["cpu_baseline"] = cpu
        if configs:
            line["configs"] = configs
        if strong:
            line["strong_1m"] = strong
        if sweep:
            line["sweep_64m"] = sweep
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
