#!/usr/bin/env python
"""bench.py — the ensemble hot path on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one pass of the hot path over one batch: solve a 2^20-trajectory Lorenz/Tsit5 FP64
ensemble with saveat = 0.1 (BASELINE.json configs[1], the configuration `metric` is quoted on).
Prints ONE JSON line (rank 0):
  value      trajectories/s, whole job, inputs resident in HBM, CUDA-event time, max over ranks
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H timed
  roofline   FP64-FMA roofline of b200_integrate (algorithmic flops / CUDA-event time vs the FMA
             peak measured live on this GPU) + the HBM side of the saveat stream
  cpu_baseline  the CPU oracle (port of the reference algorithm) on this box's host cores, bounded sample
`--impl reference` times the CPU oracle alone (the reference is Julia and cannot run here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

F_STEP_TSIT5_LORENZ = 247.0     # algorithmic FP64 flops per attempted step
F_INTERP = 85.0                 # per interpolated saveat row
F_INIT = 60.0                   # initdt: 2 RHS + 3 norms

WORKLOADS = {
    # name: (problem, alg, f32, N per GPU, saveat h or None, tspan, tolerances)
    "lorenz_tsit5_saveat_1m": dict(problem="lorenz", alg="tsit5", f32=False, N=1 << 20, saveat=0.1, tspan=(0.0, 10.0), tol={}),
    "lorenz_tsit5_saveat_1m_f32": dict(problem="lorenz", alg="tsit5", f32=True, N=1 << 20, saveat=0.1, tspan=(0.0, 10.0), tol={}),
    "lorenz_tsit5_final_1m": dict(problem="lorenz", alg="tsit5", f32=False, N=1 << 20, saveat=None, tspan=(0.0, 10.0), tol={}),
    "robertson_rodas5p_1m": dict(problem="robertson", alg="rodas5p", f32=False, N=1 << 20, saveat=None, tspan=(0.0, 1e5),
                                 tol=dict(reltol=1e-6, abstol=1e-8)),
    "robertson_rosenbrock23_1m": dict(problem="robertson", alg="ros23", f32=False, N=1 << 20, saveat=None, tspan=(0.0, 1e5),
                                      tol=dict(reltol=1e-6, abstol=1e-8)),
    # BASELINE.json configs[4]: one 64 Mi-trajectory parameter sweep sharded over the ranks (strong scaling),
    # NCCL all-gather of the final states into trajectory order + ensemble mean (all-reduce of partial sums)
    "lorenz_sweep_64m": dict(problem="lorenz_sweep", alg="tsit5", f32=False, N=1 << 26, saveat=None, tspan=(0.0, 10.0), tol={}),
    "pleiades_vern7_256k": dict(problem="pleiades", alg="vern7", f32=False, N=1 << 18, saveat=None, tspan=(0.0, 3.0),
                                tol=dict(reltol=1e-6, abstol=1e-8)),
}


def sources(pl, w):
    f32 = w["f32"]
    if w["problem"] in ("lorenz", "lorenz_sweep"):
        return pl.lorenz_source(f32), None, None, 3, 3
    if w["problem"] == "robertson":
        r, j, tg = pl.robertson_sources(f32)
        return r, j, tg, 3, 3
    if w["problem"] == "pleiades":
        # loop-form source + partially rolled stage loops: same speed as the fully unrolled form,
        # 20x shorter NVRTC compile
        return pl.pleiades_source(f32, loops=True), None, None, 28, 0
    raise ValueError(w["problem"])


def inputs(pl, w, N, offset):
    f32 = w["f32"]
    if w["problem"] == "lorenz":
        return np.array([1.0, 0.0, 0.0]), pl.lorenz_params(N, offset=offset, f32=f32)
    if w["problem"] == "robertson":
        return np.array([1.0, 0.0, 0.0]), pl.robertson_params(N, offset=offset, f32=f32)
    return pl.pleiades_u0(N, offset=offset, f32=f32), None


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
    rhs, jac, tg, n, np_ = sources(pl, w)
    alg = {"tsit5": oracle.ALG_TSIT5, "vern7": oracle.ALG_VERN7, "ros23": oracle.ALG_ROSENBROCK23,
           "rodas5p": oracle.ALG_RODAS5P}[w["alg"]]
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

    def run(N):
        u0, p = inputs(pl, w, N, 0)
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


def sweep_main(args, w, pkg, rank, world, local_rank, metric, config):
    """configs[4]: 64 Mi-trajectory rho sweep, interleaved shards, gather + mean over NCCL."""
    import importlib
    import torch
    import torch.distributed as dist
    d = importlib.import_module("ordinarydiffeq_jl_b200.distributed")
    pl, ll = pkg.problems_library, pkg.lowlevel
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = w["N"]
    h = pkg.Handle(local_rank)
    rhs = pl.lorenz_source(False)
    prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
    idx = d.shard_indices(N, world, rank)
    m = int(idx.shape[0])
    p = np.empty((m, 3), dtype=np.float64)
    p[:, 0] = 10.0
    p[:, 1] = 14.0 + 28.0 * idx.astype(np.float64) / float(N)
    p[:, 2] = 8.0 / 3.0
    bufs = ll.DeviceBuffers(prog, m, 0, dev, u0_shared=True)
    bufs.u0.copy_(torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64))
    bufs.p.copy_(torch.from_numpy(p))
    part = torch.zeros(3, dtype=torch.float64, device=dev)

    def step():
        ll.solve_device(prog, bufs, w["tspan"])
        ll.reduce_sum_device(h, pkg.F64, bufs.u_final, pkg._lib.LAYOUT_AOS, m, 3, part)
        if world > 1:
            full = d.gather_in_order(bufs.u_final, N)
            mean = d.allreduce_mean(part, N)
        else:
            full, mean = bufs.u_final, part / float(N)
        return full, mean

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        full, mean = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    assert full.shape[0] == N and bool((bufs.retcode == 1).all())
    if rank == 0:
        print(json.dumps({"metric": metric, "value": N * args.steps / (ms * 1e-3), "unit": "trajectories/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": dict(config, trajectories_total=N, trajectories_per_gpu=m,
                                         partition="interleaved blocks of 1024 trajectories; NCCL all-gather of final states + all-reduce mean"),
                          "ensemble_mean_u_tf": [float(x) for x in mean.cpu()], "gpu_launches": 4 * args.steps}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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

    if w["problem"] == "lorenz_sweep":
        return sweep_main(args, w, pkg, rank, world, local_rank, metric, config)

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU oracle")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ll = pkg.lowlevel
    h = pkg.Handle(local_rank)
    rhs, jac, tg, n, np_ = sources(pl, w)
    alg_id = {"tsit5": pkg.ALG_TSIT5, "vern7": pkg.ALG_VERN7, "ros23": pkg.ALG_ROSENBROCK23, "rodas5p": pkg.ALG_RODAS5P}[w["alg"]]
    dtype = pkg.F32 if w["f32"] else pkg.F64
    prog = h.compile(alg_id, dtype, n, np_, rhs[0], rhs[1], jac[0] if jac else None, jac[1] if jac else None,
                     tg[0] if tg else None, tg[1] if tg else None,
                     extra_options="-DB200_STAGE_UNROLL=4" if w["problem"] == "pleiades" else None)
    N = w["N"]
    grid = pkg.ranges.saveat_grid(w["saveat"], w["tspan"]) if w["saveat"] is not None else None
    nslots = ll.nslots_for(w["tspan"], grid) if grid else 0
    u0, p = inputs(pl, w, N, rank * N)        # weak scaling: every rank integrates its own N trajectories
    rs = 4 if w["f32"] else 8
    dev = torch.device("cuda", local_rank)

    # device-resident buffers (value)
    bufs = ll.DeviceBuffers(prog, N, nslots, dev, u0_shared=(u0.ndim == 1), p_shared=(p is None))
    bufs.u0.copy_(torch.from_numpy(np.ascontiguousarray(u0, dtype=np.float32 if w["f32"] else np.float64)))
    if p is not None:
        bufs.p.copy_(torch.from_numpy(np.ascontiguousarray(p)))
    kw = dict(saveat=grid, **w["tol"])

    def step_device():
        ll.solve_device(prog, bufs, w["tspan"], **kw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall = time.perf_counter()
    e_all0.record()
    for i in range(args.steps):
        ev[i][0].record()
        step_device()
        ev[i][1].record()
    e_all1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms_total = e_all0.elapsed_time(e_all1)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # statistics of the run (for the algorithmic flop count) and a sanity check that it solved
    naccept = bufs.naccept.cpu().numpy().astype(np.int64)
    nreject = bufs.nreject.cpu().numpy().astype(np.int64)
    retcode = bufs.retcode.cpu().numpy()
    assert (retcode == 1).all(), "bench: some trajectories did not reach tf"
    attempts = int(naccept.sum() + nreject.sum())
    t = torch.tensor([dev_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * N * args.steps / (ms_total * 1e-3)

    # ---- kernel-only duration of b200_integrate for the roofline: one extra pass with events around the solve call
    # (initdt + integrate; initdt is <2% of the time, see profiles/)
    kern_ms = float(np.mean(step_ms))
    flops_per_launch = None
    roofline = None
    if w["problem"] == "lorenz" and w["alg"] == "tsit5":
        n_interp = N * (len(grid) - 1) if grid else 0            # the row at tf is a copy, not an interpolation
        flops_per_launch = F_STEP_TSIT5_LORENZ * attempts + F_INTERP * n_interp + F_INIT * N
        peak_tf, _ = h.measure_fma_peak(dtype)
        achieved = flops_per_launch / (kern_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        out_bytes = N * nslots * n * rs + N * (n * rs + rs + 4 * 8)
        in_bytes = N * np_ * rs + N * rs
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
            traffic = prof.get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "fp64_fma" if not w["f32"] else "fp32_fma", "achieved": achieved, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "bound_note": "compute-bound on the scalar FP64 (FP32) FMA pipe: the path has no contraction, so "
                                  "neither the tensor-core nor the HBM roofline applies; the HBM side of the same "
                                  "launch is reported under roofline.hbm",
                    "peak_source": "measured live: b200ode_measure_fma_peak (register-resident FMA chains); "
                                   "MEASURED_PEAKS.json has no FP64 figure",
                    "algorithmic_flops_per_launch": flops_per_launch,
                    "flops_model": "247/attempted step + 85/interpolated row + 60/trajectory (DESIGN.md)",
                    "kernel_ms": kern_ms, "traffic": traffic,
                    "hbm": {"algorithmic_bytes_per_launch": out_bytes + in_bytes,
                            "achieved": (out_bytes + in_bytes) / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                            "unit": "GB/s", "frac": (out_bytes + in_bytes) / (kern_ms * 1e-3) / 1e9 / hbm_peak}}

    # ---- e2e: host (pinned) buffers through b200ode_solve, H2D and D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        def pinned(shape, dt):
            tt = torch.empty(shape, dtype=dt, pin_memory=True)
            return tt.numpy()
        rdt = torch.float32 if w["f32"] else torch.float64
        out = {"u_final": pinned((N, n), rdt), "t_final": pinned((N,), torch.float64)}
        if nslots > 0:
            out["us"] = pinned((N, nslots, n), rdt)
            out["ts"] = np.empty((nslots,), dtype=np.float64)
        for k in ("nsaved", "naccept", "nreject", "nf", "njacs", "nw", "nsolve", "retcode"):
            out[k] = pinned((N,), torch.int32)
        u0_h = np.ascontiguousarray(u0, dtype=np.float32 if w["f32"] else np.float64)
        if p is not None:
            p_pin = pinned(p.shape, rdt)
            p_pin[...] = p
        else:
            p_pin = None

        def step_host():
            return ll.solve_host(prog, u0_h, p_pin, w["tspan"], trajectories=N, out=out, **kw)
        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        k_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            r = step_host()
        barrier()
        dt_e2e = time.perf_counter() - t0
        tt = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d = u0_h.nbytes + (p_pin.nbytes if p_pin is not None else 0) + (len(grid) * rs if grid else 0)
        d2h = sum(v.nbytes for k, v in out.items() if isinstance(v, np.ndarray) and k != "ts")
        e2e = {"value": world * N * k_e2e / float(tt.item()), "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": k_e2e, "api": "b200ode_solve (C ABI, pinned host buffers)",
               "device_ms_per_step": r["total_ms"], "kernel_ms_per_step": r["kernel_ms"]}

        # informational: the same ensemble when only EnsembleAnalysis statistics are wanted
        # (b200ode_solve_meanvar: rows stay in HBM, mean/var per saved row come back)
        if nslots > 0:
            out2 = {k: v for k, v in out.items() if k not in ("us",)}
            ll.solve_host(prog, u0_h, p_pin, w["tspan"], trajectories=N, out=out2, _meanvar=(True, True), **kw)
            barrier()
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                ll.solve_host(prog, u0_h, p_pin, w["tspan"], trajectories=N, out=out2, _meanvar=(True, True), **kw)
            barrier()
            e2e["stats_only"] = {"value": world * N * k_e2e / (time.perf_counter() - t0), "unit": "trajectories/s",
                                 "api": "b200ode_solve_meanvar (timeseries mean/var reduced on the device)",
                                 "d2h_bytes_per_step": int(sum(v.nbytes for k, v in out2.items() if isinstance(v, np.ndarray) and k != "ts"))}

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sample, threads, spt = cpu_oracle_rate(pl, w, args.cpu_seconds)
        cpu = {"value": v, "unit": "trajectories/s", "cores": threads, "kind": "port",
               "sample": "%d of the %d trajectories (same inputs/options), CPU oracle with OpenMP schedule(dynamic,64)" % (sample, N),
               "steps_per_trajectory": spt}

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if w["f32"] else "f64", "data": "synthetic", "config": config,
                "clocks": sampler.summary(), "gpu_launches": 2 * args.steps,
                "gpu_launches_note": "b200_initdt + b200_integrate per step (NVRTC-compiled, loaded by libb200ode.so)",
                "attempted_steps_per_launch": attempts, "wall_s_timed_region": t_wall,
                "program": prog.info}
        if e2e:
            line["e2e"] = e2e
        if roofline:
            line["roofline"] = roofline
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
