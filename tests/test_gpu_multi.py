"""Multi-GPU parity (-m gpu; skipped on a single-GPU box): the interleaved shards of the rho sweep, integrated by one
process per GPU and gathered back over NCCL, must equal the single-GPU run bit for bit and in trajectory order;
the all-reduced ensemble mean must equal the single-GPU mean to 1e-12 (different summation tree)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    import importlib
    import torch
    import torch.distributed as dist
    import b200_import
    pkg = b200_import.load()
    d = importlib.import_module("ordinarydiffeq_jl_b200.distributed")
    pl, ll = pkg.problems_library, pkg.lowlevel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    h = pkg.Handle(rank)
    rhs = pl.lorenz_source(False)
    prog = h.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
    idx = d.shard_indices(N, world, rank)
    m = int(idx.shape[0])
    p = pl.lorenz_params(N, sweep_total=N)[idx]
    bufs = ll.DeviceBuffers(prog, m, 0, dev, u0_shared=True)
    bufs.u0.copy_(torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64))
    bufs.p.copy_(torch.from_numpy(np.ascontiguousarray(p)))
    ll.solve_device(prog, bufs, (0.0, 10.0))
    part = torch.zeros(3, dtype=torch.float64, device=dev)
    ll.reduce_sum_device(h, pkg.F64, bufs.u_final, pkg._lib.LAYOUT_AOS, m, 3, part)
    full = d.gather_in_order(bufs.u_final, N)
    steps = d.gather_in_order((bufs.naccept + bufs.nreject).to(torch.int32), N)
    mean = d.allreduce_mean(part, N)
    torch.cuda.synchronize()
    # the push over NVLink peer memory (PeerGather: sub-range launches + device-to-peer copies on per-peer streams, closed by
    # the all-reduce) must give the same array as NCCL's gather; "unavailable" when symmetric memory cannot be set up
    peer_ok = "n/a"
    if N % (world * 1024) == 0:
        pg, why = d.PeerGather.create(N, (3,), torch.float64, dev, 1024)
        if pg is None:
            peer_ok = "unavailable: %s" % why
        if pg is not None:
            cur = torch.cuda.current_stream()
            bufs.u_final.zero_()
            rounds, groups = pg.rounds, 2 if pg.rounds % 2 == 0 else 1
            rpg = rounds // groups
            pg.begin(cur)
            evs = [torch.cuda.Event() for _ in range(groups)]
            for g in range(groups):
                ll.solve_device(prog, bufs, (0.0, 10.0), first=g * rpg * 1024, count=rpg * 1024)
                evs[g].record(cur)
                pg.push(bufs.u_final, g * rpg, (g + 1) * rpg, evs[g])
            pg.finish(cur)
            d.allreduce_mean(part, N)
            torch.cuda.synchronize()
            peer_ok = bool(torch.equal(pg.full, full))
            # ... and the gather fused into the kernel (B200DeviceResult.peer_u_final): one launch whose trajectory ends store
            # into every rank's result at the global index
            pg.full.zero_()
            dist.barrier()
            ll.solve_device(prog, bufs, (0.0, 10.0), peer_out=([t.data_ptr() for t in pg.peers], world, rank, 1024))
            d.allreduce_mean(part, N)
            torch.cuda.synchronize()
            peer_ok = peer_ok and bool(torch.equal(pg.full, full))
    if rank == 0:
        q.put((full.cpu().numpy(), steps.cpu().numpy(), mean.cpu().numpy(), peer_ok))
    dist.barrier()
    dist.destroy_process_group()
    prog.close()
    h.close()


@pytest.mark.parametrize("N", [8192, 5000])      # a multiple of world*1024 (strided un-interleave) and a ragged count
def test_two_rank_nccl_sweep_equals_single_gpu(pkg, handle, N):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    pl, ll = pkg.problems_library, pkg.lowlevel
    rhs = pl.lorenz_source(False)
    prog = handle.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
    one = ll.solve_host(prog, np.array([1.0, 0.0, 0.0]), pl.lorenz_params(N, sweep_total=N), (0.0, 10.0))
    prog.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, steps, mean, peer_ok = q.get(timeout=300)
    assert peer_ok in (True, "n/a") or str(peer_ok).startswith("unavailable"), "PeerGather differs from the NCCL gather"
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert np.array_equal(full.view(np.uint64), one["u_final"].view(np.uint64))
    assert np.array_equal(steps, one["naccept"] + one["nreject"])
    assert np.allclose(mean, one["u_final"].mean(axis=0), rtol=1e-12, atol=0)


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_device_c_abi_equals_single_device(pkg, handle, ndev):
    """b200ode_multi_solve / b200ode_multi_reduce_mean (several GPUs from one process): results in global trajectory
    order equal the single-device solve bit for bit — with saveat rows and final-state only — and the mean equals the
    mean of those final states to 1e-12 (chunk-wise summation tree)."""
    import torch
    if torch.cuda.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    pl, ll = pkg.problems_library, pkg.lowlevel
    rhs = pl.lorenz_source(False)
    N = 40000 + 37                      # several chunks per device, ragged tail
    p = pl.lorenz_params(N, sweep_total=N)
    u0 = np.array([1.0, 0.0, 0.0])
    prog1 = handle.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
    mh = pkg.MultiHandle(list(range(ndev)))
    assert mh.ndev == ndev
    progm = mh.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1])
    try:
        grid = [1.0, 2.5, 4.0]
        for kw in ({}, {"saveat": grid}, {"saveat": grid, "save_start": False, "save_end": False}):
            a = ll.solve_host(prog1, u0, p, (0.0, 4.0), **kw)
            b = ll.solve_host(progm, u0, p, (0.0, 4.0), **kw)
            for k in ("naccept", "nreject", "nf", "retcode", "nsaved"):
                assert np.array_equal(a[k], b[k]), k
            assert np.array_equal(a["u_final"].view(np.uint64), b["u_final"].view(np.uint64))
            assert np.array_equal(a["t_final"], b["t_final"])
            if "saveat" in kw:
                assert np.array_equal(a["us"].view(np.uint64), b["us"].view(np.uint64))
                assert np.array_equal(a["ts"], b["ts"])
        a = ll.solve_host(prog1, u0, p, (0.0, 4.0))
        m = ll.solve_host_mean(progm, u0, p, (0.0, 4.0))
        assert np.array_equal(a["u_final"].view(np.uint64), m["u_final"].view(np.uint64))
        assert np.allclose(m["mean"], a["u_final"].mean(axis=0), rtol=1e-12, atol=0)
        # a reverse-time program (tspan[2] < tspan[1], B200ODE_OPT_REVERSE_TIME) through the multi-device entry points
        R = pkg._lib.OPT_REVERSE_TIME
        pr1 = handle.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1], extra_options=R)
        prm = mh.compile(pkg.ALG_TSIT5, pkg.F64, 3, 3, rhs[0], rhs[1], extra_options=R)
        try:
            for kw in ({}, {"saveat": [0.3, 0.2, 0.05, 0.0]}):
                a = ll.solve_host(pr1, u0, p, (0.4, 0.0), **kw)
                b = ll.solve_host(prm, u0, p, (0.4, 0.0), **kw)
                for k in ("naccept", "nreject", "nf", "retcode", "nsaved"):
                    assert np.array_equal(a[k], b[k]), k
                assert np.array_equal(a["u_final"].view(np.uint64), b["u_final"].view(np.uint64))
                assert (b["t_final"] == 0.0).all() and (b["retcode"] == 1).all()
                if "saveat" in kw:
                    assert np.array_equal(a["us"].view(np.uint64), b["us"].view(np.uint64)) and list(b["ts"]) == [0.4, 0.3, 0.2, 0.05, 0.0]
        finally:
            prm.close()
            pr1.close()
    finally:
        progm.close()
        mh.close()
        prog1.close()
