"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink).

Trajectories are independent, so integration needs no communication at all.
The ensemble is dealt to the ranks in interleaved blocks — a contiguous split would hand each GPU
a different band of the swept parameter and therefore a different amount of adaptive-step work.
Only the optional end-of-solve exchange uses a collective:
  * `gather_in_order`: all-gather of the final states back into global trajectory order,
  * `allreduce_mean`: ensemble mean from per-GPU partial sums (the device-side `reduction`).
The reference's analogue is EnsembleDistributed's pmap over worker processes (SciMLBase, EXT;
exercised at /root/reference/lib/DiffEqBase/test/downstream/distributed_ensemble.jl:43-51).
"""
import numpy as np

BLOCK = 1024


def shard_indices(N, world, rank, block=BLOCK):
    """Global trajectory indices (0-based, ascending) owned by `rank`: blocks b with b % world == rank."""
    nblocks = (N + block - 1) // block
    mine = np.arange(rank, nblocks, world, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).ravel()
    return idx[idx < N]


def shard_sizes(N, world, block=BLOCK):
    return [int(shard_indices(N, world, r, block).shape[0]) for r in range(world)]


_idx_cache = {}


def gather_in_order(local, N, block=BLOCK, group=None):
    """local: tensor [m_rank, ...] holding this rank's trajectories in shard order.
    Returns the tensor [N, ...] in global trajectory order on every rank (one all_gather; the
    un-interleave is a strided device copy when N is a multiple of world*block)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    tail = tuple(local.shape[1:])
    if N % (world * block) == 0:
        nb = N // (world * block)
        out = torch.empty((world,) + (nb, block) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view((world * nb * block,) + tail), local.contiguous(), group=group)
        # out[r, b] is global block b*world + r
        return out.permute(1, 0, 2, *range(3, out.dim())).reshape((N,) + tail)
    sizes = shard_sizes(N, world, block)
    mmax = max(sizes)
    pad = torch.zeros((mmax,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    full = torch.empty((N,) + tail, dtype=local.dtype, device=local.device)
    for r in range(world):
        key = (N, world, r, block, str(local.device))
        if key not in _idx_cache:
            _idx_cache[key] = torch.from_numpy(shard_indices(N, world, r, block)).to(local.device)
        full[_idx_cache[key]] = out[r * mmax: r * mmax + sizes[r]]
    return full


def sweep_block(N, world, chunks_per_rank=32, base=BLOCK):
    """Interleave granularity for large ensembles: ~chunks_per_rank blocks per rank (whole multiples of `base`),
    coarse enough that the ordered gather is a handful of collectives, fine enough that a parameter sweep whose cost
    varies smoothly with the index is balanced to within a block."""
    b = max(base, N // max(1, world * chunks_per_rank))
    return max(base, (b // base) * base)


def gather_in_place(local, N, block, out=None, group=None):
    """Ordered all-gather WITHOUT an un-interleave copy: with blocks dealt round-robin (block b -> rank b % world),
    round k of the gather (every rank's k-th block) is exactly the contiguous slice [k*world*block, (k+1)*world*block)
    of the result in rank order — one all_gather_into_tensor per round, written straight into its final place.
    Needs N % (world*block) == 0 (callers pick `block` with sweep_block; otherwise use gather_in_order)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert N % (world * block) == 0, "gather_in_place needs whole rounds"
    tail = tuple(local.shape[1:])
    rounds = N // (world * block)
    if out is None:
        out = torch.empty((N,) + tail, dtype=local.dtype, device=local.device)
    loc = local.contiguous()
    for k in range(rounds):
        dist.all_gather_into_tensor(out[k * world * block:(k + 1) * world * block], loc[k * block:(k + 1) * block], group=group)
    return out


def allreduce_mean(local_sum, N, group=None):
    """local_sum: float64 tensor [n] = sum of this rank's trajectories (b200ode_reduce_sum_device).
    Returns the ensemble mean [n] on every rank."""
    import torch.distributed as dist
    total = local_sum.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total / float(N)


class PeerGather:
    """Ordered all-gather by PUSH over NVLink peer memory, overlapped with the integration.

    `gather_in_place` is NCCL: its kernels need SMs, and the persistent integration kernel holds every SM until its launch
    ends, so an all-gather issued on a second stream only starts when the integration is over (measured on 8 B200s: the
    exchange stayed 3-4 ms of a 27 ms sweep step however it was pipelined).  Here the result buffer of every rank is a
    symmetric-memory allocation that all ranks map (torch.distributed._symmetric_memory: CUDA IPC / fabric handles over
    NVSwitch), and a rank copies each finished block of its shard straight into its final place — global block
    k * world + rank — of every peer's buffer with device-to-peer cudaMemcpyAsync: copy engines, no SMs, so the
    transfers of one group of rounds run under the integration of the next.  Nothing is staged or un-interleaved.
    Completion: the copies are stream-ordered before the (tiny) all-reduce of the ensemble mean that every step ends
    with; when that all-reduce has completed on a rank, every rank has passed its own copies, so `full` is complete
    everywhere.  Needs N % (world * block) == 0."""

    @classmethod
    def create(cls, N, tail, dtype, device, block, group=None):
        """PeerGather, or None when symmetric memory is not available — decided by ALL ranks together (the local
        allocation is tried first and the outcome all-reduced, so that no rank enters the collective rendezvous alone)."""
        import torch
        import torch.distributed as dist
        buf, err = None, ""
        try:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty((N,) + tuple(tail), dtype=dtype, device=device)
        except Exception as e:      # not built in, no fabric / IPC support, out of memory
            err = "%s: %s" % (type(e).__name__, e)
        flag = torch.tensor([1 if buf is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            return None, (err or "symmetric memory unavailable on another rank")
        obj = None
        try:        # the rendezvous is collective; a failure that is a property of the box hits every rank alike
            obj = cls(N, tail, dtype, device, block, group, _buf=buf)
        except Exception as e:
            err = "%s: %s" % (type(e).__name__, e)
        flag.fill_(1 if obj is not None else 0)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            return None, (err or "peer mapping failed on another rank")
        return obj, ""

    def __init__(self, N, tail, dtype, device, block, group=None, _buf=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert N % (self.world * block) == 0, "PeerGather needs whole rounds"
        self.N, self.block, self.rounds = N, block, N // (self.world * block)
        shape = (N,) + tuple(tail)
        self.full = _buf if _buf is not None else symm.empty(shape, dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.full, self.group)
        self.peers = [self.full if r == self.rank else self.handle.get_buffer(r, shape, dtype) for r in range(self.world)]
        # one copy stream per destination: copies to different peers run on different copy engines / NVLink ports at once
        # (a single stream moved the 400 MB of a group at ~130 GB/s: 3.3 ms exposed after the last kernel on 8 GPUs)
        self.streams = [torch.cuda.Stream(device=device) for _ in range(self.world)]
        self.torch = torch

    def push(self, local, k0, k1, after):
        """Copy rounds k0 .. k1-1 of `local` (this rank's shard, [rounds * block, ...]) into every rank's result, on the
        copy stream, once `after` (an event on the integration stream) has fired."""
        torch, b, w, r0 = self.torch, self.block, self.world, self.rank
        for j in range(w):
            dst, st = self.peers[(r0 + j) % w], self.streams[j]
            with torch.cuda.stream(st):
                st.wait_event(after)
                for k in range(k0, k1):
                    off = (k * w + r0) * b
                    dst[off:off + b].copy_(local[k * b:(k + 1) * b], non_blocking=True)

    def begin(self, cur):
        for st in self.streams:
            st.wait_stream(cur)                 # the previous consumers of this rank's buffers have been enqueued on `cur`

    def finish(self, cur):
        for st in self.streams:
            cur.wait_stream(st)                 # ... and the caller's closing collective orders the ranks
