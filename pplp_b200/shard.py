"""Multi-GPU plumbing: queries are independent, so a batch is sharded by contiguous slices, one process per GPU, with no
collective inside the hot path.  The only exchanges are the timing reduction and an optional final gather
of results (blinded distances / verdicts, or result ciphertexts), both through torch.distributed (NCCL on GPUs, gloo in
the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(nq, rank, world):
    """Contiguous slice [lo, hi) of nq queries owned by `rank`; sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(nq, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(nq, world):
    return [shard_range(nq, r, world)[1] - shard_range(nq, r, world)[0] for r in range(world)]


def max_over_ranks(value, device="cpu"):
    """Max of a scalar over all ranks (how every multi-GPU time is reported)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cross_shard(npts, ncl, rank, world):
    """BASELINE.json config 5 (every client against every server point): the server points are sharded, every rank keeps
    all `ncl` client ciphertexts.  Returns (point_lo, point_hi, pair_lo, pair_hi): the rank's points and the slice of the
    global pair order (pair = point * ncl + client) its results occupy — contiguous, so gather_rows applies."""
    lo, hi = shard_range(npts, rank, world)
    return lo, hi, lo * ncl, hi * ncl


def cross_sizes(npts, ncl, world):
    return [s * ncl for s in shard_sizes(npts, world)]


def gather_rows(local, nq, dst=0, sizes=None):
    """Gathers per-query rows (blinded distances, verdicts or whole result ciphertexts, first dimension = local queries)
    from every rank to `dst` in query order.  Returns the [nq, ...] tensor on dst, None elsewhere.  `sizes`: rows per
    rank when the batch is not sharded by shard_range (e.g. cross_sizes)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(nq, world) if sizes is None else list(sizes)
    assert sum(sizes) == nq
    assert local.shape[0] == sizes[rank]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst)
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    dist.gather(buf, None, dst=dst)
    return None


class ChunkedGather:
    """Result chunks to `dst`, double-buffered and asynchronous: chunk i's gather runs on the communication stream while the
    caller produces chunk i+1 (end-to-end = max(compute, gather)).  Usage per chunk i:
        buf = cg.buffer(i)       # waits until the gather that last used this buffer (chunk i-2) is done
        ... fill buf ...
        cg.submit(i)
    then cg.finish().  On dst, received(i) is the list of per-rank tensors of chunk i (valid until chunk i+2 is submitted)."""

    def __init__(self, like, dst=0):
        self.dst = dst
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.bufs = [torch.empty_like(like) for _ in range(2)]
        self.recv = [[torch.empty_like(like) for _ in range(self.world)] for _ in range(2)] if self.rank == dst else [None, None]
        self.pending = [None, None]

    def buffer(self, i):
        if self.pending[i & 1] is not None:
            self.pending[i & 1].wait()
            self.pending[i & 1] = None
        return self.bufs[i & 1]

    def submit(self, i):
        if self.world == 1:
            self.recv[i & 1][0].copy_(self.bufs[i & 1])
            return
        self.pending[i & 1] = dist.gather(self.bufs[i & 1], self.recv[i & 1] if self.rank == self.dst else None, dst=self.dst, async_op=True)

    def received(self, i):
        return self.recv[i & 1]

    def finish(self):
        for k in range(2):
            if self.pending[k] is not None:
                self.pending[k].wait()
                self.pending[k] = None
