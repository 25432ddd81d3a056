"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over gloo: contiguous query sharding, max-over-ranks timing,
and the final in-order gather of per-query results."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pplp_b200.shard import ChunkedGather, cross_shard, cross_sizes, gather_rows, max_over_ranks, shard_range, shard_sizes


def test_shard_ranges_partition_exactly():
    for nq in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(nq, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nq
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(nq, world)
            assert sum(sizes) == nq and max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, nq):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(nq, rank, world)
        # each rank "computes" its slice: blinded distance = f(query index), result rows = 4 words per query
        idx = torch.arange(lo, hi, dtype=torch.int64)
        blind = idx * 7919 + 13
        rows = torch.stack([idx, idx + 1, idx * 2, idx * idx], dim=1)
        g1 = gather_rows(blind, nq)
        g2 = gather_rows(rows, nq)
        t = max_over_ranks(1.0 + rank)
        assert t == float(world)
        if rank == 0:
            full = torch.arange(nq, dtype=torch.int64)
            assert torch.equal(g1, full * 7919 + 13)
            assert torch.equal(g2, torch.stack([full, full + 1, full * 2, full * full], dim=1))
        else:
            assert g1 is None and g2 is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nq", [(2, 11), (2, 4096), (3, 10)])
def test_gather_in_query_order_over_gloo(world, nq):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, nq), nprocs=world, join=True)


def _cross_worker(rank, world, port, npts, ncl):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plo, phi, lo, hi = cross_shard(npts, ncl, rank, world)
        assert (lo, hi) == (plo * ncl, phi * ncl)
        # the rank "evaluates" its points against all clients: result of pair (t, c) = t * 1000 + c, pair order t * ncl + c
        t = torch.arange(plo, phi, dtype=torch.int64).repeat_interleave(ncl)
        c = torch.arange(ncl, dtype=torch.int64).repeat(phi - plo)
        g = gather_rows(t * 1000 + c, npts * ncl, sizes=cross_sizes(npts, ncl, world))
        if rank == 0:
            tt = torch.arange(npts, dtype=torch.int64).repeat_interleave(ncl)
            cc = torch.arange(ncl, dtype=torch.int64).repeat(npts)
            assert torch.equal(g, tt * 1000 + cc)
        else:
            assert g is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,npts,ncl", [(2, 7, 5), (3, 4, 3)])
def test_config5_point_sharding_gathers_in_pair_order(world, npts, ncl):
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_cross_worker, args=(world, port, npts, ncl), nprocs=world, join=True)


def _chunk_worker(rank, world, port, nchunks):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cg = ChunkedGather(torch.empty(64, 3, dtype=torch.int64))
        seen = []
        for i in range(nchunks):
            buf = cg.buffer(i)                       # waits for chunk i-2's gather before the buffer is overwritten
            if rank == 0 and i >= 2:
                seen.append([t.clone() for t in cg.received(i)])   # ... so chunk i-2's data is complete here
            buf.copy_(torch.full((64, 3), 1000 * i + rank, dtype=torch.int64))
            cg.submit(i)
        cg.finish()
        if rank == 0:
            for i in range(max(0, nchunks - 2), nchunks):
                seen_i = cg.received(i)
                assert all(bool((seen_i[r] == 1000 * i + r).all()) for r in range(world)), i
            for i, parts in enumerate(seen):
                assert all(bool((parts[r] == 1000 * i + r).all()) for r in range(world)), i
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nchunks", [(2, 5), (3, 2)])
def test_chunked_double_buffered_gather(world, nchunks):
    """The exchange bench.py times: chunk i's gather in flight while chunk i+1 is produced, two buffers."""
    port = 33500 + (os.getpid() % 2000) + world
    mp.spawn(_chunk_worker, args=(world, port, nchunks), nprocs=world, join=True)
