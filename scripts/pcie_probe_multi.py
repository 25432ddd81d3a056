"""Host<->device copy bandwidth with every rank copying AT THE SAME TIME (one process per GPU), pinned buffers from cudaHostAlloc
placed on each GPU's NUMA node where the cpuset allows it.  Answers: what is the box's aggregate host-fabric ceiling for the
host-buffer entry points (pplp_circuit_a_host), and does NUMA placement move it?

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/pcie_probe_multi.py [--no-bind]
       python scripts/pcie_probe_multi.py            (one GPU)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-bind", action="store_true")
    ap.add_argument("--gib", type=float, default=1.0)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    from pplp_b200 import capi, numa
    rep = {"bound": False, "why": "--no-bind"} if a.no_bind else numa.bind_to_gpu_node(local)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    L = capi.lib()
    n = int(a.gib * (1 << 30)) // 8
    hsrc = numa.pinned_empty(L, (n,), torch.int64, write_combined=True)
    hdst = numa.pinned_empty(L, (n,), torch.int64)
    d1 = torch.empty(n, dtype=torch.int64, device="cuda")
    d2 = torch.zeros(n, dtype=torch.int64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, reps=4):
        fn(); barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return reps * n * 8 / dt / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            d1.copy_(hsrc, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            hdst.copy_(d2, non_blocking=True)

    def both():
        h2d(); d2h()

    res = {"h2d": timed(h2d), "d2h": timed(d2h), "bidir_each_way": timed(both)}
    t = torch.tensor([res["h2d"], res["d2h"], res["bidir_each_way"]], dtype=torch.float64, device="cuda")
    lo = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        reps = [None] * world
        dist.all_gather_object(reps, rep)
    else:
        reps = [rep]
    if rank == 0:
        out = {"ranks": world, "bind": not a.no_bind, "GiB_per_copy": a.gib,
               "aggregate_GBps": dict(zip(["h2d", "d2h", "bidir_each_way"], [float(x) for x in t])),
               "slowest_rank_GBps": dict(zip(["h2d", "d2h", "bidir_each_way"], [float(x) for x in lo])),
               "numa": reps, "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t"),
               "note": "all ranks copy concurrently; H2D source is write-combined pinned memory, each buffer allocated after binding the rank to its GPU's node"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
