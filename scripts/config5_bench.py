"""BASELINE.json config 5: server-side batch, every one of NCL clients against NPTS server points, each point with its own
blinds (r, s, w) and Bloom filter; every evaluation followed (bench only: it plays the client) by decryption and the
Bloom-filter query.  Real keys, real encryptions, verdicts checked against the plaintext distances.
Reports pairs/s for (a) the evaluation alone, (b) evaluation + decrypt + Bloom verdict, with the HBM figures:
  compulsory bytes per pair = 16*k*N (output write) + 48*k*N / NPTS (client ciphertexts, read once per launch)
Multi-GPU: under torchrun the server points are sharded across ranks (pplp_b200.shard.cross_shard), every rank keeps all
clients, no collective in the data path; times are the max over ranks and pairs/s is the whole-job figure.
usage: python scripts/config5_bench.py [--clients 1000] [--points 10000] [--tile 16] [--max-tiles 12]
       python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/config5_bench.py ..."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from pplp_b200 import engine
from pplp_b200.shard import cross_shard, max_over_ranks

ap = argparse.ArgumentParser()
ap.add_argument("--clients", type=int, default=1000)
ap.add_argument("--points", type=int, default=10000)
ap.add_argument("--tile", type=int, default=16, help="server points per launch (output = tile*clients ciphertexts)")
ap.add_argument("--max-tiles", type=int, default=12, help="time this many tiles and extrapolate (0 = all)")
ap.add_argument("--radius", type=int, default=128)
ap.add_argument("--n", type=int, default=8192)
a = ap.parse_args()
T56 = 1 << 56
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = engine.Context(a.n, t=T56, device=local)
k, n = ctx.k, a.n
sk, pk = ctx.keygen(np.arange(1, 9, dtype=np.uint64))
rng = np.random.default_rng(5)
ncl, npts_all = a.clients, a.points
p_lo, p_hi, _, _ = cross_shard(npts_all, ncl, rank, world)
npts = npts_all                      # all ranks draw the same server points and clients; each evaluates [p_lo, p_hi)
px = rng.integers(1000, 1 << 27, npts, dtype=np.uint64); py = rng.integers(1000, 1 << 27, npts, dtype=np.uint64)
rsw = np.stack([rng.integers(0, 1 << 32, npts, dtype=np.uint64), rng.integers(1, 1 << 32, npts, dtype=np.uint64), rng.integers(1, 1 << 16, npts, dtype=np.uint64)], axis=1)
# clients sit near server points that some rank actually times, so that both verdicts occur in every rank's checked tile
cand = np.concatenate([np.arange(cross_shard(npts_all, ncl, r, world)[0],
                                 min(cross_shard(npts_all, ncl, r, world)[1], cross_shard(npts_all, ncl, r, world)[0] + a.tile * (a.max_tiles or npts_all)))
                       for r in range(world)])
home = cand[rng.integers(0, len(cand), ncl)]
cx = px[home] + rng.integers(0, 150, ncl).astype(np.uint64); cy = py[home] + rng.integers(0, 150, ncl).astype(np.uint64)
# client side: x^2+y^2, 2x, 2y as constant plaintexts (src/client.cc:96-113), limb-major batch of ncl ciphertexts each
LM = engine.LAYOUT_LIMB_MAJOR
seeds = rng.integers(0, 1 << 63, size=(3, ncl, 8), dtype=np.uint64)
plains = [(cx * cx + cy * cy) % T56, (2 * cx) % T56, (2 * cy) % T56]
cts = [ctx.encrypt(pk, ctx.dev(seeds[i]), ctx.dev(plains[i].reshape(ncl, 1)), layout=LM) for i in range(3)]
out = ctx.empty(*ctx.ct_shape(a.tile * ncl, 2, None, LM))
bf_base = p_lo                       # Bloom filters of this rank's points only (of the timed ones when --max-tiles is set)
bf_hi = min(p_hi, p_lo + a.tile * a.max_tiles) if a.max_tiles else p_hi
bf_all = engine.BloomBatch(ctx, a.radius, fpp=1e-4, rsw=rsw[bf_base:bf_hi]).build()
my_pts = p_hi - p_lo
ntiles = (my_pts + a.tile - 1) // a.tile
timed = ntiles if a.max_tiles == 0 else min(ntiles, a.max_tiles)
d_px, d_py = ctx.dev(px), ctx.dev(py)
d_r, d_s = ctx.dev(np.ascontiguousarray(rsw[:, 0])), ctx.dev(np.ascontiguousarray(rsw[:, 1]))


def run_tile(t, with_client):
    lo, hi = p_lo + t * a.tile, min(p_hi, p_lo + (t + 1) * a.tile)
    ctx.circuit_a_cross(cts[0], cts[1], cts[2], d_px[lo:hi], d_py[lo:hi], d_r[lo:hi], d_s[lo:hi], out=out, layout=LM)
    if not with_client:
        return None
    npair = (hi - lo) * ncl
    blind = ctx.decrypt(out, sk, ncoeff=1, layout=LM)[:npair, 0]
    fidx = torch.arange(lo - bf_base, hi - bf_base, dtype=torch.int32, device=ctx.device).repeat_interleave(ncl)
    return blind, bf_all.query(blind.contiguous(), fidx=fidx)


res = {"workload": f"config5: {ncl} clients x {npts_all} server points, N={n}, k={k}, radius {a.radius}", "n_gpus": world, "tile_points": a.tile,
       "tiles_timed_per_gpu": timed}
for with_client in (False, True):
    run_tile(0, with_client)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(timed):
        last = run_tile(t, with_client)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), device=ctx.device)
    pairs = sum(min(my_pts, (t + 1) * a.tile) - t * a.tile for t in range(timed)) * ncl * world   # whole job (ranks hold equal shares +-1 point)
    key = "eval_decrypt_bloom" if with_client else "eval_only"
    res[key] = {"pairs_per_s": pairs / (ms * 1e-3), "ms_per_tile": ms / timed, "whole_job_s_extrapolated": ms * 1e-3 * ntiles / timed}
    if not with_client:
        wr = 16 * k * n
        res[key]["hbm_write_gbs_per_gpu"] = pairs * wr / (ms * 1e-3) / 1e9 / world
        res[key]["compulsory_bytes_per_pair"] = wr + 48 * k * n / a.tile
        res[key]["naive_bytes_per_pair"] = 64 * k * n
# correctness of the last timed tile: blinded distance and verdict for every pair
t = timed - 1
lo, hi = p_lo + t * a.tile, min(p_hi, p_lo + (t + 1) * a.tile)
blind, verdict = last
d2 = (cx[None, :].astype(np.int64) - px[lo:hi, None].astype(np.int64)) ** 2 + (cy[None, :].astype(np.int64) - py[lo:hi, None].astype(np.int64)) ** 2
expect = (rsw[lo:hi, 1][:, None] * (d2.astype(np.uint64) + rsw[lo:hi, 0][:, None])) & np.uint64(T56 - 1)
res["blind_distances_correct"] = bool((engine.to_np(blind).reshape(hi - lo, ncl) == expect).all())
near = d2 < a.radius * a.radius
got = verdict.cpu().numpy().astype(bool).reshape(hi - lo, ncl)
res["no_false_negatives"] = bool(got[near].all())
res["false_positive_rate"] = float(got[~near].mean())
res["near_pairs_in_checked_tile"] = int(near.sum())
ok = all(bool(res[f]) for f in ("blind_distances_correct", "no_false_negatives"))
if world > 1:
    flag = torch.tensor([1 if ok else 0], device=ctx.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["all_ranks_correct"] = bool(flag.item())
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
