"""Runs the whole-protocol batch (3 encrypts + Circuit A + decrypt + Bloom) a few times; used under
`ncu --metrics gpu__time_duration.sum` to see where the protocol's GPU time goes, and standalone for queries/s.
usage: python scripts/protocol_probe.py [--nq 1024] [--reps 3] [--n 8192]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pplp_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--chunk", type=int, default=1024)
a = ap.parse_args()
ctx = engine.Context(a.n, t=1 << 56, device=0)
seed = np.arange(1, 9, dtype=np.uint64)
sk, pk = ctx.keygen(seed)
rng = np.random.default_rng(1)
r, s, w = 0x12345678, 0x9ABCDEF1, 0xBEEF
nq = a.nq
xb = np.full(nq, 123456888, dtype=np.uint64); yb = np.full(nq, 132465777, dtype=np.uint64)
xa = xb + rng.integers(0, 300, nq).astype(np.uint64); ya = yb + rng.integers(0, 300, nq).astype(np.uint64)
seeds = rng.integers(0, 1 << 63, size=(nq * 3, 8), dtype=np.uint64)
bf = engine.BloomBatch(ctx, 128, fpp=1e-4, rsw=[(r, s, w)]).build()
d = [ctx.dev(x) for x in (xa, ya, xb, yb, seeds)]
ctx.proximity_batch(pk, sk, d[0], d[1], d[2], d[3], d[4], bf, chunk=a.chunk)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.reps):
    blind, verdict, flags = ctx.proximity_batch(pk, sk, d[0], d[1], d[2], d[3], d[4], bf, chunk=a.chunk)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
d2 = (xa.astype(np.int64) - xb.astype(np.int64)) ** 2 + (ya.astype(np.int64) - yb.astype(np.int64)) ** 2
ok = (engine.to_np(blind) == ((np.uint64(s) * (d2.astype(np.uint64) + np.uint64(r))) & np.uint64((1 << 56) - 1))).all()
print({"queries_per_s": nq * a.reps / dt, "nq": nq, "correct": bool(ok), "near": float(verdict.float().mean())})
