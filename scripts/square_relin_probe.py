"""Square + relinearise on a batch of ciphertexts (N=8192, BFVDefault): throughput, and the workload for ncu captures of
the BEHZ and key-switching kernels.  usage: python scripts/square_relin_probe.py [--nq 512] [--reps 5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pplp_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, default=512)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
n = 8192
ctx = engine.Context(n, t=1 << 20, device=0)
k, nq = ctx.k, a.nq
sk, pk = ctx.keygen(np.arange(1, 9, dtype=np.uint64))
rk = ctx.relin_keygen(np.stack([np.arange(8, dtype=np.uint64) + np.uint64(100 * (i + 1)) for i in range(k)]), sk)
quot = ctx.relin_prepare(rk)
ct = ctx.empty(nq, 2, k, n)
for j in range(k):
    ct[:, :, j].random_(0, ctx.q[j])


def timed(f):
    for _ in range(2):
        out = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = f()
    e1.record()
    torch.cuda.synchronize()
    return nq * a.reps / (e0.elapsed_time(e1) * 1e-3), out


sq_rate, sq = timed(lambda: ctx.square(ct))
rl_rate, _ = timed(lambda: ctx.relinearize(sq, rk, quot))
print({"squares_per_s": sq_rate, "relin_per_s": rl_rate, "nq": nq, "n": n, "k": k})
