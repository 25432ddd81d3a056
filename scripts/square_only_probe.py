"""One batch of BEHZ squares (N=8192, BFVDefault) — the workload for ncu instruction-count captures of the square's kernels.
usage: python scripts/square_only_probe.py [--nq 512] [--reps 2]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pplp_b200 import engine
ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, default=512)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
ctx = engine.Context(8192, t=1 << 20, device=0)
ct = ctx.empty(a.nq, 2, ctx.k, 8192)
for j in range(ctx.k):
    ct[:, :, j].random_(0, ctx.q[j])
for _ in range(2):
    ctx.square(ct)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ctx.square(ct)
e1.record()
torch.cuda.synchronize()
print({"squares_per_s": a.nq * a.reps / (e0.elapsed_time(e1) * 1e-3), "nq": a.nq})
