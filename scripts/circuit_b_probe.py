"""One batch of Circuit-B groups (N=8192, BFVDefault, t = Batching(8192, 56)) on random residues — the workload for ncu launch
lists of pplp_circuit_b's kernels.  usage: python scripts/circuit_b_probe.py [--groups 512] [--chunk 256] [--reps 4]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pplp_b200 import engine
ap = argparse.ArgumentParser()
ap.add_argument("--groups", type=int, default=512)
ap.add_argument("--chunk", type=int, default=256)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
n, t = 8192, 0xfffffffffb4001
ctx = engine.Context(n, t=t, device=0)
k, K, g = ctx.k, len(ctx.q), a.groups
def rand_ct(*lead):
    x = ctx.empty(*lead, k, n)
    for j in range(k):
        x[..., j, :].random_(0, ctx.q[j])
    return x
cx, cy = rand_ct(g, 2), rand_ct(g, 2)
rk = ctx.empty(k, 2, K, n)
for j in range(K):
    rk[:, :, j].random_(0, ctx.q[j])
quot = ctx.relin_prepare(rk)
px, py, pr = (ctx.empty(g, n).random_(0, t) for _ in range(3))
sv = ctx.empty(g).random_(1, 8)
out = ctx.empty(*ctx.ct_shape(g, 2))
run = lambda: ctx.circuit_b(cx, cy, px, py, pr, sv, rk, quot, out=out, chunk=a.chunk)
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    run()
e1.record()
torch.cuda.synchronize()
print({"groups_per_s": g * a.reps / (e0.elapsed_time(e1) * 1e-3), "groups": g, "chunk": a.chunk})
