"""NTT / INTT / fused-kernel microbenchmark (BASELINE.json config 4): GB/s = 16*N bytes per limb transform / time.
usage: python scripts/ntt_bench.py [--n 8192] [--rows-per-limb 4096] [--reps 10]"""
import argparse
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pplp_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[8192])
ap.add_argument("--rows-per-limb", type=int, default=4096)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--limbs", type=int, default=0, help="use only the first L primes of BFVDefault (0 = data level)")
a = ap.parse_args()
res = []
for n in a.n:
    q = engine.bfv_default(n)
    ctx = engine.Context(n, q=q, t=1 << 20, device=0)
    level = ctx.first_level
    k = ctx.limbs(level)
    rows = a.rows_per_limb * 8192 // n
    data = ctx.empty(k, 1, rows, n)
    for j in range(k):
        data[j].random_(0, ctx.q[j])
    out = {"n": n, "limbs": k, "rows": rows * k, "bits": max(x.bit_length() for x in ctx.q[:k])}
    for inv in (False, True):
        for _ in range(3):
            ctx.ntt_(data, level=level, inverse=inv, layout=engine.LAYOUT_LIMB_MAJOR)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
        ev[0].record()
        for i in range(a.reps):
            ctx.ntt_(data, level=level, inverse=inv, layout=engine.LAYOUT_LIMB_MAJOR)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[-1]) / a.reps                                  # mean over the back-to-back launches (reported)
        best = min(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))          # fastest single launch
        out["intt_gbs" if inv else "ntt_gbs"] = round(16 * n * rows * k / (ms * 1e-3) / 1e9, 1)
        out["intt_ms" if inv else "ntt_ms"] = round(ms, 4)
        out["intt_best_gbs" if inv else "ntt_best_gbs"] = round(16 * n * rows * k / (best * 1e-3) / 1e9, 1)
    res.append(out)
    print(json.dumps(out), flush=True)
