"""BASELINE.json config 4: NTT / INTT + relinearize (+ square) microbenchmark sweep, N = 4096..32768, 3..8 RNS limbs.

Moduli: the first L primes of BFVDefault(N), topped up with get_primes(2N, 50) when L exceeds the table.
NTT GB/s    = 16*N bytes per limb transform (read + write; twiddles are batch-amortised, L2-resident)
relin GB/s  = 40*k*N bytes per ciphertext (read c2, read + write c0, c1; keys amortised)           [k = L - 1 data limbs]
square GB/s = 8*k*N*(2 + 3) bytes per ciphertext (read 2 polynomials, write 3)
usage: python scripts/config4_sweep.py [--out gpurun_out/config4.json] [--quick]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pplp_b200 import engine


def primes_for(n, limbs):
    q = engine.bfv_default(n)[:limbs]
    cand = ((1 << 50) - 1) // (2 * n) * (2 * n) + 1
    while len(q) < limbs:   # descending primes == 1 mod 2N below 2^50, skipping ones already taken
        if cand not in q and _is_prime(cand):
            q.append(cand)
        cand -= 2 * n
    return q


def _is_prime(v):
    if v < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if v % p == 0:
            return v == p
    d, r = v - 1, 0
    while d % 2 == 0:
        d //= 2; r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, v)
        if x in (1, v - 1):
            continue
        for _ in range(r - 1):
            x = x * x % v
            if x == v - 1:
                break
        else:
            return False
    return True


def timed(fn, reps):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def sweep(ns=(4096, 8192, 16384, 32768), limb_list=(3, 4, 5, 6, 7, 8), device=0, verbose=True, reps=5):
    """Rows of {n, limbs, ntt_gbs, intt_gbs, relin_per_s, relin_gbs, square_per_s, square_gbs}; also called by bench.py (reduced grid)."""
    rows_out = []
    for n in ns:
        for limbs in limb_list:
            q = primes_for(n, limbs)
            ctx = engine.Context(n, q=q, t=1 << 20, device=device, enforce_security=False)
            if not ctx.ok:
                if verbose:
                    print("skip", n, limbs, ctx.error_message)
                continue
            rows = max(64, (1 << 25) // n)           # >= 256 MiB of data per launch at every N
            data = ctx.empty(limbs, 1, rows, n)
            for j in range(limbs):
                data[j].random_(0, q[j])
            rec = {"n": n, "limbs": limbs, "max_bits": max(x.bit_length() for x in q), "rows_per_launch": rows * limbs}
            for inv in (False, True):
                s = timed(lambda: ctx.ntt_(data, level=0, inverse=inv, layout=engine.LAYOUT_LIMB_MAJOR), reps)
                rec["intt_gbs" if inv else "ntt_gbs"] = round(16 * n * rows * limbs / s / 1e9, 1)
            # relinearize / square at the data level (k = limbs - 1)
            k = ctx.k
            nq = max(8, (1 << 22) // n)
            sk, pk = ctx.keygen(np.arange(8, dtype=np.uint64) + 1)
            rk = ctx.relin_keygen(np.tile(np.arange(8, dtype=np.uint64) + 3, (k, 1)), sk)
            quot = ctx.relin_prepare(rk)
            ct3 = ctx.empty(k, 3, nq, n)
            for j in range(k):
                ct3[j].random_(0, q[j])
            s = timed(lambda: ctx.relinearize(ct3, rk, quot, layout=engine.LAYOUT_LIMB_MAJOR), 3)
            rec["relin_per_s"] = round(nq / s, 1)
            rec["relin_gbs"] = round(40 * k * n * nq / s / 1e9, 1)
            ct2 = ct3[:, :2].contiguous()
            s = timed(lambda: ctx.square(ct2, layout=engine.LAYOUT_LIMB_MAJOR), 3)
            rec["square_per_s"] = round(nq / s, 1)
            rec["square_gbs"] = round(8 * k * n * 5 * nq / s / 1e9, 1)
            rows_out.append(rec)
            if verbose:
                print(json.dumps(rec), flush=True)
            del data, ct3, ct2, rk, quot, ctx
            torch.cuda.empty_cache()
    return rows_out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/config4.json")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--n", type=int, nargs="*", default=[4096, 8192, 16384, 32768])
    a = ap.parse_args()
    rows_out = sweep(ns=tuple(a.n), limb_list=(3, 8) if a.quick else (3, 4, 5, 6, 7, 8))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for r in rows_out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
