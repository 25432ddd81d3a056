#!/usr/bin/env python
"""bench.py — headline benchmark of the pplp hot path on B200 (contract: see the task's bench section).

Metric   proximity queries/s: the reference's server-side encrypted squared-distance evaluation with random blinding
         ("Circuit A", /root/reference src/server.cc:127-133) at BFV N=8192, BFVDefault coefficient modulus (k=4 data
         limbs), t=2^56, batched over independent client queries.  BASELINE.json configs[1].
Step     one pass of the fused evaluation kernel over one batch of Q synthetic queries per GPU (3 ciphertexts in,
         1 out, 64*k*N = 2 MiB of HBM traffic per query).  Inputs are uniform residues generated on the device (for
         timing they are indistinguishable from ciphertexts); a parity subset of REAL encryptions is evaluated and
         checked against the oracle before the clock starts.
value    whole-job queries/s with inputs resident in HBM (max over ranks, CUDA events).
e2e      the same evaluation through the C-ABI host entry (pplp_circuit_a_host): ciphertexts in page-locked HOST
         memory, H2D + kernel + D2H inside the timed region.
extras   protocol (encrypt x3 -> Circuit A -> decrypt -> Bloom verdict, host coordinates in, verdicts out) and the NTT
         GB/s microbenchmark, so the other kernels have a measured number too.
--impl reference   times the CPU restatement of the reference path (oracle/) on the box's host cores, same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when the
# first communicator comes up), so everything but the result line is sent to stderr: fd 1 is pointed at fd 2 for the
# whole run and the line is written to a saved duplicate of the original stdout.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 8192
T = 1 << 56
METRIC = "proximity_queries_per_sec"
UNIT = "queries/s"
WORKLOAD = "circuitA_bfv_n8192_k4_t2^56_batched"
LIMBS = 4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--degree", type=int, default=8192, help="poly_modulus_degree (8192 = BASELINE configs[1]; 16384 = configs[2])")
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU per step (resident run); 0 = 8192 at N=8192, scaled by 8192/N")
    ap.add_argument("--e2e-queries", type=int, default=1024, help="queries per GPU per step (host-buffer run)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = calibrate to ~10 s)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append((time.time(), parts))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [p for (ts, p) in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [p for (_, p) in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows), "power_w_max": max(pw) if pw else None}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def seed8(x):
    return np.array([(x * 0x9E3779B97F4A7C15 + i * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF for i in range(8)], dtype=np.uint64)


def synth_host_batch(q, nq, seed):
    """[nq][2][k][N] uniform residues + per-query plaintext constants (host, numpy)."""
    rng = np.random.default_rng(seed)
    k = len(q)
    c = []
    for _ in range(3):
        a = np.empty((nq, 2, k, N), dtype=np.uint64)
        for j in range(k):
            a[:, :, j, :] = rng.integers(0, q[j], size=(nq, 2, N), dtype=np.uint64)
        c.append(a)
    xb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    yb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    r = rng.integers(0, 1 << 32, nq, dtype=np.uint64)
    s = rng.integers(1, 1 << 32, nq, dtype=np.uint64)
    return c, xb, yb, r, s


# ---------------------------------------------------------------------------------------------------------------------
def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


CPU_Q_PER_THREAD = 32      # queries per thread per step (a thread sees a steady stream, not two queries)
CPU_SAMPLE_CAP = 2048      # host memory bound: 1.5 MiB of input per query at N=8192


def cpu_circuit_a_rate(steps, warmup, sample, threads):
    """CPU restatement of the reference path (oracle/, 'port'): the seven Evaluator calls of src/server.cc:127-133 per
    query, in place on host-resident ciphertexts (no copies, no allocation in the timed loop), Shoup products as in
    SEAL's multiply_poly_scalar_coeffmod.  Returns all-thread and one-thread rates plus the host STREAM figure."""
    import ctypes as C
    from tests import oracle_lib
    orc = oracle_lib.load()
    q = orc.bfv_default(N)
    octx = orc.context(N, q, T, seed=seed8(1))
    k = octx.k
    scale = max(1, (N * k) // (8192 * 4))
    if sample <= 0:
        sample = min(max(threads * CPU_Q_PER_THREAD // scale, threads), max(CPU_SAMPLE_CAP // scale, threads))
    c, xb, yb, r, s = synth_host_batch(q[:k], sample, 4)

    def run(nsteps, nth, nq):
        t0 = time.perf_counter()
        for _ in range(nsteps):
            octx.circuit_a_batch(c[0][:nq], c[1][:nq], c[2][:nq], xb[:nq], yb[:nq], r[:nq], s[:nq], nthreads=nth, inplace=True)
        return time.perf_counter() - t0

    run(max(warmup, 1), threads, sample)
    dt = run(steps, threads, sample)
    n1 = max(1, min(sample, CPU_Q_PER_THREAD // scale))
    run(1, 1, n1)
    dt1 = run(3, 1, n1)
    info = {"threads": threads, "nproc": os.cpu_count(), "cpu_model": cpu_model(), "one_thread_value": n1 * 3 / dt1,
            "thread_scaling": (sample * steps / dt) / (n1 * 3 / dt1), "queries_per_thread_per_step": sample / threads}
    try:
        f = orc.lib.orc_stream_add_gbs
        f.restype = C.c_double
        f.argtypes = [C.c_size_t, C.c_int, C.c_int]
        info["host_stream_add_GBps"] = {"one_thread": f(1 << 24, 4, 1), "all_threads": f(1 << 26, 4, threads)}
        # per query the seven calls move 3 read-modify-write passes of one ciphertext and 2 read-read-write passes
        info["evaluator_bytes_per_query"] = 12 * 16 * k * N
        info["evaluator_GBps"] = info["evaluator_bytes_per_query"] * sample * steps / dt / 1e9
    except Exception as e:
        info["host_stream_add_GBps"] = {"error": str(e)[:100]}
    return sample * steps / dt, dt / steps, sample, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    rate, per_step, sample, info = cpu_circuit_a_rate(args.steps, args.warmup, args.cpu_sample, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_per_step": sample, "poly_modulus_degree": N, "limbs": LIMBS, "plain_modulus": "2^56"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} queries/step x {args.steps} steps in place on host-resident ciphertexts, SEAL-4.1-equivalent CPU "
                                   f"restatement (oracle/), {threads} host threads", **info},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
def parity_gate(engine, ctx):
    """Real encryptions through the timed entry point must equal the oracle bit for bit before any timing counts."""
    from tests import oracle_lib
    orc = oracle_lib.load()
    octx = orc.context(N, ctx.q, T, seed=seed8(7))
    osk, opk = octx.keygen()
    nq = 4
    rng = np.random.default_rng(5)
    xa = rng.integers(0, 1 << 27, nq, dtype=np.uint64); ya = rng.integers(0, 1 << 27, nq, dtype=np.uint64)
    xb = rng.integers(1, 1 << 27, nq, dtype=np.uint64); yb = rng.integers(1, 1 << 27, nq, dtype=np.uint64)
    r = rng.integers(0, 1 << 32, nq, dtype=np.uint64); s = rng.integers(1, 1 << 32, nq, dtype=np.uint64)
    vals = [xa * xa + ya * ya, xa << np.uint64(1), ya << np.uint64(1)]
    cts = [np.stack([octx.encrypt(opk, [int(v[i])], seed=seed8(50 + 3 * i + j)) for i in range(nq)]) for j, v in enumerate(vals)]
    ref = octx.circuit_a_batch(cts[0], cts[1], cts[2], xb, yb, r, s, nthreads=2)
    lm = [ctx.dev(np.ascontiguousarray(c.transpose(2, 1, 0, 3))) for c in cts]
    out = ctx.circuit_a(lm[0], lm[1], lm[2], ctx.dev(xb), ctx.dev(yb), ctx.dev(r), ctx.dev(s), layout=engine.LAYOUT_LIMB_MAJOR)
    got = engine.to_np(out).transpose(2, 1, 0, 3)
    if not (got == ref).all():
        raise SystemExit("bench: parity gate failed — CUDA Circuit A differs from the oracle")
    dec = engine.to_np(ctx.decrypt(ctx.dev(np.ascontiguousarray(got)), ctx.dev(osk), ncoeff=1))[:, 0]
    for i in range(nq):
        d2 = (int(xa[i]) - int(xb[i])) ** 2 + (int(ya[i]) - int(yb[i])) ** 2
        if int(dec[i]) != (int(s[i]) * (d2 + int(r[i]))) % T:
            raise SystemExit("bench: parity gate failed — decrypted blind distance is wrong")
    return osk, opk


def measure_exchange(torch, dist, ctx, engine, cin, out, xb, yb, rr, ss, Q, k, world, rank, dev, barrier):
    """What leaves a GPU after the evaluation.  (1) default: 8 bytes per query (the decrypted blinded distance / verdict a
    co-located client stage produces) — gathered to rank 0, negligible.  (2) result CIPHERTEXTS to rank 0 (a server that must
    hand them to one network front end): chunks of >= 1 GiB per rank, gather of chunk i on NCCL's stream while chunk i+1 is
    evaluated, timed as compute only / gather only / overlapped; rank 0's ingest against NVLink (900 GB/s nominal per
    direction, 770 GB/s measured peer copy)."""
    from pplp_b200.shard import max_over_ranks
    per_ct = 2 * k * N * 8
    cq = max(1, min(Q // 2, (1 << 30) // per_ct))                  # queries per chunk: 1 GiB of result ciphertexts per rank
    nchunks = max(2, min(4, Q // cq))
    src = [c[:, :, :cq, :].contiguous() for c in cin]              # one contiguous chunk of inputs, re-evaluated for every chunk
    from pplp_b200.shard import ChunkedGather
    cg = ChunkedGather(ctx.empty(*ctx.ct_shape(cq, 2, None, engine.LAYOUT_LIMB_MAJOR)))      # double-buffered contiguous chunk results
    outs, recv = cg.bufs, cg.recv
    sl = lambda t, i: t[i * cq:(i + 1) * cq]

    def compute(i, out_buf):
        ctx.circuit_a(src[0], src[1], src[2], sl(xb, i), sl(yb, i), sl(rr, i), sl(ss, i), out=out_buf, layout=engine.LAYOUT_LIMB_MAJOR)

    def run(do_compute, do_gather):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(nchunks):
            buf = cg.buffer(i)                   # the buffer pair of chunk i-2 is free again
            if do_compute:
                compute(i, buf)
            if do_gather:
                cg.submit(i)
        cg.finish()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b), dev) * 1e-3

    run(True, True)                                                  # warm-up: communicators, allocations
    t_c, t_g, t_o = run(True, False), run(False, True), run(True, True)
    ok = True
    if rank == 0:
        ok = bool(torch.equal(recv[(nchunks - 1) & 1][0], outs[(nchunks - 1) & 1]))
    # default exchange: 8 bytes per query
    small = torch.arange(Q, dtype=torch.int64, device=dev)
    parts = [torch.empty_like(small) for _ in range(world)] if rank == 0 else None
    dist.gather(small, parts, dst=0)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    dist.gather(small, parts, dst=0)
    b.record()
    barrier()
    t_small = max_over_ranks(a.elapsed_time(b), dev) * 1e-3
    if rank != 0:
        return None
    ingest = (world - 1) * nchunks * cq * per_ct
    gbps = ingest / t_g / 1e9
    return {"default_exchange": {"what": "8-byte blinded distance (or verdict) per query to rank 0", "queries_per_rank": Q, "seconds": t_small,
                                 "queries_per_s": world * Q / t_small},
            "ciphertext_gather": {"chunk_GiB_per_rank": cq * per_ct / 2**30, "chunks": nchunks, "ciphertexts": world * nchunks * cq,
                                  "compute_only_s": t_c, "gather_only_s": t_g, "overlapped_s": t_o, "binds": "gather" if t_g > t_c else "compute",
                                  "overlap_efficiency": max(t_c, t_g) / t_o, "rank0_ingest_GBps": gbps, "result_ciphertexts_per_s": world * nchunks * cq / t_o,
                                  "nvlink_nominal_GBps": 900.0, "frac_of_nominal": gbps / 900.0, "nvlink_measured_peer_copy_GBps": 770.0, "frac": gbps / 770.0,
                                  "ok": ok},
            "note": "NCCL gather to rank 0 on its own stream, chunk i in flight while chunk i+1 is evaluated; the hot-path figures leave outputs on the producing GPU"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pplp_b200 import build, engine
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench: no CUDA device — pplp_b200 has no CPU fallback")
    build.build()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"
    ctx = engine.Context(N, t=T, device=local)
    k = ctx.k
    osk, opk = parity_gate(engine, ctx) if rank == 0 else (None, None)

    # ---- resident run: limb-major [k][2][Q][N] slabs, uniform residues below each prime ----
    Q = args.queries
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    cin = [ctx.empty(*ctx.ct_shape(Q, 2, None, engine.LAYOUT_LIMB_MAJOR)) for _ in range(3)]
    for c in cin:
        for j in range(k):
            c[j].random_(0, ctx.q[j], generator=g)
    out = ctx.empty(*ctx.ct_shape(Q, 2, None, engine.LAYOUT_LIMB_MAJOR))
    xb = torch.randint(1, 1 << 27, (Q,), device=dev, generator=g)
    yb = torch.randint(1, 1 << 27, (Q,), device=dev, generator=g)
    rr = torch.randint(0, 1 << 32, (Q,), device=dev, generator=g)
    ss = torch.randint(1, 1 << 32, (Q,), device=dev, generator=g)

    def step():
        ctx.circuit_a(cin[0], cin[1], cin[2], xb, yb, rr, ss, out=out, layout=engine.LAYOUT_LIMB_MAJOR)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    time.sleep(0.3)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t0 = time.time()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    t1 = time.time()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop(t0, t1)
    # sustained leg: the same launch back to back for >= 1 s (the K-step region above is tens of milliseconds), with its own
    # clock record, so the figure is also known in a steady thermal / power state
    sus_n = int(max(args.steps, min(4000, 1.1e3 / max(total_ms / args.steps, 1e-3))))
    sus_sampler = ClockSampler(local)
    sus_sampler.start()
    time.sleep(0.2)
    barrier()
    sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts0 = time.time()
    sa.record()
    for _ in range(sus_n):
        step()
    sb.record()
    barrier()
    ts1 = time.time()
    sus_ms = sa.elapsed_time(sb)
    sus_clocks = sus_sampler.stop(ts0, ts1)
    from pplp_b200.shard import gather_rows, max_over_ranks
    sus_ms_max = max_over_ranks(sus_ms, dev)
    total_ms_max = max_over_ranks(total_ms, dev)   # device time, max over ranks
    value = world * Q * args.steps / (total_ms_max * 1e-3)

    # ---- end to end: host ciphertexts (page-locked) through the C-ABI host entry ----
    Qe = args.e2e_queries
    per_ct = 2 * k * N
    from pplp_b200 import numa
    affinity0 = os.sched_getaffinity(0)
    numa_rep = numa.bind_to_gpu_node(local)     # pages of cudaHostAlloc land on the calling thread's node: bind first
    hc = [numa.pinned_empty(ctx.L, (Qe, 2, k, N), torch.int64, write_combined=True) for _ in range(3)]   # H2D sources
    hout = numa.pinned_empty(ctx.L, (Qe, 2, k, N), torch.int64)
    for t_ in hc + [hout]:
        t_.zero_()                              # first touch while bound
    os.sched_setaffinity(0, affinity0)          # the CPU baseline leg below counts the threads it may use
    for t_, c in zip(hc, cin):   # fill from the device slabs (values < q_j per limb); layout SEAL on the host
        t_.copy_(c[:, :, :Qe, :].permute(2, 1, 0, 3))
    hpar = [x[:Qe].cpu().numpy().view(np.uint64).copy() for x in (xb, yb, rr, ss)]
    e2e_steps = max(2, min(args.steps, 10))

    def e2e_step():
        ctx.circuit_a_host(hc[0], hc[1], hc[2], hout, hpar[0], hpar[1], hpar[2], hpar[3], chunk=128)

    e2e_step()
    barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - te0
    e2e_value = world * Qe * e2e_steps / max_over_ranks(e2e_s, dev)
    # the host entry returned what the resident kernel computes
    ref_slice = out[:, :, :Qe, :].permute(2, 1, 0, 3).cpu()
    if not torch.equal(ref_slice, hout):
        raise SystemExit("bench: host-buffer path and resident path disagree")
    # the host link under the same conditions: all ranks copying at once, both directions, same buffers (the ceiling of e2e)
    link = {}
    sA, sB, sC, sD = (torch.cuda.Stream() for _ in range(4))
    dtmp = [ctx.empty(Qe, 2, k, N) for _ in range(4)]
    # the entry point's own 3 : 1 traffic mix, every copy on its own stream (the entry keeps three slabs in flight)
    mix = [(dtmp[0], hc[0], sA), (dtmp[2], hc[1], sC), (dtmp[3], hc[2], sD), (hout, dtmp[1], sB)]
    for name, ops in (("h2d", [(dtmp[0], hc[0], sA)]), ("d2h", [(hout, dtmp[1], sB)]), ("bidir", [(dtmp[0], hc[0], sA), (hout, dtmp[1], sB)]), ("mix_3in_1out", mix)):
        barrier()
        tl0 = time.perf_counter()
        for _ in range(4):
            for dst_, src_, st_ in ops:
                with torch.cuda.stream(st_):
                    dst_.copy_(src_, non_blocking=True)
        torch.cuda.synchronize()
        dtl = max_over_ranks(time.perf_counter() - tl0, dev)
        link[name] = 4 * Qe * per_ct * 8 / dtl / 1e9 * world          # GB/s per direction (the larger one for the mix: x3)
        if name == "mix_3in_1out":
            link[name] *= 3
            link_ceiling = 4 * Qe * world / dtl                          # queries/s if the entry point did nothing but these copies
    del dtmp

    # ---- the one exchange of the design: results to rank 0, outside the hot-path figure ----
    gather = measure_exchange(torch, dist, ctx, engine, cin, out, xb, yb, rr, ss, Q, k, world, rank, dev, barrier) if world > 1 else None

    extras = {}
    n16384 = None
    if not args.no_extras and N == 8192:
        # BASELINE.json configs[2] (N=16384, sharded over the ranks): every rank evaluates its shard, aggregate = sum over ranks
        # over the slowest rank's device time — the same rule as the headline value
        del cin, out
        torch.cuda.empty_cache()
        try:
            n16384 = extra_n16384(engine, torch, world, dev, barrier, rank)
        except Exception as e:
            n16384 = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    if not args.no_extras and rank == 0:
        extras = run_extras(engine, ctx, torch, osk, opk)
    if n16384 is not None:
        extras["n16384"] = n16384
    if gather:
        extras["gather"] = gather

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    bytes_per_query = 64 * k * N
    avg_kernel_ms = float(np.mean(kernel_ms))
    achieved = bytes_per_query * Q / (avg_kernel_ms * 1e-3) / 1e9
    traffic = None
    pipes = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            pipes = json.load(open(tp)).get("circuit_a_kernel", {})
            traffic = pipes.get("dram_bytes_per_query")
            # ncu capture at a smaller batch (N=8192, k=4): scaled per query, and per ciphertext size, to this launch
            traffic = traffic * Q * (N * k) / (8192 * 4) if traffic else None
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_per_gpu_per_step": Q, "poly_modulus_degree": N, "limbs": k, "plain_modulus": "2^56",
                   "layout": "limb-major [limb][poly][query][N]", "parallelism": f"query-sharded x{world}, no collective in the hot path",
                   "cache": f"inputs {3 * Q * per_ct * 8 / 2**30:.1f} GiB per step >> 126 MB L2 (no flush needed)"},
        "roofline": {"bound": "hbm", "kernel": "circuit_a_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": "from profile: ncu --set full dram__bytes_read+write of this kernel (profiles/traffic.json), scaled per query",
                     "fp64_pipe_pct": pipes.get("fp64_pipe_pct"), "int_pipe_pct": pipes.get("int_pipe_pct"), "dram_pct_of_ncu_peak": pipes.get("dram_pct"),
                     "pipes_source": "from profile (profiles/traffic.json: sm__pipe_fp64_cycles_active / sm__pipe_fma+alu, ncu --set full)", "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_query * Q,
                     "avg_launch_ms": avg_kernel_ms, "note": "event pairs bracket each pplp_circuit_a call (scalar-prepare kernel + main kernel)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * Qe * per_ct * 8 + 4 * Qe * 8, "d2h_bytes_per_step": Qe * per_ct * 8,
                "queries_per_step": Qe, "steps": e2e_steps, "api": "pplp_circuit_a_host (pinned host ciphertexts, SEAL layout)",
                "host_link_GBps_all_ranks": link, "link_ceiling_queries_per_s": link_ceiling,
                "numa": numa_rep,
                "note": "pinned buffers from cudaHostAlloc after binding the rank to its GPU's NUMA node (write-combined H2D sources); "
                        "host_link = the same buffers copied by every rank at once, so e2e / link_ceiling says how much of the host fabric the entry point uses"},
        "gpu_launches": 2 * args.steps,
        "clocks": clocks,
        "sustained": {"value": world * Q * sus_n / (sus_ms_max * 1e-3), "unit": UNIT, "launches": sus_n, "seconds": sus_ms_max * 1e-3,
                      "achieved_GBps_per_gpu": bytes_per_query * Q * sus_n / (sus_ms_max * 1e-3) / 1e9, "clocks": sus_clocks,
                      "note": "same launch repeated for >= 1 s after the K timed steps"},
    }
    if world == 1:
        threads = host_threads()
        rate, per_step, sample, info = cpu_circuit_a_rate(5, 1, args.cpu_sample, threads)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{sample} queries/step x 5 steps in place on host-resident ciphertexts, SEAL-4.1-equivalent CPU "
                                          f"restatement (oracle/), {threads} host threads", **info}
    if extras:
        line["extras"] = extras
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_extras(engine, ctx, torch, osk, opk):
    """Measured numbers for the other kernels of the path (not the headline): full protocol and NTT GB/s."""
    ex = {}
    try:
        sk, pk = ctx.dev(osk), ctx.dev(opk)
        nq, radius = 4096, 128
        rng = np.random.default_rng(99)
        r, s, w = 0x12345678, 0x9ABCDEF1, 0xBEEF
        xb = np.full(nq, 123456888, dtype=np.uint64); yb = np.full(nq, 132465777, dtype=np.uint64)
        xa = xb + rng.integers(0, 300, nq).astype(np.uint64); ya = yb + rng.integers(0, 300, nq).astype(np.uint64)
        seeds = rng.integers(0, 1 << 63, size=(nq * 3, 8), dtype=np.uint64)
        bf = engine.BloomBatch(ctx, radius, fpp=1e-4, rsw=[(r, s, w)]).build()
        ctx.proximity_batch_host(pk, sk, xa, ya, xb, yb, seeds, bf, chunk=2048)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            blind, verdict, flags = ctx.proximity_batch_host(pk, sk, xa, ya, xb, yb, seeds, bf, chunk=2048)
        dt = time.perf_counter() - t0
        d2 = (xa.astype(np.int64) - xb.astype(np.int64)) ** 2 + (ya.astype(np.int64) - yb.astype(np.int64)) ** 2
        expect = (np.uint64(s) * (d2.astype(np.uint64) + np.uint64(r))) & np.uint64(T - 1)
        ex["protocol_e2e"] = {"value": nq * reps / dt, "unit": UNIT, "queries_per_step": nq,
                              "what": "host coords -> 3 encrypts + Circuit A + decrypt + Bloom verdict -> host verdicts (pplp_proximity_batch_host)",
                              "blind_distances_correct": bool((blind == expect).all()), "near_fraction": float(verdict.mean())}
        # the same protocol on the host cores (CPU restatement), a small sample: context for the number above
        from tests import oracle_lib
        from tests.oracle_lib import OracleBloom
        orc = oracle_lib.load()
        octx = orc.context(N, ctx.q, T, seed=seed8(7))
        ob = OracleBloom(orc.lib, "orc", radius * radius, 1e-4)
        tb0 = time.perf_counter()
        ob.insert_blinded_range(r, s, w, radius * radius)
        set_bf_ms = (time.perf_counter() - tb0) * 1e3
        threads = host_threads()
        ns = max(threads, 2 * threads)
        t0 = time.perf_counter()
        cb, cv, stage_ns = octx.protocol_batch(opk, osk, xa[:ns], ya[:ns], xb[:ns], yb[:ns], r, s, w, seeds[: ns * 3], bloom=ob, nthreads=threads)
        cdt = time.perf_counter() - t0
        ex["protocol_cpu"] = {"value": ns / cdt, "unit": UNIT, "cores": threads, "sample": f"{ns} queries", "kind": "port",
                              "agrees_with_gpu": bool((cb == blind[:ns]).all() and (cv == verdict[:ns]).all()),
                              "stage_ms_per_query": {k: float(v) / ns / 1e6 for k, v in zip(["d_enc", "d_homoCalc", "d_dec", "d_bfQuery"], stage_ns)},
                              "d_setBF_ms_per_filter": set_bf_ms, "d_setBF_note": f"radius {radius}: {radius * radius} keys x 13 hashes, one host core (src/server.cc:95-98)"}
    except Exception as e:   # extras never invalidate the headline
        ex["protocol_e2e"] = {"error": str(e)[:200]}
    try:
        rows_q, k = 4096, ctx.k
        data = ctx.empty(k, 1, rows_q, N)
        for j in range(k):
            data[j].random_(0, ctx.q[j])
        for inv in (False, True):
            for _ in range(3):
                ctx.ntt_(data, inverse=inv, layout=engine.LAYOUT_LIMB_MAJOR)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            reps = 10
            for _ in range(reps):
                ctx.ntt_(data, inverse=inv, layout=engine.LAYOUT_LIMB_MAJOR)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            ex["intt_gbs" if inv else "ntt_gbs"] = 16 * N * rows_q * k / (ms * 1e-3) / 1e9
        ex["ntt_note"] = f"16*N bytes per limb transform, {rows_q * k} rows of N={N} per launch"
    except Exception as e:
        ex["ntt_gbs"] = {"error": str(e)[:200]}
    try:   # BASELINE.json config 5 in small: every resident client against a tile of server points (pplp_circuit_a_cross)
        ncl, npts, k = 512, 16, ctx.k
        cts = []
        for _ in range(3):
            c = ctx.empty(k, 2, ncl, N)
            for j in range(k):
                c[j].random_(0, ctx.q[j])
            cts.append(c)
        rng = np.random.default_rng(55)
        par = [ctx.dev(rng.integers(1, 1 << 27, npts, dtype=np.uint64)) for _ in range(2)] + [ctx.dev(rng.integers(1, 1 << 32, npts, dtype=np.uint64)) for _ in range(2)]
        out = ctx.empty(k, 2, ncl * npts, N)
        for _ in range(2):
            ctx.circuit_a_cross(cts[0], cts[1], cts[2], par[0], par[1], par[2], par[3], out=out, layout=engine.LAYOUT_LIMB_MAJOR)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 5
        for _ in range(reps):
            ctx.circuit_a_cross(cts[0], cts[1], cts[2], par[0], par[1], par[2], par[3], out=out, layout=engine.LAYOUT_LIMB_MAJOR)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        ex["config5_cross"] = {"value": ncl * npts / (ms * 1e-3), "unit": "client-point pairs/s", "clients": ncl, "points_per_launch": npts,
                               "hbm_write_gbs": ncl * npts * 16 * k * N / (ms * 1e-3) / 1e9,
                               "what": "pplp_circuit_a_cross: client ciphertexts read once per launch, 16*k*N bytes written per pair"}
    except Exception as e:
        ex["config5_cross"] = {"error": str(e)[:200]}
    for name, fn in (("circuit_b", extra_circuit_b), ("bloom_build", extra_bloom_build), ("config4_sweep", extra_config4)):
        try:
            ex[name] = fn(engine, torch)
        except Exception as e:
            ex[name] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        torch.cuda.empty_cache()
    return ex


def _event_time(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def extra_circuit_b(engine, torch):
    """north_star's direct form at N=8192, slot-batched with a prime t (PlainModulus::Batching(8192, 56)): per ciphertext
    group  s * (relin((cx - px)^2) + relin((cy - py)^2) + r)  through pplp_circuit_b; N slot-wise queries per group.
    Checked against the oracle's call-by-call evaluation on one group, timed next to the oracle on the host cores."""
    import concurrent.futures
    from tests import oracle_lib
    n, t = 8192, 0xfffffffffb4001
    ctx = engine.Context(n, t=t, device=torch.cuda.current_device())
    k = ctx.k
    orc = oracle_lib.load()
    octx = orc.context(n, ctx.q, t, seed=seed8(11))
    osk, opk = octx.keygen()
    ork = octx.relin_keygen(osk)
    rk = ctx.dev(ork)
    quot = ctx.relin_prepare(rk)
    rng = np.random.default_rng(77)
    groups = 512
    enc = lambda v: ctx.batch_encode(ctx.dev(np.ascontiguousarray(v)))
    xb = rng.integers(0, 1 << 27, (groups, n), dtype=np.uint64); yb = rng.integers(0, 1 << 27, (groups, n), dtype=np.uint64)
    rr = rng.integers(0, 1 << 16, (groups, n), dtype=np.uint64)
    px, py, pr = enc(xb), enc(yb), enc(rr)
    sv = ctx.dev(rng.integers(1, 8, groups, dtype=np.uint64))
    xa, ya = 123456789, 132456888
    pa = engine.to_np(enc(np.stack([np.full(n, xa, dtype=np.uint64), np.full(n, ya, dtype=np.uint64)])))
    ex_, ey_ = octx.encrypt(opk, pa[0], seed=seed8(31)), octx.encrypt(opk, pa[1], seed=seed8(32))
    cx = ctx.dev(np.broadcast_to(ex_, (groups,) + ex_.shape).copy()); cy = ctx.dev(np.broadcast_to(ey_, (groups,) + ey_.shape).copy())
    out = ctx.empty(*ctx.ct_shape(groups, 2))
    run = lambda: ctx.circuit_b(cx, cy, px, py, pr, sv, rk, quot, out=out, chunk=256)
    sec = _event_time(torch, run, 4)
    # parity + algebra on group 0
    hp = [engine.to_np(x[0]) for x in (px, py, pr)]
    s0 = int(engine.to_np(sv)[0])

    def oracle_group(a, b, c):
        ox = octx.eval_plain("sub_plain", ex_, a); oy = octx.eval_plain("sub_plain", ey_, b)
        ox2 = octx.relinearize(octx.square(ox), ork); oy2 = octx.relinearize(octx.square(oy), ork)
        return octx.eval_plain("multiply_plain", octx.eval_plain("add_plain", octx.eval_ct("add", ox2, oy2), c), [s0])

    t0 = time.perf_counter()
    od = oracle_group(*hp)
    one = time.perf_counter() - t0
    agrees = bool((engine.to_np(out[0]) == od).all())
    slots = engine.to_np(ctx.batch_decode(ctx.decrypt(out[:1], ctx.dev(osk))))[0]
    d2 = (xa - xb[0].astype(object)) ** 2 + (ya - yb[0].astype(object)) ** 2
    algebra = bool(all(int(v) == (s0 * (int(a) + int(b))) % t for v, a, b in zip(slots[:64], d2[:64], rr[0][:64])))
    dsk = ctx.dev(osk)
    budget = {"fresh_input": int(ctx.noise_budget(cx[:1], dsk)[0].item()), "output_min": int(ctx.noise_budget(out, dsk).min().item()),
              "what": "Decryptor::invariant_noise_budget in bits (pplp_noise_budget), input ciphertext and the minimum over the step's outputs"}
    threads = host_threads()
    with concurrent.futures.ThreadPoolExecutor(threads) as pool:     # ctypes releases the GIL: one group per host thread
        t0 = time.perf_counter()
        list(pool.map(lambda i: oracle_group(*hp), range(threads)))
        cdt = time.perf_counter() - t0
    peak, _ = peaks()
    bytes_group = 280 * n * 8           # planning estimate of compulsory HBM bytes per ciphertext group
    gps = groups / sec
    return {"groups_per_s": gps, "value": gps * n, "unit": "slot-wise queries/s", "queries_per_group": n, "groups_per_step": groups,
            "squares_per_s": 2 * gps, "relinearizations_per_s": 2 * gps,
            "workload": "circuitB_bfv_n8192_k4_t=Batching(8192,56)_slot_batched: sub_plain x2, square x2, relinearize x2, add, add_plain, multiply_plain(mono)",
            "roofline": {"bound": "FP64 pipe (squares over the 44-bit auxiliary base, relinearisation: every transform and base conversion in exact FP64 products); HBM shown for reference", "algorithmic_bytes_per_group": bytes_group,
                         "achieved": bytes_group * gps / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_group * gps / 1e9 / peak},
            "roofline_fp64": _circuit_b_fp64_roofline(gps, n),
            "agrees_with_oracle": agrees, "slots_match_algebra": algebra, "noise_budget_bits": budget,
            "cpu_baseline": {"groups_per_s": threads / cdt, "value": threads / cdt * n, "unit": "slot-wise queries/s", "cores": threads, "kind": "port",
                             "one_thread_seconds_per_group": one, "sample": f"{threads} groups, one per host thread (oracle/ restatement of SEAL's bfv_square + switch_key_inplace)"}}


def _circuit_b_fp64_roofline(gps, n):
    """The pipe that bounds Circuit B: FP64 instructions per coefficient index (counted from the ncu captures under
    profiles/r02_circuit_b_fp64_counts.csv: square 5 100, relinearisation 1 950; by hand 5 400 and 1 900) against
    64 FP64 lanes per clock per SM at the maximum SM clock."""
    per_square, per_relin = 5100, 1950   # sm__inst_executed_pipe_fp64.sum x 32 / (ciphertexts x N), profiles/r02_circuit_b_fp64_counts.csv
    instr_group = (2 * per_square + 2 * per_relin) * n
    try:
        import torch
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    except Exception:
        sms = 148
    try:
        mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
    except Exception:
        mhz = 1965.0
    peak = 64.0 * sms * mhz * 1e6            # FP64 instructions per second
    return {"bound": "fp64_pipe", "fp64_instr_per_coefficient": {"square": per_square, "relinearize": per_relin}, "fp64_instr_per_group": instr_group,
            "achieved": instr_group * gps / 1e12, "peak": peak / 1e12, "unit": "T FP64 instr/s", "frac": instr_group * gps / peak,
            "source": "instruction counts measured with ncu (sm__inst_executed_pipe_fp64.sum, profiles/r02_circuit_b_fp64_counts.csv); peak = 64 lanes x SMs x max SM clock"}


def extra_bloom_build(engine, torch):
    """d_setBF (src/server.cc:83-98, include/benchmark.h): Bloom-filter construction rates, GPU vs the oracle on one host core."""
    from tests import oracle_lib
    from tests.oracle_lib import OracleBloom
    ctx = engine.Context(4096, t=T, device=torch.cuda.current_device())
    orc = oracle_lib.load()
    res = {}
    rng = np.random.default_rng(5)
    for radius, nf, fpp in ((128, 2048, 1e-4), (128, 1, 1e-4), (4096, 1, 1e-4)):
        rsw = np.stack([rng.integers(0, 1 << 32, nf, dtype=np.uint64), rng.integers(1, 1 << 32, nf, dtype=np.uint64), rng.integers(1 << 15, 1 << 16, nf, dtype=np.uint64)], axis=1)
        bf = engine.BloomBatch(ctx, radius, fpp=fpp, rsw=rsw)
        sec = _event_time(torch, lambda: bf.build(), 5)
        ob = OracleBloom(orc.lib, "orc", radius * radius, fpp)
        t0 = time.perf_counter()
        ob.insert_blinded_range(int(rsw[0, 0]), int(rsw[0, 1]), int(rsw[0, 2]), radius * radius)
        cpu = time.perf_counter() - t0
        ok = bool((bf.table_bytes(0) == ob.table()).all())
        res[f"r{radius}_x{nf}"] = {"filters": nf, "keys_per_filter": radius * radius, "hashes_per_key": bf.k, "table_bytes": bf.m_bits // 8, "gpu_ms": sec * 1e3,
                                   "inserts_per_s": nf * radius * radius / sec, "filters_per_s": nf / sec, "d_setBF_cpu_ms_per_filter": cpu * 1e3,
                                   "cpu_inserts_per_s_one_core": radius * radius / cpu, "table_equals_oracle": ok}
    return res


def extra_n16384(engine, torch, world, dev, barrier, rank):
    """BASELINE.json configs[2]: the same Circuit A workload at N=16384 (k=8), query-sharded over all ranks (every rank runs it;
    value = queries of all ranks / slowest rank's device time), plus this rank's NTT figures at that degree."""
    from pplp_b200.shard import max_over_ranks
    n = 16384
    ctx = engine.Context(n, t=T, device=torch.cuda.current_device())
    k, Q = ctx.k, 1024
    cin = [ctx.empty(*ctx.ct_shape(Q, 2, None, engine.LAYOUT_LIMB_MAJOR)) for _ in range(3)]
    for c in cin:
        for j in range(k):
            c[j].random_(0, ctx.q[j])
    out = ctx.empty(*ctx.ct_shape(Q, 2, None, engine.LAYOUT_LIMB_MAJOR))
    par = [torch.randint(1, 1 << 27, (Q,), device=out.device) for _ in range(2)] + [torch.randint(1, 1 << 32, (Q,), device=out.device) for _ in range(2)]
    run = lambda: ctx.circuit_a(cin[0], cin[1], cin[2], par[0], par[1], par[2], par[3], out=out, layout=engine.LAYOUT_LIMB_MAJOR)
    for _ in range(3):
        run()
    reps = 20
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record()
    barrier()
    sec = max_over_ranks(a.elapsed_time(b), dev) * 1e-3 / reps
    peak, _ = peaks()
    gbs = 64 * k * n * Q / sec / 1e9
    res = {"value": world * Q / sec, "unit": UNIT, "n_gpus": world, "workload": f"circuitA_bfv_n{n}_k{k}_t2^56_batched", "queries_per_gpu_per_step": Q,
           "achieved_GBps_per_gpu": gbs, "frac_of_hbm_peak": gbs / peak}
    del cin, out
    if rank == 0:
        rows = 2048
        data = ctx.empty(k, 1, rows, n)
        for j in range(k):
            data[j].random_(0, ctx.q[j])
        for inv in (False, True):
            s_ = _event_time(torch, lambda: ctx.ntt_(data, inverse=inv, layout=engine.LAYOUT_LIMB_MAJOR), 5)
            res["intt_gbs" if inv else "ntt_gbs"] = 16 * n * rows * k / s_ / 1e9
    torch.cuda.empty_cache()
    return res


def extra_config4(engine, torch):
    """BASELINE.json configs[3], reduced grid: NTT / INTT / relinearize / square at N = 4096..32768 x {3, 8} limbs."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import config4_sweep
    return config4_sweep.sweep(limb_list=(3, 8), device=torch.cuda.current_device(), verbose=False, reps=3)


if __name__ == "__main__":
    a = parse_args()
    N = a.degree
    _k = {4096: 2, 8192: 4, 16384: 8, 32768: 15}.get(N)
    if _k is None:
        raise SystemExit("bench: --degree must be 4096, 8192, 16384 or 32768")
    LIMBS = _k
    WORKLOAD = f"circuitA_bfv_n{N}_k{_k}_t2^56_batched"
    if a.queries <= 0:
        a.queries = max(256, 8192 * 8192 * 4 // (N * _k))       # same bytes per step as 8192 queries at N=8192
    a.e2e_queries = max(64, a.e2e_queries * 8192 * 4 // (N * _k))
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
